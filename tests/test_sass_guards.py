"""Static guards on the SASS of the built library (no GPU needed: `cuobjdump -sass` reads the cubin).

Two performance bugs of round 2 were invisible in the source and obvious in the SASS: "lane k stores element k of a
register array" compiles to a jump table (`LDC` + `BRX`) whose targets the warp walks one divergent path at a time -- 800 of
the 1 850 cycles of K3's FACTOR panel and 9 % of the point pass (profiles/r02_notes.md).  This test keeps them out, and
checks that the kernels DESIGN.md section 4 describes as DMMA / TMA kernels really contain those instructions.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rsba_b200", "lib", "librsba_cuda.so")


@pytest.fixture(scope="module")
def sass_by_kernel():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        pytest.skip("librsba_cuda.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    assert kernels, "no kernels found in the library"
    return kernels


def _count(lines, mnemonic):
    pat = re.compile(r"\b" + re.escape(mnemonic) + r"\b")
    return sum(1 for l in lines if pat.search(l))


def _find(kernels, fragment):
    hits = {k: v for k, v in kernels.items() if fragment in k}
    assert hits, f"no kernel matching {fragment!r} in the library"
    return hits


def test_no_jump_tables_outside_the_k3_task_switch(sass_by_kernel):
    offenders = {k: _count(v, "BRX") for k, v in sass_by_kernel.items() if _count(v, "BRX")}
    allowed = {k: n for k, n in offenders.items() if "k3_dag_kernel" in k}
    others = {k: n for k, n in offenders.items() if k not in allowed}
    assert not others, f"indirect branches (jump tables) in: {others}"
    # the task-type switch of the persistent kernel (FACTOR / TRSM / UPDATE / BACKTILE / BACKFIN)
    assert all(n <= 2 for n in allowed.values()), allowed


def test_hot_kernels_have_no_local_memory_beyond_the_libm_slow_paths(sass_by_kernel):
    # sincos' Payne-Hanek slow path keeps a small stack frame (two LDL / two STL); anything above that would be a
    # dynamically indexed register array or a spill storm in a kernel DESIGN.md times
    for frag in ("point_pass_kernel", "frame_pass_kernel", "schur_syrk_kernel", "k3_dag_kernel", "point_step_group_kernel"):
        for k, lines in _find(sass_by_kernel, frag).items():
            n = sum(1 for l in lines if "LDL" in l)
            assert n <= 12, (k, n)


@pytest.mark.parametrize("fragment", ["schur_syrk_kernel", "k3_dag_kernel", "frame_pass_kernel"])
def test_fp64_tensor_path(sass_by_kernel, fragment):
    for k, lines in _find(sass_by_kernel, fragment).items():
        assert sum(1 for l in lines if "DMMA" in l) > 0, f"{k}: no DMMA (mma.sync.m8n8k4.f64)"


def test_tma_bulk_copies(sass_by_kernel):
    # the SYRK's panel ring (cp.async.bulk global -> shared, mbarrier completion) and K1's tile store
    for k, lines in _find(sass_by_kernel, "schur_syrk_kernel").items():
        assert sum(1 for l in lines if "UBLKCP" in l) > 0, f"{k}: no bulk copy"
        assert sum(1 for l in lines if "SYNCS" in l) > 0, f"{k}: no mbarrier instructions"
    k1_jac = {k: v for k, v in _find(sass_by_kernel, "k1_kernel").items() if "ILb1E" in k}
    assert k1_jac
    for k, lines in k1_jac.items():
        assert sum(1 for l in lines if "UBLKCP" in l) > 0, f"{k}: no bulk store"
