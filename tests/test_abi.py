"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/rsba_cuda.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import rsba_b200.api as api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rsba_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsba_cuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(api.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(api.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rsba_cuda.h but not exported"
    assert sorted(api.SYMBOLS) == declared


def test_struct_layouts_match_header():
    lib = api.load_library()
    o = api.default_options()
    assert o.max_num_iterations == 50 and o.initial_trust_region_radius == 1e4
    assert o.min_relative_decrease == 1e-3 and o.function_tolerance == 1e-6
    assert o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.jacobi_scaling == 1 and o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert lib.rsba_cuda_version().startswith(b"rsba_b200")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.RsbaError) as e:
        api.Problem(0)
    assert e.value.code == api.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under rsba_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rsba_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "librsba_ref" not in src, f
