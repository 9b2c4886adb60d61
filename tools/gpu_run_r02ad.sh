# round 2, call ad: point pass with L2 prefetch one resident wave ahead: full GPU suite, unfused subset, smoke, bench
O=gpurun_out/r02ad; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
RSBA_CUDA_FUSED=0 timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_edge.py tests/test_gpu_priors.py -m gpu -x -q > $O/pytest_unfused.txt 2>&1; echo "pytest unfused rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 600 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_gpu.txt; tail -2 $O/pytest_unfused.txt
python - $O/bench_default.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, 'frac', round(d['roofline']['frac'],3), d['gpu_launches'], d['lm']['final_cost'])
P
