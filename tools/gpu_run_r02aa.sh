# round 2, call aa: regression after the FACTOR panel-store fix (full GPU suite, unfused subset, smoke, default bench with CPU
# baseline, reference arm, K3 trace, ncu launch list)
O=gpurun_out/r02aa; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" >> $O/rc.txt
RSBA_CUDA_FUSED=0 timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_edge.py tests/test_gpu_priors.py -m gpu -x -q > $O/pytest_unfused.txt 2>&1; echo "pytest unfused rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace.npz > $O/k3_trace_band.json 2> $O/trace.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_gpu.txt; tail -2 $O/pytest_unfused.txt; tail -1 $O/smoke.txt
python - $O/bench_default.json $O/bench_reference.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, 'frac', round(d['roofline']['frac'],3), d['gpu_launches'], d['e2e']['value'], d['e2e'].get('cold_call_ms_total'), d['cpu_baseline'])
r=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(r.get('value'), r.get('cpu_baseline'))
P
