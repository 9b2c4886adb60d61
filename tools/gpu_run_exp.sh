O=gpurun_out/r02exp5; mkdir -p $O
timeout 120 tools/factor_ablation > $O/ablation.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_k3.py -x -q > $O/pytest_k3.txt 2>&1; echo "k3 rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace.npz > $O/trace.json 2> $O/trace.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -2 $O/pytest_k3.txt
python - $O/trace.json $O/bench_c3.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(d['ms'], d['factor']['run_us_median'], d['factor_phase_cycles_median'])
for f in sys.argv[2:]:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d['kernel_ms']['cholesky'], d['lm']['final_cost'])
P
