# round 2, call t: K3 with the Dinv / z flag split and the panel barrier -- parity, trace, racecheck of K3 alone, bench
O=gpurun_out/r02t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_k3.py tests/test_gpu_lm.py tests/test_gpu_zz_full_size.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace_band.npz > $O/trace_band.json 2> $O/trace_band.err; echo "trace rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool racecheck python tools/bench_k3.py 12 3 > $O/racecheck_k3.log 2>&1; echo "racecheck k3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -2 $O/pytest.txt; grep -E "RACECHECK SUMMARY" $O/racecheck_k3.log; grep "Error: Race" $O/racecheck_k3.log | sed -E 's/.* in ([a-z0-9_]+\.cu:[0-9]+).*/\1/' | sort | uniq -c | head
python - $O/trace_band.json $O/bench_c3.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms'], d['factor']['run_us_median'], d['trsm'], d['factor_phase_cycles_median'])
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), d['kernel_ms']['cholesky'], d['lm']['final_cost'])
P
