# round 2, call l: SYRK with K-split two-patch items, session / multi handler tests, last linearisation without S
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense.json 2> $O/bench_c3dense.err; echo "dense rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_gpu.txt
for f in bench_c3 bench_c3dense; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['stage_ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'], d['roofline']['frac'])
P
done
