# round 2, call y: int8 MMA throughput (Ozaki sizing), final regression: full GPU suite + smoke + bench both arms
O=gpurun_out/r02y; mkdir -p $O
timeout 120 ./tools/imma_peak > $O/imma_peak.txt 2>&1; echo "imma rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/imma_peak.txt; tail -3 $O/pytest_gpu.txt; tail -1 $O/smoke.txt
python - $O/bench_default.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], d['e2e']['cold_call_ms_total'])
P
