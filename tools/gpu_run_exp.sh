O=gpurun_out/r02exp13; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_lm.py tests/test_gpu_edge.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -2 $O/pytest.txt
python - $O/bench_c3.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'])
P
