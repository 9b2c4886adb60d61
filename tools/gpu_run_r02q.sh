# round 2, call q: pooled device allocations (cold call), Armijo diagnostic, full GPU suite
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
RSBA_CUDA_TRACE=1 timeout 300 python tools/cold_call.py C3 10 4 > $O/cold_pool.txt 2>&1; echo "cold rc=$?" >> $O/rc.txt
RSBA_CUDA_NO_POOL=1 timeout 300 python tools/cold_call.py C3 10 4 > $O/cold_nopool.txt 2>&1; echo "cold nopool rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -4 $O/pytest_gpu.txt; grep "cold call" $O/cold_pool.txt; echo; grep "cold call" $O/cold_nopool.txt; grep "upload + alloc" $O/cold_pool.txt
python - $O/bench_c3.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['e2e'])
P
