# round 2, call m: SYRK pipeline-shape variants; session tests
O=gpurun_out/r02n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_k3.py -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for v in 0 5 6 7; do
  RSBA_CUDA_SYRK_VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_v$v.json 2> $O/bench_c3_v$v.err; echo "bench v$v rc=$?" >> $O/rc.txt
done
cat $O/rc.txt; tail -4 $O/pytest_gpu.txt
for v in 0 5 6 7; do python - $O/bench_c3_v$v.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'])
P
done
