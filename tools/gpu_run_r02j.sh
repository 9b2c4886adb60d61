# round 2, call j: SYRK2 with rotated roles vs SYRK1, in-loop kernel timers, dense case
O=gpurun_out/r02j; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_lm.py tests/test_gpu_edge.py tests/test_gpu_uncalibrated.py -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
RSBA_CUDA_KP_OCC=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
RSBA_CUDA_KP_OCC=4 RSBA_CUDA_SYRK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_syrk1.json 2> $O/bench_c3_syrk1.err; echo "bench syrk1 rc=$?" >> $O/rc.txt
RSBA_CUDA_KP_OCC=4 RSBA_CUDA_SYRK=1 RSBA_CUDA_FINE_TIMERS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_fine.json 2> $O/bench_c3_fine.err; echo "bench fine rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense.json 2> $O/bench_c3dense.err; echo "dense rc=$?" >> $O/rc.txt
RSBA_CUDA_SYRK=1 timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense_syrk1.json 2> $O/bench_c3dense_syrk1.err; echo "dense syrk1 rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_gpu.txt
for f in bench_c3 bench_c3_syrk1 bench_c3_fine bench_c3dense bench_c3dense_syrk1; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['stage_ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d.get('inloop_stage_ms_last_iteration'))
P
done
