# round 2, call r: K3 FACTOR with the trailing update kept off the panel warps' sub-partitions
O=gpurun_out/r02r; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_k3.py -x -q > $O/pytest_k3.txt 2>&1; echo "k3 rc=$?" >> $O/rc.txt
for m in 0 1; do
  RSBA_CUDA_K3_PANEL=$m timeout 300 python tools/trace_k3.py 125 3 4 $O/trace_band_$m.npz > $O/trace_band_$m.json 2> $O/trace_band_$m.err; echo "trace $m rc=$?" >> $O/rc.txt
  RSBA_CUDA_K3_PANEL=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_$m.json 2> $O/bench_c3_$m.err; echo "bench $m rc=$?" >> $O/rc.txt
done
RSBA_CUDA_K3_PANEL=1 timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense_1.json 2> $O/bench_c3dense_1.err
cat $O/rc.txt; tail -2 $O/pytest_k3.txt
for m in 0 1; do python - $O/trace_band_$m.json $O/bench_c3_$m.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms'], d['factor']['run_us_median'], d['factor_phase_cycles_median'])
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), d['kernel_ms']['cholesky'])
P
done
python - $O/bench_c3dense_1.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],3), d['kernel_ms']['cholesky'])
P
