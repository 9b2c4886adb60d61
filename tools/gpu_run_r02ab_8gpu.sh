# round 2, call ab (8 GPUs), after the FACTOR panel-store fix: multi-GPU parity tests, C3 at N = 2 / 8, C5 at N = 8 (BASELINE.json configs[3], configs[4])
O=gpurun_out/r02ab; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.txt 2>&1; echo "pytest multi rc=$?" >> $O/rc.txt
run() { # N config steps port
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $1 --steps $3 --warmup 3 --config $2 --no-cpu-baseline > $O/bench_$2_$1gpu.json 2> $O/bench_$2_$1gpu.err; echo "bench $2 N=$1 rc=$?" >> $O/rc.txt
}
run 8 C3 10 29611
run 8 C5 10 29612
run 4 C3 10 29614
run 2 C3 10 29613
cat $O/rc.txt; tail -3 $O/pytest_multi.txt
for f in $O/bench_*gpu.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],1), d['stage_ms_per_step'], d.get('parity_vs_1gpu'))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
P
done
