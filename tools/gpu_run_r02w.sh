# round 2, call w: point pass + back-substitution with one thread per observation over groups of whole points
O=gpurun_out/r02w; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
RSBA_CUDA_KP=warp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_warp.json 2> $O/bench_c3_warp.err; echo "bench warp rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'point_group|point_step_group|frame_pass' -s 6 -c 6 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -4 $O/pytest_gpu.txt
for f in bench_c3 bench_c3_warp; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'])
P
done
