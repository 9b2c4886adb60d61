# round 2, call o: point pass split into eval + solve kernels
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
RSBA_CUDA_KP_OCC=5 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_occ5.json 2> $O/bench_c3_occ5.err; echo "bench occ5 rc=$?" >> $O/rc.txt
RSBA_CUDA_KP=fused timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_kpfused.json 2> $O/bench_c3_kpfused.err; echo "bench kpfused rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'point_eval|point_pass|frame_pass|schur_syrk' -s 8 -c 8 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -4 $O/pytest_gpu.txt
for f in bench_c3 bench_c3_occ5 bench_c3_kpfused; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'])
P
done
