# round 2, call i: SYRK by Cholesky-tile pairs (k2_schur2.cu), packed point-pass records
O=gpurun_out/r02i; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
RSBA_CUDA_SYRK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_syrk1.json 2> $O/bench_c3_syrk1.err; echo "bench syrk1 rc=$?" >> $O/rc.txt
RSBA_CUDA_KP_OCC=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_kp4.json 2> $O/bench_c3_kp4.err; echo "bench kp4 rc=$?" >> $O/rc.txt
RSBA_CUDA_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_unfused.json 2> $O/bench_c3_unfused.err; echo "bench unfused rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'schur_syrk2|point_pass|frame_pass' -s 6 -c 6 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_gpu.txt
for f in bench_c3 bench_c3_syrk1 bench_c3_kp4 bench_c3_unfused; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, d['lm']['final_cost'])
P
done
