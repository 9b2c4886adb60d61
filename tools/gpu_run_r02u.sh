# round 2, call u: final state -- full GPU suite, benches (C3 default with CPU baseline, C2, C5 on one GPU, orbit scene, PnP),
# ncu launch list + --set full of the main kernels
O=gpurun_out/r02u; mkdir -p $O
nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --impl reference > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err; echo "bench ref rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --config C2 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --config C5 --no-cpu-baseline > $O/bench_c5_1gpu.json 2> $O/bench_c5_1gpu.err; echo "bench c5 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense.json 2> $O/bench_c3dense.err; echo "dense rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_pnp.py > $O/bench_pnp.json 2> $O/bench_pnp.err; echo "pnp rc=$?" >> $O/rc.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'schur_syrk|point_pass|frame_pass|point_step|k3_dag|k1_kernel' -s 12 -c 12 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_gpu.txt
for f in bench_c3 bench_c2 bench_c5_1gpu bench_c3dense; do python - $O/$f.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms'].items() if v}, 'frac', round(d['roofline']['frac'],3), d.get('cpu_baseline'))
P
done
cat $O/bench_c3_reference.json | cut -c1-400
