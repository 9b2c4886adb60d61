# round 2, call f: full GPU suite, C3 bench, ncu launch list, ncu --set full of the big kernels
O=gpurun_out/r02f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 5 --warmup 3 --config C3dense --no-cpu-baseline > $O/bench_c3dense.json 2> $O/bench_c3dense.err; echo "dense rc=$?" >> $O/rc.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'schur_syrk|frame_blocks_kernel|point_step|point_blocks|k3_dag|k1_kernel' -s 12 -c 12 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_gpu.txt; cat $O/bench_c3.json | cut -c1-1500
