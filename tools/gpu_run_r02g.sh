# round 2, call g: compact Jacobian records + DMMA-Gram frame_blocks
O=gpurun_out/r02g; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'frame_blocks_kernel|point_step|point_blocks|k1_kernel' -s 8 -c 8 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_gpu.txt; cat $O/bench_c3.json | cut -c1-900
