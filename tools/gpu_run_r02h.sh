# round 2, call h: fused point / frame passes (k2_fused.cu)
O=gpurun_out/r02h; mkdir -p $O
nproc > $O/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
RSBA_CUDA_FUSED=0 timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_priors.py tests/test_gpu_edge.py -m gpu -x -q > $O/pytest_gpu_unfused.txt 2>&1; echo "pytest unfused rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python tools/stage_gaps.py C3 > $O/stage_gaps.txt 2>&1; echo "gaps rc=$?" >> $O/rc.txt
RSBA_CUDA_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_trace.json 2> $O/bench_trace.err; echo "trace rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'point_pass|frame_pass|point_step' -s 6 -c 6 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_gpu.txt; tail -3 $O/pytest_gpu_unfused.txt; cat $O/stage_gaps.txt; cat $O/bench_c3.json | cut -c1-700
