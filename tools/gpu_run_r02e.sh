mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
timeout 300 python -m pytest tests/test_gpu_k3.py -x -q > $O/pytest_k3.txt 2>&1; echo "k3 rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace_band.npz > $O/trace_band.json 2> $O/trace_band.err; echo "trace rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_gpu_lm.py tests/test_gpu_handler.py tests/test_gpu_pose_priors.py tests/test_gpu_priors.py tests/test_gpu_uncalibrated.py -x -q > $O/pytest_lm.txt 2>&1; echo "lm rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_dag.json 2> $O/bench_dag.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_k3.txt; tail -3 $O/pytest_lm.txt; cat $O/trace_band.json
