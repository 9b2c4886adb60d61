mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_k3.py -x -q > gpurun_out/r02a/pytest_k3.txt 2>&1; echo "k3 rc=$?" >> gpurun_out/r02a/rc.txt
timeout 300 python tools/bench_k3.py 125 3 > gpurun_out/r02a/bench_k3_band.json 2> gpurun_out/r02a/bench_k3_band.err; echo "bk3 rc=$?" >> gpurun_out/r02a/rc.txt
timeout 300 python tools/bench_k3.py 48 dense > gpurun_out/r02a/bench_k3_dense48.json 2> gpurun_out/r02a/bench_k3_dense48.err; echo "bk3d rc=$?" >> gpurun_out/r02a/rc.txt
timeout 900 python -m pytest tests/test_gpu_lm.py tests/test_gpu_uncalibrated.py tests/test_gpu_priors.py -x -q > gpurun_out/r02a/pytest_lm.txt 2>&1; echo "lm rc=$?" >> gpurun_out/r02a/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a/bench_dag.json 2> gpurun_out/r02a/bench_dag.err; echo "bench dag rc=$?" >> gpurun_out/r02a/rc.txt
RSBA_CUDA_K3=levels timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a/bench_levels.json 2> gpurun_out/r02a/bench_levels.err; echo "bench levels rc=$?" >> gpurun_out/r02a/rc.txt
cat gpurun_out/r02a/rc.txt; tail -5 gpurun_out/r02a/pytest_k3.txt; cat gpurun_out/r02a/bench_k3_band.json
