mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace_band.npz > $O/trace_band.json 2> $O/trace_band.err; echo "trace rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 48 dense 16 $O/trace_dense.npz > $O/trace_dense.json 2> $O/trace_dense.err; echo "trace dense rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_k3.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_dag.json 2> $O/bench_dag.err; echo "bench dag rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest.txt; cat $O/trace_band.json
