mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
./tools/fp64_latency > $O/fp64_latency.txt 2>&1; echo "lat rc=$?" >> $O/rc.txt
timeout 300 python tools/trace_k3.py 125 3 4 $O/trace_band.npz > $O/trace_band.json 2> $O/trace_band.err; echo "trace rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/fp64_latency.txt; cat $O/trace_band.json
