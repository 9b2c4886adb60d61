# round 2, call s: compute-sanitizer on the small end-to-end run (fused passes, K-split SYRK, DAG K3, pool off and on)
O=gpurun_out/r02s; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/rc.txt
RSBA_CUDA_NO_POOL=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/memcheck_nopool.log 2>&1; echo "memcheck nopool rc=$?" >> $O/rc.txt
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/rc.txt
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_small.py > $O/initcheck.log 2>&1; echo "initcheck rc=$?" >> $O/rc.txt
cat $O/rc.txt; for f in memcheck memcheck_nopool racecheck initcheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|cost" $O/$f.log | sort | uniq -c | sort -rn | head -8; done
