O=gpurun_out/r02p; mkdir -p $O
nproc > $O/nproc.txt
RSBA_CUDA_TRACE=1 timeout 300 python tools/cold_call.py C3 10 3 > $O/cold_fused.txt 2>&1
RSBA_CUDA_FUSED=0 RSBA_CUDA_TRACE=1 timeout 300 python tools/cold_call.py C3 10 3 > $O/cold_unfused.txt 2>&1
grep -v "k3\|task " $O/cold_fused.txt | tail -40; echo ----; grep "cold call" $O/cold_unfused.txt
