# round 2, call v: SYRK launch order by anti-diagonals (L2 working set)
O=gpurun_out/r02v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_k1.py tests/test_gpu_multi.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for m in 0 1; do
  RSBA_CUDA_SYRK_ORDER=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_order$m.json 2> $O/bench_c3_order$m.err; echo "bench $m rc=$?" >> $O/rc.txt
  RSBA_CUDA_SYRK_ORDER=$m timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:schur_syrk -s 4 -c 2 --csv --log-file $O/syrk_dram_order$m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -2 $O/pytest.txt; tail -1 $O/smoke.txt
for m in 0 1; do python - $O/bench_c3_order$m.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['kernel_ms']['schur_syrk'], d['e2e']['cold_call_ms_total'], d['e2e'].get('cold_call_ms_first_in_process'))
P
grep -v "^==" $O/syrk_dram_order$m.csv | tail -8 | cut -d, -f5,13-16
done
