# round 2, call ae: ncu --set full of the final kernels (two filtered passes over the same bench command)
O=gpurun_out/r02ae; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'schur_syrk|point_pass|frame_pass|point_step_group|k3_dag' -s 10 -c 10 -o $O/full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k1_kernel' -s 12 -c 6 -o $O/full_k1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full_k1.log 2>&1; echo "ncu k1 rc=$?" >> $O/rc.txt
cat $O/rc.txt; ls -la $O
