# round 2, call 10 (1 GPU): ncu of the sharded-path kernels (K11-K13) and of K14 on loopback ranks
mkdir -p gpurun_out
timeout 300 python tools/microbench.py ncu_loopback_sharded > gpurun_out/r2_10_plain.log 2>&1; echo "plain rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seg_stage_kernel|seg_reduce_kernel|adam_push_kernel|bucket_optim_kernel" -s 4 -c 8 -o gpurun_out/r2_10_prof_owner python tools/microbench.py ncu_loopback_sharded > gpurun_out/r2_10_ncu_full.log 2>&1; echo "ncu full rc $?"
ncu -i gpurun_out/r2_10_prof_owner.ncu-rep --page raw --csv > gpurun_out/r2_10_prof_owner_raw.csv 2>/dev/null; echo "export rc $?"
tail -3 gpurun_out/r2_10_plain.log; tail -3 gpurun_out/r2_10_ncu_full.log
