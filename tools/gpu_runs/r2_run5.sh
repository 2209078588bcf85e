# round 2, call 5 (1 GPU): the whole gpu suite on the current tree, smoke, ncu --set full of the staged kernels (loopback), bench N=1
mkdir -p gpurun_out
S=gpurun_out/r2_5_summary.txt; : > $S
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_5_tests.log 2>&1; echo "gpu tests rc $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_5_smoke.log 2>&1; echo "smoke rc $?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stage_kernel|exch_kernel|unstage_kernel" -s 12 -c 9 -o gpurun_out/r2_5_prof_staged python tools/microbench.py ncu_loopback > gpurun_out/r2_5_ncu_full.log 2>&1; echo "ncu full rc $?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_5_loopback_launches.csv python tools/microbench.py ncu_loopback > gpurun_out/r2_5_ncu_list.log 2>&1; echo "ncu list rc $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_5_bench_n1.json 2> gpurun_out/r2_5_bench_n1.err; echo "bench rc $?" >> $S
cat $S; tail -6 gpurun_out/r2_5_tests.log; tail -2 gpurun_out/r2_5_smoke.log
