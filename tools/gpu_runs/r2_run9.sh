# round 2, call 9 (1 GPU): whole gpu suite on the final tree, smoke (plain + ncu), bench N=1 (default, optimizer in backward, GPT-2 sharded for the exposed-comm baseline)
mkdir -p gpurun_out
S=gpurun_out/r2_9_summary.txt; : > $S
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_9_tests.log 2>&1; echo "gpu tests rc $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_9_smoke.log 2>&1; echo "smoke rc $?" >> $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_9_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/r2_9_smoke_ncu.log 2>&1; echo "smoke under ncu rc $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_9_bench_n1.json 2> gpurun_out/r2_9_bench_n1.err; echo "bench rc $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 --optimizer-in-backward --no-cpu-baseline > gpurun_out/r2_9_bench_n1_inbw.json 2> gpurun_out/r2_9_bench_n1_inbw.err; echo "bench in-backward rc $?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --model gpt2-medium --strategy sharded --no-cpu-baseline > gpurun_out/r2_9_bench_n1_gpt2_sharded.json 2> gpurun_out/r2_9_bench_n1_gpt2_sharded.err; echo "bench gpt2 sharded n1 rc $?" >> $S
cat $S; tail -6 gpurun_out/r2_9_tests.log | cut -c1-300; tail -2 gpurun_out/r2_9_smoke.log
