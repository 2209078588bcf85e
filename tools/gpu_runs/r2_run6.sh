# round 2, call 6 (2 GPUs): re-run of the multi-process tests (buffer hook, optimizer in backward), W=2 sweep with the final AUTO tables, bench N=2
mkdir -p gpurun_out
S=gpurun_out/r2_6_summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_multiprocess.py tests/test_gpu_allreduce.py tests/test_gpu_sharded.py -q -m gpu --timeout 600 > gpurun_out/r2_6_tests.log 2>&1; echo "gpu tests rc $?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29911 tools/microbench.py sweep > gpurun_out/r2_6_sweep_2.jsonl 2> gpurun_out/r2_6_sweep_2.err; echo "sweep rc $?" >> $S
timeout 900 $TR --master-port 29913 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_6_bench_n2.json 2> gpurun_out/r2_6_bench_n2.err; echo "bench n2 rc $?" >> $S
timeout 900 $TR --master-port 29914 bench.py --gpus 2 --steps 20 --warmup 5 --wire fp32 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_6_bench_n2_fp32.json 2> gpurun_out/r2_6_bench_n2_fp32.err; echo "bench n2 fp32 rc $?" >> $S
cat $S; tail -12 gpurun_out/r2_6_tests.log | cut -c1-400
