# round 2, call 11 (2 GPUs): sharded tests + GPT-2 sharded N=2 after the adam_push grid change; N=2 with the optimizer in backward
mkdir -p gpurun_out
S=gpurun_out/r2_11_summary.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_strategy.py tests/test_gpu_multiprocess.py -q -m gpu --timeout 600 > gpurun_out/r2_11_tests.log 2>&1; echo "gpu tests rc $?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29915 bench.py --gpus 2 --steps 10 --warmup 3 --model gpt2-medium --strategy sharded --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_11_bench_n2_gpt2_sharded.json 2> gpurun_out/r2_11_bench_n2_gpt2_sharded.err; echo "bench n2 gpt2 sharded rc $?" >> $S
timeout 400 $TR --master-port 29916 bench.py --gpus 2 --steps 20 --warmup 5 --optimizer-in-backward --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_11_bench_n2_inbw.json 2> gpurun_out/r2_11_bench_n2_inbw.err; echo "bench n2 in-backward rc $?" >> $S
cat $S; tail -4 gpurun_out/r2_11_tests.log | cut -c1-300
