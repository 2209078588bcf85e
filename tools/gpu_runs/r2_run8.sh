# round 2, call 8 (8 GPUs, short): ResNet-50 with the 64-register NVLS exchange kernel, and with the optimizer in backward
mkdir -p gpurun_out
S=gpurun_out/r2_8_summary.txt; : > $S
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR8 --master-port 29914 bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_8_bench_n8.json 2> gpurun_out/r2_8_bench_n8.err; echo "bench n8 rc $?" >> $S
timeout 300 $TR8 --master-port 29916 bench.py --gpus 8 --steps 20 --warmup 5 --optimizer-in-backward --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_8_bench_n8_inbw.json 2> gpurun_out/r2_8_bench_n8_inbw.err; echo "bench n8 optimizer-in-backward rc $?" >> $S
cat $S; tail -c 300 gpurun_out/r2_8_bench_n8_inbw.err
