# round 2, call 4 (8 GPUs): the W=8 and W=4 sweeps (every algorithm, NCCL, symm_mem), a small tune grid, ResNet-50 A/B, BERT / GPT-2 A/B
mkdir -p gpurun_out
S=gpurun_out/r2_4_summary.txt; : > $S
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
B2D_ITERS=20 timeout 400 $TR8 --master-port 29911 tools/microbench.py sweep > gpurun_out/r2_4_sweep_8.jsonl 2> gpurun_out/r2_4_sweep_8.err; echo "sweep8 rc $?" >> $S
B2D_TUNE_CHUNKS=8,16,32,64 B2D_TUNE_CTAS=32,64,128 timeout 400 $TR8 --master-port 29912 tools/microbench.py tune > gpurun_out/r2_4_tune_8.jsonl 2> gpurun_out/r2_4_tune_8.err; echo "tune8 rc $?" >> $S
B2D_ITERS=20 timeout 400 $TR4 --master-port 29913 tools/microbench.py sweep > gpurun_out/r2_4_sweep_4.jsonl 2> gpurun_out/r2_4_sweep_4.err; echo "sweep4 rc $?" >> $S
timeout 600 $TR8 --master-port 29914 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_4_bench_n8.json 2> gpurun_out/r2_4_bench_n8.err; echo "bench n8 rc $?" >> $S
timeout 400 $TR8 --master-port 29915 bench.py --gpus 8 --steps 20 --warmup 5 --hook nccl_bf16 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_4_bench_n8_nccl.json 2> gpurun_out/r2_4_bench_n8_nccl.err; echo "bench n8 nccl rc $?" >> $S
timeout 400 $TR8 --master-port 29916 bench.py --gpus 8 --steps 20 --warmup 5 --wire fp32 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_4_bench_n8_fp32.json 2> gpurun_out/r2_4_bench_n8_fp32.err; echo "bench n8 fp32 arena rc $?" >> $S
for m in bert-base gpt2-medium; do
  timeout 400 $TR8 --master-port 29917 bench.py --gpus 8 --steps 10 --warmup 3 --model $m --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_4_${m}_b200.json 2> gpurun_out/r2_4_${m}_b200.err; echo "$m b200 rc $?" >> $S
  timeout 400 $TR8 --master-port 29918 bench.py --gpus 8 --steps 10 --warmup 3 --model $m --hook nccl_bf16 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_4_${m}_nccl.json 2> gpurun_out/r2_4_${m}_nccl.err; echo "$m nccl rc $?" >> $S
done
timeout 400 $TR8 --master-port 29919 bench.py --gpus 8 --steps 10 --warmup 3 --model gpt2-medium --strategy sharded --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_4_gpt2_sharded.json 2> gpurun_out/r2_4_gpt2_sharded.err; echo "gpt2 sharded rc $?" >> $S
cat $S
