# round 2, call 7 (8 GPUs): final sweeps (W=8, W=4), ResNet-50 A/B (+ optimizer in backward), BERT with 75 buckets, GPT-2 sharded
mkdir -p gpurun_out
S=gpurun_out/r2_7_summary.txt; : > $S
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
B2D_ITERS=20 timeout 400 $TR8 --master-port 29911 tools/microbench.py sweep > gpurun_out/r2_7_sweep_8.jsonl 2> gpurun_out/r2_7_sweep_8.err; echo "sweep8 rc $?" >> $S
B2D_ITERS=20 timeout 400 $TR4 --master-port 29913 tools/microbench.py sweep > gpurun_out/r2_7_sweep_4.jsonl 2> gpurun_out/r2_7_sweep_4.err; echo "sweep4 rc $?" >> $S
timeout 600 $TR8 --master-port 29914 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_7_bench_n8.json 2> gpurun_out/r2_7_bench_n8.err; echo "bench n8 rc $?" >> $S
timeout 400 $TR8 --master-port 29915 bench.py --gpus 8 --steps 20 --warmup 5 --hook nccl_bf16 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_bench_n8_nccl.json 2> gpurun_out/r2_7_bench_n8_nccl.err; echo "bench n8 nccl rc $?" >> $S
timeout 400 $TR8 --master-port 29916 bench.py --gpus 8 --steps 20 --warmup 5 --optimizer-in-backward --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_bench_n8_inbw.json 2> gpurun_out/r2_7_bench_n8_inbw.err; echo "bench n8 optimizer-in-backward rc $?" >> $S
timeout 400 $TR4 --master-port 29920 bench.py --gpus 4 --steps 20 --warmup 5 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_bench_n4.json 2> gpurun_out/r2_7_bench_n4.err; echo "bench n4 rc $?" >> $S
timeout 400 $TR8 --master-port 29917 bench.py --gpus 8 --steps 10 --warmup 3 --model bert-base --bucket-cap-mb 1 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_bert_cap1_b200.json 2> gpurun_out/r2_7_bert_cap1_b200.err; echo "bert cap1 b200 rc $?" >> $S
timeout 400 $TR8 --master-port 29918 bench.py --gpus 8 --steps 10 --warmup 3 --model bert-base --bucket-cap-mb 1 --hook nccl_bf16 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_bert_cap1_nccl.json 2> gpurun_out/r2_7_bert_cap1_nccl.err; echo "bert cap1 nccl rc $?" >> $S
timeout 400 $TR8 --master-port 29919 bench.py --gpus 8 --steps 10 --warmup 3 --model gpt2-medium --strategy sharded --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_7_gpt2_sharded.json 2> gpurun_out/r2_7_gpt2_sharded.err; echo "gpt2 sharded rc $?" >> $S
cat $S
