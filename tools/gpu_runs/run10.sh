mkdir -p gpurun_out
N=${1:-8}
B2D_TRACE=1 B2D_SKIP_FP32=1 B2D_SKIP_NVLS=1 B2D_ITERS=20 B2D_TMA_CTAS=32 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 tools/microbench.py sweep > gpurun_out/r10_sweep_${N}_tma32.jsonl 2> gpurun_out/r10_sweep_${N}_tma32.err
echo "sweep tma32 exit $?" >> gpurun_out/r10_summary_$N.txt
for c in 16 64 128; do
B2D_SKIP_NCCL=1 B2D_SKIP_FP32=1 B2D_SKIP_NVLS=1 B2D_ITERS=20 B2D_TMA_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2982$((c%10)) tools/microbench.py sweep > gpurun_out/r10_sweep_${N}_tma$c.jsonl 2> gpurun_out/r10_sweep_${N}_tma$c.err
echo "sweep tma$c exit $?" >> gpurun_out/r10_summary_$N.txt
done
cat gpurun_out/r10_summary_$N.txt; grep -h '"sweep"' gpurun_out/r10_sweep_${N}_tma32.jsonl | cut -c1-600; tail -3 gpurun_out/r10_sweep_${N}_tma32.err
