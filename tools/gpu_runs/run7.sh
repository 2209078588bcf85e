mkdir -p gpurun_out
N=${1:-8}
B2D_TRACE=1 B2D_SKIP_FP32=1 B2D_ITERS=20 B2D_MEM=vmm B2D_MAX_CTAS=64 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/microbench.py sweep > gpurun_out/r7_sweep_${N}.jsonl 2> gpurun_out/r7_sweep_${N}.err
echo "sweep exit $?" >> gpurun_out/r7_summary_$N.txt
grep trace gpurun_out/r7_sweep_${N}.jsonl; grep parity gpurun_out/r7_sweep_${N}.jsonl; tail -3 gpurun_out/r7_sweep_${N}.err
