mkdir -p gpurun_out
N=${1:-2}
for t in allreduce sharded; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/r3_test_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/r3_summary.txt
  tail -3 gpurun_out/r3_test_$t.log
done
for c in 32 64 128; do
B2D_MEM=vmm B2D_MAX_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((c%10)) tools/microbench.py sweep > gpurun_out/r3_sweep_${N}_cta$c.jsonl 2> gpurun_out/r3_sweep_${N}_cta$c.err
echo "sweep cta$c exit $?" >> gpurun_out/r3_summary.txt
done
cat gpurun_out/r3_summary.txt
