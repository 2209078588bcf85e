mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r1_gpu.txt 2>&1
for t in allreduce sharded multiprocess; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/r1_test_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/r1_summary.txt
  tail -5 gpurun_out/r1_test_$t.log
done
timeout 300 python tools/microbench.py k0 > gpurun_out/r1_k0.jsonl 2> gpurun_out/r1_k0.err; echo "k0 exit $?" >> gpurun_out/r1_summary.txt
timeout 300 python tools/microbench.py loopback > gpurun_out/r1_loopback.jsonl 2> gpurun_out/r1_loopback.err; echo "loopback exit $?" >> gpurun_out/r1_summary.txt
cat gpurun_out/r1_summary.txt; cat gpurun_out/r1_k0.jsonl | tail -12
