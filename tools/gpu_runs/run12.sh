mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "1" ]; then
  for t in allreduce sharded multiprocess; do
    timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r12_test_$t.log 2>&1
    echo "test_$t exit $?" >> gpurun_out/r12_summary_1.txt; tail -4 gpurun_out/r12_test_$t.log
  done
  cat gpurun_out/r12_summary_1.txt
else
  B2D_TRACE=1 B2D_SKIP_FP32=1 B2D_ITERS=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29911 tools/microbench.py sweep > gpurun_out/r12_sweep_${N}.jsonl 2> gpurun_out/r12_sweep_${N}.err
  echo "sweep exit $?" >> gpurun_out/r12_summary_$N.txt
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29912 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r12_bench_$N.json 2> gpurun_out/r12_bench_$N.err
  echo "bench exit $?" >> gpurun_out/r12_summary_$N.txt
  cat gpurun_out/r12_summary_$N.txt; grep -h '"sweep"' gpurun_out/r12_sweep_${N}.jsonl | cut -c1-700; grep -h trace gpurun_out/r12_sweep_${N}.jsonl | grep -E '"two_shot"|nvls' | grep -E "16777216|67108864|268435456" | cut -c1-300; tail -c 1500 gpurun_out/r12_bench_$N.json
fi
