mkdir -p gpurun_out
N=${1:-1}
timeout 900 python -m pytest tests/test_gpu_strategy.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r4_test_strategy_$N.log 2>&1
echo "test_strategy exit $?" >> gpurun_out/r4_summary_$N.txt
tail -15 gpurun_out/r4_test_strategy_$N.log
python __graft_entry__.py smoke > gpurun_out/r4_smoke_$N.log 2>&1; echo "smoke exit $?" >> gpurun_out/r4_summary_$N.txt
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r4_bench_1.json 2> gpurun_out/r4_bench_1.err
  echo "bench1 exit $?" >> gpurun_out/r4_summary_$N.txt
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r4_bench_$N.json 2> gpurun_out/r4_bench_$N.err
  echo "bench$N exit $?" >> gpurun_out/r4_summary_$N.txt
fi
cat gpurun_out/r4_summary_$N.txt; tail -2 gpurun_out/r4_bench_$N.json; tail -5 gpurun_out/r4_bench_$N.err
