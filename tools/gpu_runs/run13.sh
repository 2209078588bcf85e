mkdir -p gpurun_out
for t in allreduce sharded multiprocess strategy; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r13_test_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/r13_summary.txt; tail -3 gpurun_out/r13_test_$t.log
done
python __graft_entry__.py smoke > gpurun_out/r13_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r13_summary.txt
timeout 600 python bench.py > gpurun_out/r13_bench_1.json 2> gpurun_out/r13_bench_1.err; echo "bench exit $?" >> gpurun_out/r13_summary.txt
cat gpurun_out/r13_summary.txt; tail -c 2500 gpurun_out/r13_bench_1.json
