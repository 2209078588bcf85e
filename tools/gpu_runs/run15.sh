mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_allreduce.py tests/test_gpu_sharded.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r15_tests.log 2>&1; echo "tests exit $?" > gpurun_out/r15_summary.txt
python __graft_entry__.py smoke > gpurun_out/r15_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r15_summary.txt
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "bench exit $?" >> gpurun_out/r15_summary.txt
cat gpurun_out/r15_summary.txt; tail -2 gpurun_out/r15_tests.log; tail -c 600 gpurun_out/r15_bench.json
