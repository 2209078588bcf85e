mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_strategy.py -m gpu -q --timeout 140 -p no:cacheprovider > gpurun_out/r16_test_strategy_2gpu.log 2>&1; echo "exit $?" > gpurun_out/r16_summary.txt
cat gpurun_out/r16_summary.txt; tail -15 gpurun_out/r16_test_strategy_2gpu.log
