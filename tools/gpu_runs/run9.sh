mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_allreduce.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r9_test_allreduce.log 2>&1
echo "test_allreduce exit $?" >> gpurun_out/r9_summary.txt
tail -15 gpurun_out/r9_test_allreduce.log
timeout 900 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q --timeout 300 -p no:cacheprovider -k "ddp_with" > gpurun_out/r9_test_mp.log 2>&1
echo "test_mp_ddp exit $?" >> gpurun_out/r9_summary.txt
tail -15 gpurun_out/r9_test_mp.log
cat gpurun_out/r9_summary.txt
