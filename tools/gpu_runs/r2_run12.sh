# round 2, call 12 (1 GPU, short): sanity of the per-slot event change — allreduce parity tests + smoke
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_allreduce.py -q -m gpu --timeout 120 > gpurun_out/r2_12_tests.log 2>&1; echo "tests rc $?"
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2_12_smoke.log 2>&1; echo "smoke rc $?"
tail -2 gpurun_out/r2_12_tests.log; tail -1 gpurun_out/r2_12_smoke.log
