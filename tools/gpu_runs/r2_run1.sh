# round 2, call 1 (1 GPU): parity tests of the new staged exchange, smoke (also under ncu), bench N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,count --format=csv > gpurun_out/r2_1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_allreduce.py -x -q -m gpu --timeout 300 > gpurun_out/r2_1_test_allreduce.log 2>&1; echo "allreduce rc $?" >> gpurun_out/r2_1_summary.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multiprocess.py tests/test_gpu_strategy.py -x -q -m gpu --timeout 600 > gpurun_out/r2_1_test_rest.log 2>&1; echo "rest rc $?" >> gpurun_out/r2_1_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_1_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_1_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_1_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/r2_1_smoke_ncu.log 2>&1; echo "smoke under ncu rc $?" >> gpurun_out/r2_1_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_1_bench_n1.json 2> gpurun_out/r2_1_bench_n1.err; echo "bench rc $?" >> gpurun_out/r2_1_summary.txt
cat gpurun_out/r2_1_summary.txt; tail -3 gpurun_out/r2_1_test_allreduce.log; tail -3 gpurun_out/r2_1_test_rest.log; tail -2 gpurun_out/r2_1_smoke.log; tail -c 600 gpurun_out/r2_1_bench_n1.json
