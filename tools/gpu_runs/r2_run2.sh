# round 2, call 2 (2 GPUs): NVLS + cross-GPU parity tests, sweep/tune at W=2, bench N=2 with the parity block, per-rank ncu
mkdir -p gpurun_out
S=gpurun_out/r2_2_summary.txt; : > $S
nvidia-smi --query-gpu=index,name --format=csv >> $S 2>&1
nvidia-smi topo -m > gpurun_out/r2_2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_allreduce.py -x -q -m gpu --timeout 300 -k "nvls or staged_p2p or in_place or auto" > gpurun_out/r2_2_test_nvls.log 2>&1; echo "nvls tests rc $?" >> $S
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -x -q -m gpu --timeout 300 > gpurun_out/r2_2_test_mp.log 2>&1; echo "multiprocess tests rc $?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29911 tools/microbench.py sweep > gpurun_out/r2_2_sweep.jsonl 2> gpurun_out/r2_2_sweep.err; echo "sweep rc $?" >> $S
timeout 600 $TR --master-port 29912 tools/microbench.py tune > gpurun_out/r2_2_tune.jsonl 2> gpurun_out/r2_2_tune.err; echo "tune rc $?" >> $S
timeout 900 $TR --master-port 29913 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_2_bench_n2.json 2> gpurun_out/r2_2_bench_n2.err; echo "bench n2 rc $?" >> $S
timeout 900 $TR --master-port 29914 bench.py --gpus 2 --steps 20 --warmup 5 --wire fp32 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_2_bench_n2_fp32.json 2> gpurun_out/r2_2_bench_n2_fp32.err; echo "bench n2 fp32 (arena buckets) rc $?" >> $S
ncu --query-metrics 2>/dev/null | grep -i -E "nvl|pcie" | head -40 > gpurun_out/r2_2_nvl_metrics.txt
bash tools/gpu_runs/ncu_multirank.sh 2 staged,nvls 16777216 r2_2_staged >> $S 2>&1
bash tools/gpu_runs/ncu_multirank.sh 2 one_shot,two_shot 16777216 r2_2_fused gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum >> $S 2>&1
timeout 600 python -m pytest tests/test_gpu_strategy.py -x -q -m gpu --timeout 300 > gpurun_out/r2_2_test_strategy.log 2>&1; echo "strategy tests rc $?" >> $S
cat $S; tail -3 gpurun_out/r2_2_test_nvls.log; tail -3 gpurun_out/r2_2_test_mp.log; tail -3 gpurun_out/r2_2_test_strategy.log; tail -c 400 gpurun_out/r2_2_bench_n2.err
