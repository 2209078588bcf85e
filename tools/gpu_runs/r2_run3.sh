# round 2, call 3 (2 GPUs): full gpu test suite on the new code, W=2 sweep/tune after the CTA-footprint change, bench N=2 (bf16 + fp32 arena buckets), sharded bench, per-rank ncu retry
mkdir -p gpurun_out
S=gpurun_out/r2_3_summary.txt; : > $S
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/r2_3_tests.log 2>&1; echo "gpu tests rc $?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29911 tools/microbench.py sweep > gpurun_out/r2_3_sweep.jsonl 2> gpurun_out/r2_3_sweep.err; echo "sweep rc $?" >> $S
timeout 600 $TR --master-port 29912 tools/microbench.py tune > gpurun_out/r2_3_tune.jsonl 2> gpurun_out/r2_3_tune.err; echo "tune rc $?" >> $S
timeout 900 $TR --master-port 29913 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_3_bench_n2.json 2> gpurun_out/r2_3_bench_n2.err; echo "bench n2 rc $?" >> $S
timeout 900 $TR --master-port 29914 bench.py --gpus 2 --steps 20 --warmup 5 --wire fp32 --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_3_bench_n2_fp32.json 2> gpurun_out/r2_3_bench_n2_fp32.err; echo "bench n2 fp32 (arena buckets) rc $?" >> $S
timeout 900 $TR --master-port 29915 bench.py --gpus 2 --steps 10 --warmup 3 --model gpt2-medium --strategy sharded --no-parity --no-sweep --no-cpu-baseline > gpurun_out/r2_3_bench_n2_gpt2_sharded.json 2> gpurun_out/r2_3_bench_n2_gpt2_sharded.err; echo "bench n2 gpt2 sharded rc $?" >> $S
# per-rank ncu, separate lock/tmp dirs, exchange kernel only, few metrics
N=2
B2D_NCU_ALGO=staged,nvls B2D_NCU_WIRE_BYTES=16777216 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29977 --no-python \
  bash -c 'mkdir -p /tmp/ncu_r$RANK; TMPDIR=/tmp/ncu_r$RANK ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum --clock-control none -k regex:"exch_kernel" -c 6 --csv --log-file gpurun_out/ncu_r2_3_exch_rank${RANK}.csv python tools/microbench.py ncu_target' > gpurun_out/ncu_multirank_r2_3.log 2>&1
echo "ncu multirank retry exit $?" >> $S
cat $S; tail -4 gpurun_out/r2_3_tests.log
