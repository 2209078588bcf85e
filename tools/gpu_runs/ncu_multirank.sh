# UNTESTED (written after the round-1 GPU budget was spent) — the plan of DESIGN.md §9.6:
# every rank under its own ncu, single-pass metrics only (no kernel replay: the kernels wait for peers),
# everything under a short timeout so a stuck capture cannot hold the box.
#   bash tools/gpu_runs/ncu_multirank.sh 2 two_shot 16777216
N=${1:-2}; ALGO=${2:-two_shot}; BYTES=${3:-16777216}
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvlrx__bytes.sum,nvltx__bytes.sum,lts__t_bytes.sum
B2D_NCU_ALGO=$ALGO B2D_NCU_WIRE_BYTES=$BYTES timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29977 --no-python \
  bash -c 'ncu --metrics '"$METRICS"' --clock-control none --replay-mode application --target-processes all \
           -k regex:"k[0-9]" --csv --log-file gpurun_out/ncu_rank${RANK}_'"$ALGO"'.csv \
           python tools/microbench.py ncu_target' > gpurun_out/ncu_multirank_$ALGO.log 2>&1
echo "exit $?"; tail -5 gpurun_out/ncu_multirank_$ALGO.log
