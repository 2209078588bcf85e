# Every rank under its OWN ncu (one profiler per process; kernels of different ranks still run concurrently).
#   bash tools/gpu_runs/ncu_multirank.sh <n_gpus> <algo,algo,...> <wire_bytes> <tag> [metrics]
# The staged kernels tolerate ncu's kernel replay (they wait only at their start for monotone flags set by earlier
# launches); the single-kernel algorithms spin on concurrently running peers, so give those a metric set that
# needs ONE pass only.  Everything runs under a short timeout and the library's own 8 s peer watchdog.
N=${1:-2}; ALGO=${2:-staged}; BYTES=${3:-16777216}; TAG=${4:-x}
METRICS=${5:-gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvlrx__bytes.sum,nvltx__bytes.sum}
mkdir -p gpurun_out
B2D_NCU_ALGO=$ALGO B2D_NCU_WIRE_BYTES=$BYTES timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29977 --no-python \
  bash -c 'ncu --metrics '"$METRICS"' --clock-control none -k regex:"exch_kernel|stage_kernel|k1_one|k2_two|wait_published" \
           --csv --log-file gpurun_out/ncu_'"$TAG"'_rank${RANK}.csv python tools/microbench.py ncu_target' > gpurun_out/ncu_multirank_$TAG.log 2>&1
echo "ncu multirank $TAG exit $?"; tail -3 gpurun_out/ncu_multirank_$TAG.log
