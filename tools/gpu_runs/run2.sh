mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
for mem in vmm ipc; do
  B2D_MEM=$mem timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/microbench.py sweep > gpurun_out/r2_sweep_${N}_$mem.jsonl 2> gpurun_out/r2_sweep_${N}_$mem.err
  echo "sweep $N $mem exit $?" >> gpurun_out/r2_summary_$N.txt
done
B2D_MEM=vmm B2D_MAX_CTAS=16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/microbench.py sweep > gpurun_out/r2_sweep_${N}_vmm_cta16.jsonl 2> gpurun_out/r2_sweep_${N}_vmm_cta16.err
B2D_MEM=vmm B2D_MAX_CTAS=128 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/microbench.py sweep > gpurun_out/r2_sweep_${N}_vmm_cta128.jsonl 2> gpurun_out/r2_sweep_${N}_vmm_cta128.err
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/r2_test_mp_$N.log 2>&1
echo "test_mp exit $?" >> gpurun_out/r2_summary_$N.txt
cat gpurun_out/r2_summary_$N.txt; cat gpurun_out/r2_sweep_${N}_vmm.jsonl; tail -3 gpurun_out/r2_sweep_${N}_vmm.err
