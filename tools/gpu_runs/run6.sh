mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m | head -12 > gpurun_out/r6_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r6_test_mp_$N.log 2>&1
echo "test_mp exit $?" >> gpurun_out/r6_summary_$N.txt
for c in 32 64; do
B2D_MEM=vmm B2D_MAX_CTAS=$c timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$((c%10)) tools/microbench.py sweep > gpurun_out/r6_sweep_${N}_cta$c.jsonl 2> gpurun_out/r6_sweep_${N}_cta$c.err
echo "sweep cta$c exit $?" >> gpurun_out/r6_summary_$N.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r6_bench_$N.json 2> gpurun_out/r6_bench_$N.err
echo "bench$N exit $?" >> gpurun_out/r6_summary_$N.txt
cat gpurun_out/r6_summary_$N.txt; tail -1 gpurun_out/r6_bench_$N.json; tail -3 gpurun_out/r6_bench_$N.err; tail -3 gpurun_out/r6_test_mp_$N.log
