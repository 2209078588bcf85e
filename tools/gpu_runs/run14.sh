mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29933 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r14_bench_8.json 2> gpurun_out/r14_bench_8.err; echo "bench8 exit $?" >> gpurun_out/r14_summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29934 bench.py --gpus $N --steps 20 --warmup 5 --algo nvls > gpurun_out/r14_bench_8_nvls.json 2> gpurun_out/r14_bench_8_nvls.err; echo "bench8 nvls exit $?" >> gpurun_out/r14_summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29935 bench.py --gpus $N --steps 20 --warmup 5 --hook nccl_bf16 > gpurun_out/r14_bench_8_nccl.json 2> gpurun_out/r14_bench_8_nccl.err; echo "bench8 nccl exit $?" >> gpurun_out/r14_summary.txt
cat gpurun_out/r14_summary.txt; for f in gpurun_out/r14_bench_8.json gpurun_out/r14_bench_8_nvls.json gpurun_out/r14_bench_8_nccl.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], (d.get("allreduce_isolated") or {}).get("GBps_total"))
PY
done
