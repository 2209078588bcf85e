mkdir -p gpurun_out
N=${1:-4}
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + RANDOM % 90)) bench.py --gpus $N --steps 10 --warmup 4 --no-cpu-baseline "$@" > gpurun_out/r11_${tag}_n$N.json 2> gpurun_out/r11_${tag}_n$N.err; echo "$tag exit $?" >> gpurun_out/r11_summary_$N.txt; tail -c 700 gpurun_out/r11_${tag}_n$N.json; echo; }
run resnet50_b200 --model resnet50
run resnet50_ncclbf16 --model resnet50 --hook nccl_bf16
run bert_b200 --model bert-base
run bert_ncclbf16 --model bert-base --hook nccl_bf16
run bert_b200_cap5 --model bert-base --bucket-cap-mb 5
run bert_ncclbf16_cap5 --model bert-base --hook nccl_bf16 --bucket-cap-mb 5
run gpt2_sharded --model gpt2-medium --strategy sharded
run gpt2_ddp_b200 --model gpt2-medium
run gpt2_ddp_ncclbf16 --model gpt2-medium --hook nccl_bf16
cat gpurun_out/r11_summary_$N.txt
