mkdir -p gpurun_out
# N=1 validation of every bench.py mode (short)
for cfg in "--model resnet50" "--model bert-base --batch 8" "--model bert-base --batch 8 --hook nccl_bf16" "--model gpt2-medium --strategy sharded --batch 2 --seq 512" "--model gpt2-medium --batch 2 --seq 512"; do
  tag=$(echo $cfg | tr -d ' -' | tr '/' '_')
  timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/r8_$tag.json 2> gpurun_out/r8_$tag.err
  echo "$cfg exit $?" >> gpurun_out/r8_summary.txt
  tail -c 1500 gpurun_out/r8_$tag.json; echo; tail -2 gpurun_out/r8_$tag.err
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 > gpurun_out/r8_ref1.json 2>/dev/null; tail -c 600 gpurun_out/r8_ref1.json
cat gpurun_out/r8_summary.txt
