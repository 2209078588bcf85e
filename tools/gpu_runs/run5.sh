mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max > gpurun_out/r5_cpu.txt 2>&1; nproc >> gpurun_out/r5_cpu.txt; lscpu | head -20 >> gpurun_out/r5_cpu.txt
for t in 16 32 64; do
  /usr/bin/time -v -o gpurun_out/r5_cpu_time_$t.txt env B2D_CPU_THREADS=$t timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --cpu-batch 8 2>/dev/null | tail -1 > gpurun_out/r5_ref_t$t.json
  echo "threads $t: $(python -c "import json;d=json.load(open('gpurun_out/r5_ref_t$t.json'));print(d['value'], d['ms_per_step'])")" >> gpurun_out/r5_cpu.txt
done
B2D_CPU_THREADS=32 timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --cpu-batch 8 2>/dev/null | tail -1 > gpurun_out/r5_ref_w2.json
cat gpurun_out/r5_cpu.txt
# launch list of one bench run (short): every kernel with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_launch.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/r5_summary.txt
# the K0 kernel, full set, 3 launches of the steady state
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k0_cast -s 6 -c 3 -o gpurun_out/r5_prof_k0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_k0.log 2>&1
echo "ncu k0 exit $?" >> gpurun_out/r5_summary.txt
cat gpurun_out/r5_summary.txt; ls -la gpurun_out | tail -8
