#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the dominant kernels.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/ncu_bench.log
for k in ir ir3 head4 head0 conv0 conv2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'patch_ir|signal2weights|patch_conv1x1' -s 3 -c 1 \
     -f -o gpurun_out/prof_$k python scripts/run_kernel.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
done
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_full.log 2>&1
echo "bench exit $?"
tail -c 600 gpurun_out/bench_full.log
ls -la gpurun_out | head -30
