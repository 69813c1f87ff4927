#!/bin/bash
# ncu evidence: launch list of one bench run + full capture of the dominant kernel. Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/breakdown.py > gpurun_out/breakdown.log 2>&1
echo "breakdown exit $?" >> gpurun_out/breakdown.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_ir -c 2 \
   -o gpurun_out/prof_ir python scripts/run_kernel.py ir > gpurun_out/ncu_ir.log 2>&1
echo "ncu ir exit $?" >> gpurun_out/ncu_ir.log
tail -4 gpurun_out/breakdown.log; tail -2 gpurun_out/ncu_bench.log; tail -2 gpurun_out/ncu_ir.log; ls -la gpurun_out
