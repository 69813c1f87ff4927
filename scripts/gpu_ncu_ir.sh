#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_ir_tc -s 2 -c 1 \
   -f -o gpurun_out/prof_ir_tc python scripts/run_kernel.py ir > gpurun_out/ncu_ir_tc.log 2>&1
echo "exit $?" >> gpurun_out/ncu_ir_tc.log
tail -3 gpurun_out/ncu_ir_tc.log
