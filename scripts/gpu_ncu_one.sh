#!/bin/bash
# usage: gpu_ncu_one.sh <run_kernel tag> <kernel regex>
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_$1 python scripts/run_kernel.py $1 > gpurun_out/ncu_$1.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_$1.log
