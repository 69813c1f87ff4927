#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/tc_debug.py > gpurun_out/sanitizer.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer.log
grep -v "^$" gpurun_out/sanitizer.log | head -60
