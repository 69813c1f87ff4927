#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1
echo "tc_debug exit $?" >> gpurun_out/tc_debug.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
cat gpurun_out/tc_debug.log; tail -15 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench.log
