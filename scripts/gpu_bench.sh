#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench exit: $?"
