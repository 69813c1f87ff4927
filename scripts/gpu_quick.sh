#!/bin/bash
mkdir -p gpurun_out
HSB_VERBOSE=1 timeout 120 python scripts/run_kernel.py ir 2>&1 | tail -3
HSB_VERBOSE=1 timeout 120 python scripts/run_kernel.py ir3 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench exit: $?"
