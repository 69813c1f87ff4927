#!/bin/bash
mkdir -p gpurun_out
export HSB_VERBOSE=0
timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-200
