#!/bin/bash
mkdir -p gpurun_out
export HSB_VERBOSE=0
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1500
