#!/bin/bash
mkdir -p gpurun_out
export HSB_VERBOSE=0
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "se_gate or bias_act or gate or engine or epilogue" 2>&1 | tail -4
HSB_FUSED_SE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_se1.log 2>&1; tail -1 gpurun_out/bench_se1.log | cut -c1-160
HSB_FUSED_SE=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_se0.log 2>&1; tail -1 gpurun_out/bench_se0.log | cut -c1-160
