#!/bin/bash
# First GPU call of round 2: (1) the 5-D tensor-map -> MN-major UMMA operand probe, (2) the HSB_IR_X5D variant of the
# fused MetaBlock kernel (x tile without a re-stage) against the IR parity tests and the timing harness.
# Build the variants first, here:  python -m hyperseg_b200.build --variant x5d HSB_IR_X5D
#                                  python -m hyperseg_b200.build --variant dw2 HSB_IR_DW2   (adjacent-column depthwise)
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hyperseg_b200/csrc -I include \
     scripts/probe/x5d_probe.cu -o gpurun_out/x5d_probe -lcuda 2>&1 | grep -v warning | tail -3
timeout 120 gpurun_out/x5d_probe 2>&1 | tail -16 | tee gpurun_out/x5d_probe.log
L=$PWD/hyperseg_b200
HSB_LIBRARY=$L/libhsb200.so timeout 300 python scripts/time_kernel.py ir ir3 2>&1 | tail -1
for v in x5d dw2; do
  if [ -f $L/libhsb200_$v.so ]; then
    HSB_LIBRARY=$L/libhsb200_$v.so timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ir or tensor_core" 2>&1 | tail -4
    HSB_LIBRARY=$L/libhsb200_$v.so timeout 300 python scripts/time_kernel.py ir ir3 2>&1 | tail -1
  fi
done
if [ -f $L/libhsb200_x5d.so ]; then
  HSB_LIBRARY=$L/libhsb200_x5d.so HSB_IR_XBLOCKED=1 timeout 300 python scripts/x5d_blocked_check.py 2>&1 | tail -4
fi
# (3) experimental fused depthwise conv of the encoder (engine flag), parity then A/B of the whole step
HSB_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "depthwise" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160
HSB_FUSED_DW=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160
