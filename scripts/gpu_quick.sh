#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_debug.py > gpurun_out/train_debug.log 2>&1
timeout 600 python -m pytest tests/test_grads.py -m gpu -q 2>&1 | tail -15
