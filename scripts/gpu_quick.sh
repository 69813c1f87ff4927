#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/time_kernel.py ir 2>&1 | tail -1
timeout 120 python scripts/time_kernel.py ir3 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
