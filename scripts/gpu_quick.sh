#!/bin/bash
for c in 1 2 3; do HSB_VERBOSE=1 HSB_IR_CTAS=$c timeout 120 python scripts/time_kernel.py ir 2>&1 | tail -2; done
for c in 1 2 3 4; do HSB_IR_CTAS=$c timeout 120 python scripts/time_kernel.py ir3 2>&1 | tail -1; done
