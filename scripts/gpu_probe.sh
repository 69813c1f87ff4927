#!/bin/bash
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7 8 9 10; do timeout 60 ./scripts/probe/tma_probe $i 2>&1 | grep -v "encode fn" | grep -v sticky; done | tee gpurun_out/tma_probe.log
