#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1
echo "bench n2 exit $?"
tail -1 gpurun_out/bench_n2.log | cut -c1-1200
