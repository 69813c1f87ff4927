#!/bin/bash
# Round-2 bench collection on one B200: every BASELINE configuration that fits one GPU, JSON lines into gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 420 python bench.py > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err
tail -c 600 gpurun_out/r02_bench_m.err
timeout 300 python bench.py --config s-city --no-cpu-baseline > gpurun_out/r02_bench_s_city.json 2> gpurun_out/r02_bench_s_city.err
timeout 400 python bench.py --config s-camvid --sweep --no-cpu-baseline > gpurun_out/r02_bench_s_camvid.json 2> gpurun_out/r02_bench_s_camvid.err
timeout 300 python bench.py --config l-voc-train --no-cpu-baseline > gpurun_out/r02_bench_l_voc_train.json 2> gpurun_out/r02_bench_l_voc_train.err
for f in gpurun_out/r02_bench_*.json; do echo "== $f"; head -c 1500 $f; echo; done
tail -5 gpurun_out/r02_bench_l_voc_train.err
