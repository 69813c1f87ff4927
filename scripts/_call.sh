#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, smoke, ncu launch list of an ungraphed bench step, one `ncu --set full`
# capture per decoder kernel at its HyperSeg-M batch-8 shape (summarised by scripts/summarize_ncu.py), kernel timings.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | grep "smoke" | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_nograph.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-200
for k in ir2 ir2_3 head4a head3a head0 conv0 conv1 conv2; do
  case $k in ir2*) pat=patch_ir2;; head*) pat=signal2weights_tc;; conv0) pat=conv1x1;; *) pat=patch_conv1x1;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o gpurun_out/prof_$k -f python scripts/run_kernel.py $k 2>&1 | tail -1
done
timeout 200 python scripts/time_heads.py 2>&1 | tail -9 > gpurun_out/r02_kernel_times.txt
timeout 200 python scripts/time_conv.py 2>&1 | tail -4 >> gpurun_out/r02_kernel_times.txt
timeout 300 python scripts/check_ir2.py --time-only 2>&1 | grep "new\|old" >> gpurun_out/r02_kernel_times.txt
cat gpurun_out/r02_kernel_times.txt
