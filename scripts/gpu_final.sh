#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, bench (both arms), ncu launch list + full captures.
set -u
mkdir -p gpurun_out
export HSB_VERBOSE=0
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
for k in ir ir3 head4 conv0 epi; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'patch_ir|signal2weights|patch_conv1x1|bias_act' -s 3 -c 1 \
     -f -o gpurun_out/prof_$k python scripts/run_kernel.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
done
python scripts/time_kernel.py 2>&1 | tail -1 | tee gpurun_out/time_kernels.log
