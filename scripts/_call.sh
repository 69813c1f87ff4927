timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "head or arranged or uint8" 2>&1 | tail -2
timeout 200 python scripts/time_heads.py 2>&1 | tail -9
HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_prof.so timeout 120 python scripts/run_kernel.py head4a 2>&1 | grep hsb-prof | tail -3
for k in conv0 conv2; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:patch_conv1x1 -s 2 -c 1 -o gpurun_out/r02_$k python scripts/run_kernel.py $k 2>&1 | tail -1; done
