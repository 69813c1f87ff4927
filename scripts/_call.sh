mkdir -p gpurun_out
timeout 300 python scripts/ir2_sizes.py 2>&1 | tail -20
if [ ${PIPESTATUS[0]} -ne 0 ]; then timeout 600 compute-sanitizer --launch-timeout 60 python scripts/ir2_sizes.py 2>&1 | grep -v "^$" | head -60; fi
timeout 300 python scripts/check_ir2.py 2>&1 | grep -v " OK" | tail -12
HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_prof.so timeout 120 python scripts/run_kernel.py ir2 2>&1 | grep hsb-prof | tail -9
