timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "head or arranged" 2>&1 | tail -4
timeout 200 python scripts/time_heads.py 2>&1 | tail -9
timeout 300 ncu --set full --clock-control none --import-source on -k regex:signal2weights_tc -s 2 -c 1 -o gpurun_out/r02_head4a_v3 python scripts/run_kernel.py head4a 2>&1 | tail -1
