echo "== unroll 4 (default)"; timeout 200 python scripts/time_glue.py 2>&1 | tail -1
echo "== unroll 1"; HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_t1.so timeout 200 python scripts/time_glue.py 2>&1 | tail -1
echo "== unroll 8"; HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_t8.so timeout 200 python scripts/time_glue.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "argmax or decoder_input or engine" 2>&1 | tail -2
