for v in wt2r8 wt4r4 wt4r8; do echo "== $v"; HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_$v.so timeout 200 python scripts/check_ir2.py 2>&1 | grep "ir new\|worst"; done
