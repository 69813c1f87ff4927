for i in 1 2; do
timeout 300 python scripts/check_ir2.py --time-only 2>&1 | grep "new"
HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_allpoll.so timeout 300 python scripts/check_ir2.py --time-only 2>&1 | grep "new" | sed 's/^/allpoll /'
done
