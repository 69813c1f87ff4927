timeout 300 python scripts/time_glue.py 2>&1 | tail -4
