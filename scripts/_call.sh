timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python scripts/time_heads.py 2>&1 | tail -9
