timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | grep "smoke" | tail -8
