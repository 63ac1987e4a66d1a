timeout 900 python -m pytest tests/test_gpu_mbtree.py -x -q 2>&1 | grep -E "^E|passed|failed" | head -12
