timeout 600 python -m pytest tests/test_gpu_aq.py tests/test_gpu_c_host.py -x -q 2>&1 | tail -8
