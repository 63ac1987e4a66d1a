timeout 1200 python -m pytest tests/test_gpu_configs.py -x -q --durations=10 2>&1 | tail -25
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
