mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_golden.py tests/test_gpu_me.py tests/test_gpu_lookahead.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/frame_bench.py 2>&1 | tail -1 | tee gpurun_out/frame_bench.json
