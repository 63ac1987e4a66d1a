timeout 300 python -m pytest tests/test_gpu_c_host.py -x -q 2>&1 | tail -2
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_aq.py tests/test_gpu_mbtree.py tests/test_gpu_mc.py tests/test_gpu_frame.py tests/test_golden.py -m gpu -x -q -k "not 3840 and not 1920 and not 1918 and not 352" 2>&1 | tail -4
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_frame.py tests/test_gpu_lookahead.py -x -q -k "hpel and not 200 or cfg0" 2>&1 | tail -3
