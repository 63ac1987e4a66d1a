timeout 900 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_slicetype.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', round(d['value'],1), round(d['e2e']['value'],1))"; done
X264CU_STATS=1 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>&1 | grep -E "x264cu" | cut -c1-300 | tail -3
