python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_slicetype.py -x -q 2>&1 | tail -2
X264CU_STATS=1 python bench.py --workload lookahead --quick --steps 10 2>&1 | grep -E "x264cu slice|busy|metric" | cut -c1-200
for cfg in "8 4" "12 4" "16 8" "8 2"; do
  set -- $cfg
  X264CU_RUN_AHEAD=$1 X264CU_PREFETCH_GROUP=$2 python bench.py --workload lookahead --quick --steps 8 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['e2e']['value'])"
done
