mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 300 gpurun_out/b_default.err
python -c "import json; d=json.load(open('gpurun_out/b_default.json')); print('default', d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['cpu_baseline'])"
timeout 400 python bench.py --workload lookahead --weightp 1 > gpurun_out/b_la_w.json 2> gpurun_out/b_la_w.err
python -c "import json; d=json.load(open('gpurun_out/b_la_w.json')); print('weightp', d['value'], d['e2e']['value'], d['cpu_baseline'])"
for cfg in "24 12" "32 16" "32 12"; do set -- $cfg
 X264CU_RUN_AHEAD=$1 X264CU_PREFETCH_GROUP=$2 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ra/pg $cfg', round(d['value'],1), round(d['e2e']['value'],1))"
done
X264CU_STATS=1 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>&1 | grep -E "x264cu" | cut -c1-300 | head -3
