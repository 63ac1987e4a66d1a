mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 300 gpurun_out/b_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err
timeout 400 python bench.py --workload lookahead --weightp 1 > gpurun_out/b_la_w.json 2> gpurun_out/b_la_w.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/la_launches.csv python bench.py --workload lookahead --steps 1 --warmup 3 --quick > gpurun_out/ncu_la.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -f -o gpurun_out/search_full python bench.py --workload lookahead --steps 1 --warmup 3 --quick > gpurun_out/ncu_search.log 2>&1
X264CU_STATS=1 timeout 300 python bench.py --workload lookahead --quick --steps 10 2>&1 | grep -E "x264cu" | cut -c1-300 | head -3
python -c "
import json
for f in ('b_default','b_ref','b_la_w'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
"
