mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 300 gpurun_out/b_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err
timeout 400 python bench.py --workload lookahead --weightp 1 > gpurun_out/b_la_w.json 2> gpurun_out/b_la_w.err
python -c "
import json
for f in ('b_default','b_ref','b_la_w'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), (d.get('cpu_baseline') or {}).get('value'), d.get('roofline',{}).get('share_of_step'), d.get('gpu_launches'))
"
