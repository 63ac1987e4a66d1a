mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -c 300 gpurun_out/b_default.err
python -c "
import json
d=json.load(open('gpurun_out/b_default.json')); print('default', round(d['value'],1), round(d['e2e']['value'],1), d['cpu_baseline']['value'], d['roofline']['share_of_step'], d['satd_16x16']['roofline']['frac'])
"
