mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_2gpu.json 2> gpurun_out/b_2gpu.err; tail -c 600 gpurun_out/b_2gpu.err
python -c "import json; d=json.loads(open('gpurun_out/b_2gpu.json').read().strip().splitlines()[-1]); print('2gpu', d['n_gpus'], d['value'], d['e2e']['value'], d.get('sharded_stream'), d['satd_16x16']['value'])"
