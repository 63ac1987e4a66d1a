mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_check.py 2> gpurun_out/shard.err | tail -1 > gpurun_out/shard_2gpu.json; cat gpurun_out/shard_2gpu.json; tail -5 gpurun_out/shard.err | cut -c1-300
