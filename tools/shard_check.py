#!/usr/bin/env python
"""ONE 4K picture stream sharded over N GPUs (x264cu_slicetype_set_shard + x264_b200.dist.ShardExchange over NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_check.py
Every rank is fed the same pictures; checks that every rank's frame types and MB-tree offsets equal those of an unsharded run on
the same GPU, then times the sharded stream (CUDA events on rank 0's context stream, max over ranks) next to the unsharded one.
Prints one JSON line on rank 0.  Not part of the product path."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import x264_b200 as x
from x264_b200 import dist as xd

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = x.Context(local)
n = 96
frames = bench.make_la_frames(2160, n, lambda b: ctx.malloc_host(b))          # the SAME stream on every rank
d_frames = ctx.malloc(frames.nbytes + 256)
ctx.h2d(d_frames, frames)
W, H = bench.LA_W, bench.LA_H


def run(shard, steps, collect):
    st = x.Slicetype(ctx, W, H, **bench.LA_ST, **bench.LA_OPTS)
    ex = xd.ShardExchange(dist, device=dev) if shard else None
    if shard:
        st.set_shard(rank, world, ex)
    types, qp = [], {}
    def one_pass():
        for i in range(n):
            fr, ty = st.step_device(d_frames + i * W * H, W)
            if fr >= 0 and collect:
                types.append((fr, ty))
                if ty not in (4, 5):
                    qp[fr] = st.get_qp_offset(fr)
    one_pass()                       # warm-up (also fills the pipeline)
    ctx.sync()
    dist.barrier(device_ids=[local])
    types.clear(); qp.clear()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    ms = ctx.timer_stop()
    ctx.sync()
    wall = time.perf_counter() - t0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    stats = (ex.calls, ex.bytes) if ex else (0, 0)
    st.close()
    return types, qp, float(t.item()), wall, stats

steps = 3
t_single, qp_single, ms_single, _, _ = run(False, steps, True)
t_shard, qp_shard, ms_shard, _, (calls, nbytes) = run(True, steps, True)
same = t_single == t_shard and all(np.array_equal(qp_single[k], qp_shard[k]) for k in qp_single) and len(qp_single) > 10
flags = [None] * world
dist.all_gather_object(flags, bool(same))
if rank == 0:
    print(json.dumps({"n_gpus": world, "pictures_per_pass": n, "steps": steps,
                      "sharded_equals_single_gpu_on_every_rank": all(flags), "decisions_compared": len(t_single), "qp_offset_arrays_compared": len(qp_single),
                      "single_gpu_frames_per_s": n * steps / (ms_single * 1e-3), "sharded_stream_frames_per_s": n * steps / (ms_shard * 1e-3),
                      "exchanges": calls, "bytes_per_exchange_total": nbytes / max(calls, 1)}))
dist.barrier(device_ids=[local])
ctx.close()
dist.destroy_process_group()
