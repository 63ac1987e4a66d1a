#!/usr/bin/env python
"""One search launch of 7 searches (one 4K picture against its 4 predecessors / 3 successors) -- the light-load shape --
for `ncu -k regex:search_kernel -c 1 python tools/la_one_launch.py [n_pictures]`.  Not part of the product path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import x264_b200 as x

ctx = x.Context(0)
frames = bench.make_la_frames(2160, 8, lambda b: ctx.malloc_host(b))
d = ctx.malloc(frames.nbytes + 256)
ctx.h2d(d, frames)
la = x.Lookahead(ctx, bench.LA_W, bench.LA_H, n_slots=8, **bench.LA_OPTS)
for i in range(8):
    la.frame_put_device(i, d + i * bench.LA_W * bench.LA_H, bench.LA_W)
jobs = [(4, 4 - k, 0, k) for k in range(1, 5)] + [(4 - k, 4, 1, k) for k in range(1, 4)]
la.search_batch(jobs)
la.join()
ctx.sync()
print("done")
