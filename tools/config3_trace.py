#!/usr/bin/env python
"""BASELINE configs[3] (7680x4320, rc-lookahead 250, bframes 16, b-adapt 2) on ONE GPU, chunk by chunk: host wall time, pictures
decided, searches launched and device time of the search launches per chunk of pictures fed.  Shows where the stream's time goes
(first analysis of the 250-picture window vs. steady state) and that the steady-state figure bench.py reports is a steady state.
  python tools/config3_trace.py [pictures] [chunk] [speculate 0|1] [width height] [bframes rc_lookahead]
(a small width x height leaves only the host logic and the launch overheads: what a sharded stream cannot split)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import x264_b200 as x

total = int(sys.argv[1]) if len(sys.argv) > 1 else 588
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 48
if len(sys.argv) > 5:
    bench.C3_W, bench.C3_H = int(sys.argv[4]), int(sys.argv[5])
if len(sys.argv) > 7:
    bench.C3_OPTS = dict(bench.C3_OPTS, bframes=int(sys.argv[6]))
    bench.C3_ST = dict(bench.C3_ST, rc_lookahead=int(sys.argv[7]))
ctx = x.Context(0)
frames = bench.make_la_frames(4320, bench.C3_CLIP, lambda b: np.empty(b, np.uint8), bench.C3_W, bench.C3_H)
d_frames = ctx.malloc(frames.nbytes + 256)
ctx.h2d(d_frames, frames)
st = x.Slicetype(ctx, bench.C3_W, bench.C3_H, **bench.C3_ST, **bench.C3_OPTS, weighted_pred=0)
if len(sys.argv) > 3:
    st.set_speculation(int(sys.argv[3]))          # default: only on a sharded stream
rows = []
t_all = time.perf_counter()
prev = st.search_stats()
types = []
for k0 in range(0, total, chunk):
    t0 = time.perf_counter()
    got = 0
    for i in range(k0, min(k0 + chunk, total)):
        fr, ty = st.step_device(d_frames + (i % bench.C3_CLIP) * bench.C3_W * bench.C3_H, bench.C3_W)
        if fr >= 0:
            got += 1
            types.append(ty)
    wall = time.perf_counter() - t0
    cur = st.search_stats()
    rows.append({"fed": [k0, min(k0 + chunk, total)], "decided": got, "wall_ms": wall * 1e3, "fed_per_s": (min(k0 + chunk, total) - k0) / wall,
                 "search_launches": cur[1] - prev[1], "searches": cur[2] - prev[2], "search_device_ms": cur[0] - prev[0]})
    prev = cur
while len(types) < total:                          # flush: the remaining decisions
    fr, ty = st.step(None)
    if fr < 0:
        break
    types.append(ty)
ctx.sync()
wall_all = time.perf_counter() - t_all
names = {1: "I", 2: "P", 3: "b", 4: "B", 5: "i"}
print(json.dumps({"speculation": st.speculation_stats(), "pictures": total, "wall_s": wall_all, "overall_fed_per_s": total / wall_all, "chunks": rows,
                  "types_tail": "".join(str(t) for t in types[-64:])}))
st.close()
ctx.close()
