#!/usr/bin/env python
"""Phase timing of the lookahead search kernel (clock64 inside the kernel).  Needs a library built with LA_PROFILE=1:
    LA_PROFILE=1 python x264_b200/build.py && python tools/la_phase_profile.py
Prints average cycles per macroblock spent polling the row below, setting up, in the predictor stage, the full-pel
search, the sub-pel refinement.  Not part of the product path."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import x264_b200 as x

ctx = x.Context(0)
frames = bench.make_la_frames(2160, 8, lambda b: ctx.malloc_host(b))
d = ctx.malloc(frames.nbytes + 256)
ctx.h2d(d, frames)
la = x.Lookahead(ctx, bench.LA_W, bench.LA_H, n_slots=8, **bench.LA_OPTS)
for i in range(5):
    la.frame_put_device(i, d + i * bench.LA_W * bench.LA_H, bench.LA_W)
jobs = [(4, 4 - k, 0, k) for k in range(1, 5)] + [(4 - k, 4, 1, k) for k in range(1, 4)]
out = (C.c_ulonglong * 8)()
ctx.L.x264cu_debug_la_profile(out, 1)
la.search_batch(jobs)
la.get_intra(0)
ctx.L.x264cu_debug_la_profile(out, 1)
v = np.array(list(out), dtype=np.float64)
n = v[6]
names = ["poll row below", "setup+mvp+skip test", "predictors", "full-pel search", "sub-pel refine", "window slide"]
print("macroblocks %d, searched %d" % (n, v[7]))
for i, nm in enumerate(names):
    print("%-22s %8.0f cycles / MB" % (nm, v[i] / n))
print("%-22s %8.0f cycles / MB" % ("total (excl. store)", v[:6].sum() / n))
