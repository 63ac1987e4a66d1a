#!/usr/bin/env python
"""Warm CUDA-event timing of the two frame-preparation kernels at 3840x2160 (streaming, HBM-bound):
x264cu_frame_init_lowres (2*W*H algorithmic bytes) and x264cu_hpel_filter (4*W*H).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import x264_b200 as x
from _libs import PaddedPlane, PAD

ctx = x.Context(0)
w, h = 3840, 2160
rng = np.random.default_rng(1)
luma = rng.integers(0, 256, (h, w), dtype=np.uint8)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
res = {"picture": "%dx%d" % (w, h), "hbm_peak_gbs": peak}
reps = 50
# lowres
wl, ll = w // 2, h // 2
pl = PaddedPlane(wl, ll)
d_src = ctx.upload(luma)
d_planes = ctx.malloc(4 * pl.buf.size + 256)
darr = (C.c_void_p * 4)(*[d_planes + i * pl.buf.size + pl.origin for i in range(4)])
f = lambda: ctx.check(ctx.L.x264cu_frame_init_lowres(ctx.h, d_src, w, w, h, darr, pl.stride))
for _ in range(5):
    f()
ctx.sync(); ctx.timer_start()
for _ in range(reps):
    f()
ms = ctx.timer_stop() / reps
res["frame_init_lowres"] = {"us": ms * 1e3, "algorithmic_bytes": 2 * w * h, "gbs": 2 * w * h / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 2 * w * h / (ms * 1e-3) / 1e9 / peak}
# hpel
src = PaddedPlane(w, h)
src.inner()[:] = luma
d = [ctx.upload(src.buf)] + [ctx.malloc(src.buf.size + 256) for _ in range(3)]
org = src.origin
g = lambda: ctx.check(ctx.L.x264cu_hpel_filter(ctx.h, d[0] + org, src.stride, w, h, d[1] + org, d[2] + org, d[3] + org, 1))
for _ in range(5):
    g()
ctx.sync(); ctx.timer_start()
for _ in range(reps):
    g()
ms = ctx.timer_stop() / reps
res["hpel_filter"] = {"us": ms * 1e3, "algorithmic_bytes": 4 * w * h, "gbs": 4 * w * h / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 4 * w * h / (ms * 1e-3) / 1e9 / peak}
print(json.dumps(res))
ctx.close()
