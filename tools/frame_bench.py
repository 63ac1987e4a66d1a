#!/usr/bin/env python
"""CUDA-event timing of the two frame-preparation kernels at 3840x2160 (streaming, HBM-bound): x264cu_frame_init_lowres
(2*W*H algorithmic bytes per picture) and x264cu_hpel_filter (4*W*H).
  cold : a stack of pictures per launch, inputs + outputs larger than the 126 MB L2 (the number that counts)
  warm : one picture per launch, the same picture again and again (what profiles/r01b_frame_kernels.json held)
Prints one JSON line."""
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
L = ctx.L
vp, ss, ci = C.c_void_p, C.c_ssize_t, C.c_int
L.x264cu_frame_init_lowres_batch.argtypes = [vp, vp, ss, ss, ci, ci, ci, C.POINTER(vp), ss, ss]
L.x264cu_hpel_filter_batch.argtypes = [vp, vp, ss, ss, ci, ci, ci, vp, vp, vp, ci]
w, h = 3840, 2160
rng = np.random.default_rng(1)
luma = rng.integers(0, 256, (h, w), dtype=np.uint8)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
res = {"picture": "%dx%d" % (w, h), "hbm_peak_gbs": peak}
reps = 20


def timed(fn, n=reps):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(n):
        fn()
    return ctx.timer_stop() / n


def entry(ms_per_picture, alg, extra):
    gbs = alg / (ms_per_picture * 1e-3) / 1e9
    return dict({"us_per_picture": ms_per_picture * 1e3, "algorithmic_bytes": alg, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}, **extra)


# ---- lowres -------------------------------------------------------------------------------------------------
NP = 16
wl, ll = w // 2, h // 2
pl = PaddedPlane(wl, ll)
pitch_src = w * h
d_src = ctx.malloc(NP * pitch_src + 256)
for i in range(NP):
    ctx.h2d(d_src + i * pitch_src, np.roll(luma, 17 * i, axis=1))
plane_bytes = (pl.buf.size + 255) & ~255
pitch_dst = 4 * plane_bytes
d_planes = ctx.malloc(NP * pitch_dst + 256)
darr = (vp * 4)(*[d_planes + i * plane_bytes + pl.origin for i in range(4)])
cold = timed(lambda: ctx.check(L.x264cu_frame_init_lowres_batch(ctx.h, d_src, w, pitch_src, NP, w, h, darr, pl.stride, pitch_dst))) / NP
warm = timed(lambda: ctx.check(L.x264cu_frame_init_lowres(ctx.h, d_src, w, w, h, darr, pl.stride)), 50)
res["frame_init_lowres"] = {"cold": entry(cold, 2 * w * h, {"pictures_per_launch": NP, "MB_touched_per_launch": NP * (pitch_src + pitch_dst) / 1e6}),
                            "warm": entry(warm, 2 * w * h, {"pictures_per_launch": 1})}
ctx.free(d_src)
ctx.free(d_planes)
# ---- hpel ---------------------------------------------------------------------------------------------------
NH = 8
src = PaddedPlane(w, h)
src.inner()[:] = luma
pitch = (src.buf.size + 255) & ~255
d = [ctx.malloc(NH * pitch + 256) for _ in range(4)]
for i in range(NH):
    src.inner()[:] = np.roll(luma, 9 * i, axis=0)
    ctx.h2d(d[0] + i * pitch, src.buf)
org = src.origin
cold = timed(lambda: ctx.check(L.x264cu_hpel_filter_batch(ctx.h, d[0] + org, src.stride, pitch, NH, w, h, d[1] + org, d[2] + org, d[3] + org, 1))) / NH
warm = timed(lambda: ctx.check(L.x264cu_hpel_filter(ctx.h, d[0] + org, src.stride, w, h, d[1] + org, d[2] + org, d[3] + org, 1)), 50)
res["hpel_filter"] = {"cold": entry(cold, 4 * w * h, {"pictures_per_launch": NH, "MB_touched_per_launch": NH * 4 * pitch / 1e6}),
                      "warm": entry(warm, 4 * w * h, {"pictures_per_launch": 1})}
print(json.dumps(res))
ctx.close()
