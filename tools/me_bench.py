#!/usr/bin/env python
"""BASELINE configs[2] shape: full-resolution motion search of every 16x16 macroblock of a 3840x2160 picture (preset slower:
UMH, merange 64... subme 9 here) through x264cu_me_search_batch, timed with CUDA events, beside the oracle port on one host
core for a sample of the same jobs.  Prints one JSON line.  Not part of the product path (bench.py's lines are the lookahead and
the SATD table)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import x264_b200 as x
import _libs
from _libs import oracle, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe

_libs._bind_me()
_libs._bind_mc()
ctx = x.Context(0)
w, h = 3840, 2160
rng = np.random.default_rng(4)
base = _libs.synth_luma(w + 16, h + 16, seed=99)
ref_l = np.ascontiguousarray(base[8:8 + h, 8:8 + w])
fenc_l = np.ascontiguousarray(base[5:5 + h, 11:11 + w])
fenc = PaddedPlane(w, h)
fenc.inner()[:] = fenc_l
st = fenc.stride
F = PaddedPlane(w, h, stride=st)
F.inner()[:] = ref_l
d_pl = [ctx.upload(F.buf)] + [ctx.malloc(F.buf.size + 256) for _ in range(3)]
ctx.timer_start()
ctx.check(ctx.L.x264cu_hpel_filter(ctx.h, d_pl[0] + F.origin, st, w, h, d_pl[1] + F.origin, d_pl[2] + F.origin, d_pl[3] + F.origin, 1))
hpel_ms = ctx.timer_stop()
mbw, mbh = w // 16, h // 16
n = mbw * mbh
res = {"picture": "%dx%d" % (w, h), "jobs": n, "hpel_filter_ms": hpel_ms}
for name, method, me_range, subpel in (("umh_merange64_subme9", 2, 64, 9), ("umh_merange16_subme7", 2, 16, 7), ("hex_merange16_subme7", 1, 16, 7),
                                       ("dia_merange16_subme2", 0, 16, 2), ("esa_merange16_subme7", 3, 16, 7)):
    jobs = np.zeros(n, x.me_job_dtype)
    yy, xx = np.meshgrid(np.arange(mbh), np.arange(mbw), indexing="ij")
    jobs["i_pixel"] = 0
    jobs["fenc_off"] = (fenc.origin + yy * 16 * st + xx * 16).reshape(-1)
    jobs["ref_off"] = jobs["fenc_off"]
    jobs["mvp"] = rng.integers(-20, 21, (n, 2))
    jobs["mvc"][:, :8] = rng.integers(-30, 31, (n, 8, 2))
    jobs["i_mvc"] = rng.integers(0, 6, n)
    mvr = 4 * 512
    jobs["mv_min_spel"][:, 0] = np.maximum(4 * (-16 * xx - 24), -mvr).reshape(-1)
    jobs["mv_min_spel"][:, 1] = np.maximum(4 * (-16 * yy - 24), -mvr).reshape(-1)
    jobs["mv_max_spel"][:, 0] = np.minimum(4 * (16 * (mbw - xx - 1) + 24), mvr - 1).reshape(-1)
    jobs["mv_max_spel"][:, 1] = np.minimum(4 * (16 * (mbh - yy - 1) + 24), mvr - 1).reshape(-1)
    jobs["halfpel_thresh"] = -1
    d_fenc = ctx.upload(fenc.buf)
    d_jobs = ctx.upload(jobs)
    d_res = ctx.malloc(n * x.me_result_dtype.itemsize)
    params = x.MeParams(method, subpel, me_range, 1, 1, 512, 0, 0, 0, 0)
    arr = (C.c_void_p * 4)(*d_pl)
    call = lambda: ctx.check(ctx.L.x264cu_me_search_batch(ctx.h, C.byref(params), d_fenc, st, arr, d_pl[0], st, d_jobs, n, d_res))
    call(); ctx.sync()
    ctx.timer_start()
    reps = 3
    for _ in range(reps):
        call()
    ms = ctx.timer_stop() / reps
    res[name] = {"ms_per_picture": ms, "searches_per_s": n / (ms * 1e-3)}
    if name.startswith("umh_merange64"):
        # the oracle port on one host core, 400 of the same jobs
        planes = [ctx.download(p, (h + 64, st), np.uint8) for p in d_pl]
        o = oracle()
        nt = 2 * 4 * 512
        tab = np.zeros(2 * nt + 1, np.uint16)
        o.orc_cost_mv_table(tab, nt, 1)
        sample = rng.choice(n, 400, replace=False)
        t0 = time.perf_counter()
        for k in sample:
            j = jobs[k]
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = method, subpel, me_range, 1
            for i in range(2):
                c.mv_min_spel[i], c.mv_max_spel[i] = int(j["mv_min_spel"][i]), int(j["mv_max_spel"][i])
                c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = int(j["mv_min_spel"][i]) >> 2, int(j["mv_max_spel"][i]) >> 2
            m = OrcMe()
            m.i_pixel = 0
            m.p_cost_mv = tab.ctypes.data + 2 * nt
            for i in range(4):
                m.p_fref[i] = planes[i].ctypes.data + int(j["ref_off"])
            m.p_fref_w = m.p_fref[0]
            m.p_fenc = fenc.buf.ctypes.data + int(j["fenc_off"])
            m.fenc_stride, m.stride = st, st
            m.weight = OrcWeight(0, 0, 0, 0)
            m.mvp[0], m.mvp[1] = int(j["mvp"][0]), int(j["mvp"][1])
            mvc_arr = np.ascontiguousarray(j["mvc"])
            o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), int(j["i_mvc"]), None)
        dt = time.perf_counter() - t0
        res[name]["cpu_port_searches_per_s_1_core"] = 400 / dt
    for p in (d_fenc, d_jobs, d_res):
        ctx.free(p)
print(json.dumps(res))
ctx.close()
