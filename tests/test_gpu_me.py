"""GPU parity of x264cu_me_search_batch (the batched x264_me_search_ref twin) against the oracle, which tests/test_oracle_me.py
pins to the compiled reference: every partition size, DIA / HEX / UMH, every sub-pel level, weighted references, half-pel
thresholds, random predictors and windows, on textured, flat (tie-heavy) and noisy content.  (mv, cost) bit-exact."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe, make_ref_planes, PIXEL_W, PIXEL_H
from test_oracle_me import _content, W, H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_me()
    c = x.Context(0)
    yield c
    c.close()


def run_group(ctx, kind, method, subpel, me_range, satd, wt, rng, n_jobs=96):
    o = oracle()
    fenc_l, ref_l = _content(kind, rng)
    planes = make_ref_planes(ref_l)
    st = planes[0].stride
    fenc = PaddedPlane(W, H, stride=st)
    fenc.inner()[:] = fenc_l
    if wt[0]:
        wplane = PaddedPlane(W, H, stride=st)
        ow = OrcWeight(*wt)
        o.orc_weight_scale_plane(ptr(wplane.buf), st, ptr(planes[0].buf), st, st, H + 64, C.byref(ow))
    else:
        wplane = planes[0]
    mv_range = 64
    n = 2 * 4 * mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    lam = int(rng.choice([1, 4]))
    o.orc_cost_mv_table(tab, n, lam)
    jobs = np.zeros(n_jobs, x.me_job_dtype)
    want = []
    for k in range(n_jobs):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
        mvr = 4 * mv_range
        lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
        lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
        i_mvc = int(rng.integers(0, 9))
        spread = int(rng.choice([2, 12, 50]))
        mvp = rng.integers(-spread, spread + 1, 2)
        mvcs = rng.integers(-spread, spread + 1, (8, 2))
        if rng.random() < 0.3:
            mvp[:] = 0
        if rng.random() < 0.3 and i_mvc:
            mvcs[0] = mvp
        use_thresh = rng.random() < 0.2
        thresh = int(rng.integers(50, 3000)) if use_thresh else -1
        off = planes[0].off(bx, by)
        j = jobs[k]
        j["i_pixel"], j["fenc_off"], j["ref_off"] = ip, fenc.off(bx, by), off
        j["mvp"], j["mvc"], j["i_mvc"] = mvp, mvcs, i_mvc
        j["mv_min_spel"], j["mv_max_spel"], j["halfpel_thresh"] = lim_min, lim_max, thresh
        c = OrcMeCtx()
        c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = method, subpel, me_range, satd
        for i in range(2):
            c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
            c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = lim_min[i] >> 2, lim_max[i] >> 2
        m = OrcMe()
        m.i_pixel = ip
        m.p_cost_mv = tab.ctypes.data + 2 * n
        for i in range(4):
            m.p_fref[i] = planes[i].buf.ctypes.data + off
        m.p_fref_w = wplane.buf.ctypes.data + off
        m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
        m.fenc_stride, m.stride = st, st
        m.weight = OrcWeight(*wt)
        m.mvp[0], m.mvp[1] = int(mvp[0]), int(mvp[1])
        mvc_arr = np.ascontiguousarray(mvcs.astype(np.int16))
        th = C.c_int(thresh)
        o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), i_mvc, C.byref(th) if use_thresh else None)
        want.append((m.mv[0], m.mv[1], m.cost, th.value if use_thresh else -1))
    d_fenc = ctx.upload(fenc.buf)
    d_pl = [ctx.upload(p.buf) for p in planes]
    d_w = ctx.upload(wplane.buf) if wt[0] else d_pl[0]
    params = x.MeParams(method, subpel, me_range, satd, lam, mv_range, *wt)
    res = x.me_search_batch(ctx, params, d_fenc, st, d_pl, d_w, st, jobs)
    for p in [d_fenc] + d_pl + ([d_w] if wt[0] else []):
        ctx.free(p)
    for k in range(n_jobs):
        got = (int(res[k]["mv"][0]), int(res[k]["mv"][1]), int(res[k]["cost"]), int(res[k]["halfpel_thresh"]))
        assert got == want[k], (kind, method, subpel, me_range, satd, wt, k, jobs[k], got, want[k])


@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
@pytest.mark.parametrize("method", [0, 1, 2, 3])
def test_me_search_batch_matches_oracle(ctx, kind, method):
    rng = np.random.default_rng(17 * method + len(kind))
    for subpel in (0, 1, 2, 3, 4, 5, 6, 7, 9):
        me_range = int(rng.choice([4, 8, 16] if method != 2 else [16, 24, 32]))
        satd = int(subpel > 1 and rng.random() < 0.8)
        wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.3 else (0, 0, 0, 0)
        run_group(ctx, kind, method, subpel, me_range, satd, wt, rng)


def test_me_search_batch_4k_frame_of_macroblocks(ctx):
    """BASELINE config 3 shape: every 16x16 macroblock of a 4K frame, UMH merange 64, subme 9 -- against the oracle on a
    random subset (the oracle is scalar), and a determinism check on all 32 400 jobs"""
    _libs._bind_mc()
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
    ctx.check(ctx.L.x264cu_hpel_filter(ctx.h, d_pl[0] + F.origin, st, w, h, d_pl[1] + F.origin, d_pl[2] + F.origin, d_pl[3] + F.origin, 1))
    planes = [ctx.download(p, (h + 64, st), np.uint8) for p in d_pl]
    mbw, mbh = w // 16, h // 16
    jobs = np.zeros(mbw * mbh, x.me_job_dtype)
    yy, xx = np.meshgrid(np.arange(mbh), np.arange(mbw), indexing="ij")
    jobs["i_pixel"] = 0
    jobs["fenc_off"] = (fenc.origin + yy * 16 * st + xx * 16).reshape(-1)
    jobs["ref_off"] = jobs["fenc_off"]
    jobs["mvp"] = rng.integers(-20, 21, (mbw * mbh, 2))
    jobs["mvc"] = rng.integers(-30, 31, (mbw * mbh, 8, 2))
    jobs["i_mvc"] = rng.integers(0, 6, mbw * mbh)
    mvr = 4 * 512
    jobs["mv_min_spel"][:, 0] = np.maximum(4 * (-16 * xx - 24), -mvr).reshape(-1)
    jobs["mv_min_spel"][:, 1] = np.maximum(4 * (-16 * yy - 24), -mvr).reshape(-1)
    jobs["mv_max_spel"][:, 0] = np.minimum(4 * (16 * (mbw - xx - 1) + 24), mvr - 1).reshape(-1)
    jobs["mv_max_spel"][:, 1] = np.minimum(4 * (16 * (mbh - yy - 1) + 24), mvr - 1).reshape(-1)
    jobs["halfpel_thresh"] = -1
    d_fenc = ctx.upload(fenc.buf)
    params = x.MeParams(2, 9, 64, 1, 1, 512, 0, 0, 0, 0)
    r1 = x.me_search_batch(ctx, params, d_fenc, st, d_pl, d_pl[0], st, jobs)
    r2 = x.me_search_batch(ctx, params, d_fenc, st, d_pl, d_pl[0], st, jobs)
    assert np.array_equal(r1["mv"], r2["mv"]) and np.array_equal(r1["cost"], r2["cost"])
    o = oracle()
    n = 2 * 4 * 512
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, 1)
    for k in rng.choice(mbw * mbh, 300, replace=False):
        j = jobs[k]
        c = OrcMeCtx()
        c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 2, 9, 64, 1
        for i in range(2):
            c.mv_min_spel[i], c.mv_max_spel[i] = int(j["mv_min_spel"][i]), int(j["mv_max_spel"][i])
            c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = int(j["mv_min_spel"][i]) >> 2, int(j["mv_max_spel"][i]) >> 2
        m = OrcMe()
        m.i_pixel = 0
        m.p_cost_mv = tab.ctypes.data + 2 * n
        for i in range(4):
            m.p_fref[i] = planes[i].ctypes.data + int(j["ref_off"])
        m.p_fref_w = m.p_fref[0]
        m.p_fenc = fenc.buf.ctypes.data + int(j["fenc_off"])
        m.fenc_stride, m.stride = st, st
        m.weight = OrcWeight(0, 0, 0, 0)
        m.mvp[0], m.mvp[1] = int(j["mvp"][0]), int(j["mvp"][1])
        mvc_arr = np.ascontiguousarray(j["mvc"])
        o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), int(j["i_mvc"]), None)
        assert (m.mv[0], m.mv[1], m.cost) == (int(r1[k]["mv"][0]), int(r1[k]["mv"][1]), int(r1[k]["cost"])), (k, j)
