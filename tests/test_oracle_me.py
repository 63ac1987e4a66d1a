"""Pins oracle/oracle_me.c against the compiled reference's x264_me_search_ref + refine_subpel (encoder/me.c:182-992):
random predictors / candidate lists / windows, DIA, HEX and UMH, every sub-pel level, textured, flat (tie-heavy)
and noisy content, weighted and unweighted references."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import (oracle, ref, have_ref, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe, XrefMeArgs, synth_luma,
                   make_ref_planes, PIXEL_W, PIXEL_H)

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

W, H = 112, 96


def _content(kind, rng):
    if kind == "texture":
        ref_l = synth_luma(W + 16, H + 16, seed=int(rng.integers(1 << 30)))
        dx, dy = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        fenc = ref_l[dy:dy + H, dx:dx + W].astype(np.int16) + rng.integers(-3, 4, (H, W))
        return np.clip(fenc, 0, 255).astype(np.uint8), np.ascontiguousarray(ref_l[4:4 + H, 4:4 + W])
    if kind == "flat":
        return (rng.integers(126, 130, (H, W)).astype(np.uint8), rng.integers(126, 130, (H, W)).astype(np.uint8))
    return rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)


def run_case(hnd, subme_param, cost_tab, centre, kind, rng, n_blocks):
    o, r = oracle(), ref()
    fenc_l, ref_l = _content(kind, rng)
    planes = make_ref_planes(ref_l)
    st = planes[0].stride
    fenc = PaddedPlane(W, H, stride=st)
    fenc.inner()[:] = fenc_l
    for _ in range(n_blocks):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
        method = int(rng.integers(0, 3))
        subpel = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 9]))
        me_range = int(rng.choice([4, 8, 16] if method < 2 else [16, 24, 32]))
        mvr = 4 * 64
        lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
        lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
        i_mvc = int(rng.integers(0, 5))
        spread = int(rng.choice([2, 12, 60]))
        mvp = rng.integers(-spread, spread + 1, 2)
        mvcs = rng.integers(-spread, spread + 1, (16, 2))
        if rng.random() < 0.3:
            mvp[:] = 0
        if rng.random() < 0.3 and i_mvc:
            mvcs[0] = mvp
        wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.25 else (0, 0, 0, 0)
        if wt[0]:
            wplane = PaddedPlane(W, H, stride=st)
            ow = OrcWeight(*wt)
            o.orc_weight_scale_plane(ptr(wplane.buf), st, ptr(planes[0].buf), st, st, H + 2 * 32, C.byref(ow))
        else:
            wplane = planes[0]
        off = planes[0].off(bx, by)
        a = XrefMeArgs()
        a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = ip, method, subpel, me_range, 12
        for i in range(2):
            a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i] = lim_min[i], lim_max[i], int(mvp[i])
        a.i_mvc = i_mvc
        for i in range(16):
            a.mvc[i][0], a.mvc[i][1] = int(mvcs[i][0]), int(mvcs[i][1])
        a.wt_en, a.wt_scale, a.wt_denom, a.wt_offset = wt
        use_thresh = rng.random() < 0.2
        a.use_thresh, a.halfpel_thresh = int(use_thresh), int(rng.integers(50, 3000))
        r.xref_me_search(hnd, C.byref(a), ptr(fenc.buf, fenc.off(bx, by)), st,
                         *[ptr(p.buf, off) for p in planes], ptr(wplane.buf, off), st)
        c = OrcMeCtx()
        c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = method, subpel, me_range, int(subme_param > 1)
        for i in range(2):
            c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
            c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = lim_min[i] >> 2, lim_max[i] >> 2
        m = OrcMe()
        m.i_pixel = ip
        m.p_cost_mv = cost_tab.ctypes.data + 2 * centre
        for i in range(4):
            m.p_fref[i] = planes[i].buf.ctypes.data + off
        m.p_fref_w = wplane.buf.ctypes.data + off
        m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
        m.fenc_stride, m.stride = st, st
        m.weight = OrcWeight(*wt)
        m.mvp[0], m.mvp[1] = int(mvp[0]), int(mvp[1])
        mvc_arr = np.ascontiguousarray(mvcs.astype(np.int16))
        th = C.c_int(a.halfpel_thresh)
        o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), i_mvc, C.byref(th) if use_thresh else None)
        key = (kind, ip, method, subpel, me_range, tuple(mvp), i_mvc, wt, bx, by, use_thresh)
        assert (m.mv[0], m.mv[1], m.cost) == (a.mv[0], a.mv[1], a.cost), key
        if use_thresh:
            assert th.value == a.thresh_out, key


@pytest.mark.parametrize("subme_param", [1, 7])
@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
def test_me_search_matches_reference(subme_param, kind):
    _libs._bind_me()
    r = ref()
    hnd = r.xref_open(W, H, b"medium", ("subme=%d" % subme_param).encode(), 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(100 * subme_param + len(kind))
        for _ in range(6):
            run_case(hnd, subme_param, tab, n, kind, rng, 60)
    finally:
        r.xref_close(hnd)


@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
def test_esa_matches_reference(kind):
    """exhaustive search (me.c:618-771, the ADS + SAD branch): the reference runs on a frame it builds itself (its integral
    image comes from x264_frame_filter); every partition size, three ranges, sub-pel levels 0 / 2 / 7"""
    _libs._bind_me()
    o, r = oracle(), ref()
    r.xref_me_search_frame.argtypes = [C.c_void_p, C.POINTER(XrefMeArgs), C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
    hnd = r.xref_open(W, H, b"medium", b"me=esa:merange=32:subme=7:partitions=all", 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(77 + len(kind))
        for rep in range(3):
            fenc_l, ref_l = _content(kind, rng)
            ref_l = np.ascontiguousarray(ref_l)
            planes = make_ref_planes(ref_l)
            st = planes[0].stride
            fenc = PaddedPlane(W, H, stride=st)
            fenc.inner()[:] = fenc_l
            for _ in range(40):
                ip = int(rng.integers(0, 7))
                bw, bh = PIXEL_W[ip], PIXEL_H[ip]
                bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
                by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
                subpel = int(rng.choice([0, 2, 7]))
                me_range = int(rng.choice([4, 8, 16]))
                mvr = 4 * 64
                lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
                lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
                i_mvc = int(rng.integers(0, 5))
                mvp = rng.integers(-12, 13, 2)
                mvcs = rng.integers(-12, 13, (16, 2))
                a = XrefMeArgs()
                a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = ip, 3, subpel, me_range, 12
                for i in range(2):
                    a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i] = lim_min[i], lim_max[i], int(mvp[i])
                a.i_mvc = i_mvc
                for i in range(16):
                    a.mvc[i][0], a.mvc[i][1] = int(mvcs[i][0]), int(mvcs[i][1])
                assert r.xref_me_search_frame(hnd, C.byref(a), ptr(fenc.buf, fenc.off(bx, by)), st, ptr(ref_l), W, bx, by) == 0
                c = OrcMeCtx()
                c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 3, subpel, me_range, 1
                for i in range(2):
                    c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
                    c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = lim_min[i] >> 2, lim_max[i] >> 2
                off = planes[0].off(bx, by)
                m = OrcMe()
                m.i_pixel = ip
                m.p_cost_mv = tab.ctypes.data + 2 * n
                for i in range(4):
                    m.p_fref[i] = planes[i].buf.ctypes.data + off
                m.p_fref_w = planes[0].buf.ctypes.data + off
                m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
                m.fenc_stride, m.stride = st, st
                m.weight = OrcWeight(0, 0, 0, 0)
                m.mvp[0], m.mvp[1] = int(mvp[0]), int(mvp[1])
                mvc_arr = np.ascontiguousarray(mvcs.astype(np.int16))
                o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), i_mvc, None)
                assert (m.mv[0], m.mv[1], m.cost) == (a.mv[0], a.mv[1], a.cost), (kind, ip, subpel, me_range, tuple(mvp), i_mvc, bx, by)
    finally:
        r.xref_close(hnd)
