"""Pins oracle/oracle_me.c against the compiled reference's x264_me_search_ref + refine_subpel (encoder/me.c:182-992):
random predictors / candidate lists / windows, DIA, HEX and UMH, every sub-pel level, textured, flat (tie-heavy)
and noisy content, weighted and unweighted references."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import (oracle, ref, have_ref, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe, XrefMeArgs, XrefChroma, synth_luma,
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


def chroma_pair(kind, rng, st):
    """NV12 chroma of the two pictures (W bytes x H/2 rows, padded like the luma: a superset of the reference's 16-pixel border);
    the reference picture's chroma is the source's moved by a fraction of a pixel so that the eighth-pel weights matter"""
    base = synth_luma(W + 16, H // 2 + 16, seed=int(rng.integers(1 << 30)), kind="noise" if kind == "noise" else "texture")
    fc, rc = PaddedPlane(W, H // 2, stride=st), PaddedPlane(W, H // 2, stride=st)
    fc.inner()[:] = base[4:4 + H // 2, 4:4 + W]
    rc.inner()[:] = np.clip(base[5:5 + H // 2, 6:6 + W].astype(np.int16) + rng.integers(-2, 3, (H // 2, W)), 0, 255).astype(np.uint8)
    fc.fill_border()
    rc.fill_border()
    return fc, rc


def run_case(hnd, subme_param, cost_tab, centre, kind, rng, n_blocks, chroma=False):
    o, r = oracle(), ref()
    fenc_l, ref_l = _content(kind, rng)
    planes = make_ref_planes(ref_l)
    st = planes[0].stride
    fenc = PaddedPlane(W, H, stride=st)
    fenc.inner()[:] = fenc_l
    if chroma:
        fenc_c, ref_c = chroma_pair(kind, rng, st)
    for _ in range(n_blocks):
        ip = int(rng.integers(0, 4 if chroma and rng.random() < 0.8 else 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
        method = int(rng.integers(0, 3))
        subpel = int(rng.choice([5, 6, 7, 9, 2, 4] if chroma else [0, 1, 2, 3, 4, 5, 6, 7, 9]))
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
        if chroma:
            xc = XrefChroma()
            coff = ref_c.off(bx & ~1, by // 2)             # NV12: chroma pair x/2 starts at byte 2*(x/2)
            xc.fenc_uv, xc.fenc_uv_stride = fenc_c.buf.ctypes.data + fenc_c.off(bx & ~1, by // 2), st
            xc.fref_uv, xc.fref_uv_stride = ref_c.buf.ctypes.data + coff, st
            wuv = [(1, int(rng.integers(40, 90)), int(rng.integers(0, 7)), int(rng.integers(-6, 7))) if rng.random() < 0.3 else (0, 0, 0, 0)
                   for _ in range(2)]
            for k in range(2):
                for j in range(4):
                    xc.wt[k][j] = wuv[k][j]
            r.xref_me_search_chroma(hnd, C.byref(a), ptr(fenc.buf, fenc.off(bx, by)), st,
                                    *[ptr(p.buf, off) for p in planes], ptr(wplane.buf, off), st, C.byref(xc))
        else:
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
        if chroma:
            c.chroma_me = 1
            m.p_fref_uv, m.stride_uv, m.p_fenc_uv, m.fenc_uv_stride = xc.fref_uv, st, xc.fenc_uv, st
            m.weight_uv[0], m.weight_uv[1] = OrcWeight(*wuv[0]), OrcWeight(*wuv[1])
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
def test_me_search_chroma_matches_reference(kind):
    """chroma ME (h->mb.b_chroma_me: P slices at subme >= 5, common/macroblock.c:507): the chroma branch of COST_MV_SATD
    (me.c:826-857) and mc_chroma (common/mc.c:251-283), weighted chroma planes included"""
    _libs._bind_me()
    r = ref()
    hnd = r.xref_open(W, H, b"medium", b"subme=7", 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(4242 + len(kind))
        for _ in range(6):
            run_case(hnd, 7, tab, n, kind, rng, 60, chroma=True)
    finally:
        r.xref_close(hnd)


def _exhaustive_case(kind, method, open_opts, seed, subpels, ranges, weighted_share, reps=3, blocks=40):
    """ESA / TESA against the reference on a frame the reference builds itself (its integral image comes from x264_frame_filter)"""
    _libs._bind_me()
    o, r = oracle(), ref()
    r.xref_me_search_frame.argtypes = [C.c_void_p, C.POINTER(XrefMeArgs), C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
    hnd = r.xref_open(W, H, b"medium", open_opts, 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(seed + len(kind))
        for rep in range(reps):
            fenc_l, ref_l = _content(kind, rng)
            ref_l = np.ascontiguousarray(ref_l)
            planes = make_ref_planes(ref_l)
            st = planes[0].stride
            fenc = PaddedPlane(W, H, stride=st)
            fenc.inner()[:] = fenc_l
            for _ in range(blocks):
                ip = int(rng.integers(0, 7))
                bw, bh = PIXEL_W[ip], PIXEL_H[ip]
                bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
                by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
                subpel = int(rng.choice(subpels))
                me_range = int(rng.choice(ranges))
                mvr = 4 * 64
                lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
                lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
                i_mvc = int(rng.integers(0, 5))
                mvp = rng.integers(-12, 13, 2)
                mvcs = rng.integers(-12, 13, (16, 2))
                wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < weighted_share else (0, 0, 0, 0)
                a = XrefMeArgs()
                a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = ip, method, subpel, me_range, 12
                for i in range(2):
                    a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i] = lim_min[i], lim_max[i], int(mvp[i])
                a.i_mvc = i_mvc
                for i in range(16):
                    a.mvc[i][0], a.mvc[i][1] = int(mvcs[i][0]), int(mvcs[i][1])
                a.wt_en, a.wt_scale, a.wt_denom, a.wt_offset = wt
                assert r.xref_me_search_frame(hnd, C.byref(a), ptr(fenc.buf, fenc.off(bx, by)), st, ptr(ref_l), W, bx, by) == 0
                if wt[0]:
                    wplane = PaddedPlane(W, H, stride=st)
                    ow = OrcWeight(*wt)
                    o.orc_weight_scale_plane(ptr(wplane.buf), st, ptr(planes[0].buf), st, st, H + 2 * 32, C.byref(ow))
                else:
                    wplane = planes[0]
                c = OrcMeCtx()
                c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = method, subpel, me_range, 1
                for i in range(2):
                    c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
                    c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = lim_min[i] >> 2, lim_max[i] >> 2
                off = planes[0].off(bx, by)
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
                o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), i_mvc, None)
                assert (m.mv[0], m.mv[1], m.cost) == (a.mv[0], a.mv[1], a.cost), (kind, method, ip, subpel, me_range, tuple(mvp), i_mvc, bx, by, wt)
    finally:
        r.xref_close(hnd)


@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
def test_esa_matches_reference(kind):
    """exhaustive search (me.c:618-771, the ADS + SAD branch): every partition size, three ranges, sub-pel levels 0 / 2 / 7,
    a third of the searches against a weighted reference (whose ADS prefilter reads the UNWEIGHTED plane's sums)"""
    _exhaustive_case(kind, 3, b"me=esa:merange=32:subme=7:partitions=all", 77, [0, 2, 7], [4, 8, 16], 0.35)


@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
def test_tesa_matches_reference(kind):
    """transformed exhaustive search (me.c:656-747): fpelcmp is SATD for the whole search (encoder.c:1409-1427); ADS threshold,
    SAD threshold, the thinned candidate list, then SATD; ranges on each side of the sad_thresh steps (16 / 24 / 32)"""
    _exhaustive_case(kind, 4, b"me=tesa:merange=32:subme=7:partitions=all", 177, [0, 2, 3, 7, 9], [4, 8, 16, 24, 32], 0.25)


@pytest.mark.parametrize("subme_param", [1, 7])
@pytest.mark.parametrize("kind", ["texture", "flat"])
def test_refine_bidir_satd_matches_reference(subme_param, kind):
    """x264_me_refine_bidir_satd (me.c:1027-1183): joint refinement of a bi-predicted pair, every partition size, implicit
    bipred weights 32 / 21 / 43 / -10 (pixel_avg_weight), vectors near the window edge (early return) included"""
    _libs._bind_me()
    o, r = oracle(), ref()
    PP = C.POINTER(C.c_void_p)
    r.xref_me_refine_bidir_satd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    r.xref_me_refine_bidir_satd.restype = None
    o.orc_me_refine_bidir_satd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    o.orc_me_refine_bidir_satd.restype = None
    hnd = r.xref_open(W, H, b"medium", ("subme=%d" % subme_param).encode(), 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(31 * subme_param + len(kind))
        moved = 0
        for rep in range(4):
            fenc_l, ref0_l = _content(kind, rng)
            _, ref1_l = _content(kind, rng)
            if kind == "texture":                    # list 1 = the same scene displaced the other way
                ref1_l = np.ascontiguousarray(np.roll(ref0_l, (3, -2), (0, 1)))
            pl0, pl1 = make_ref_planes(np.ascontiguousarray(ref0_l)), make_ref_planes(np.ascontiguousarray(ref1_l))
            st = pl0[0].stride
            fenc = PaddedPlane(W, H, stride=st)
            fenc.inner()[:] = fenc_l
            for _ in range(50):
                ip = int(rng.integers(0, 7))
                bw, bh = PIXEL_W[ip], PIXEL_H[ip]
                bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
                by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
                mvr = 4 * 64
                lim_min = np.array([max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)], np.int32)
                lim_max = np.array([min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)], np.int32)
                spread = int(rng.choice([6, 30, 120]))
                mv0 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max).astype(np.int16)
                mv1 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max).astype(np.int16)
                mvp0 = (mv0 + rng.integers(-6, 7, 2)).astype(np.int16)
                mvp1 = (mv1 + rng.integers(-6, 7, 2)).astype(np.int16)
                weight = int(rng.choice([32, 32, 21, 43, -10, 64]))
                off = pl0[0].off(bx, by)
                a0, a1 = mv0.copy(), mv1.copy()
                f0 = (C.c_void_p * 4)(*[p.buf.ctypes.data + off for p in pl0])
                f1 = (C.c_void_p * 4)(*[p.buf.ctypes.data + off for p in pl1])
                r.xref_me_refine_bidir_satd(hnd, ip, 12, ptr(fenc.buf, fenc.off(bx, by)), st, f0, f1, st, ptr(a0), ptr(mvp0), ptr(a1), ptr(mvp1),
                                            weight, ptr(lim_min), ptr(lim_max))
                c = OrcMeCtx()
                c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, 7, 16, int(subme_param > 1)
                for i in range(2):
                    c.mv_min_spel[i], c.mv_max_spel[i] = int(lim_min[i]), int(lim_max[i])
                ms = []
                for pl, mv, mvp in ((pl0, mv0, mvp0), (pl1, mv1, mvp1)):
                    m = OrcMe()
                    m.i_pixel = ip
                    m.p_cost_mv = tab.ctypes.data + 2 * n
                    for i in range(4):
                        m.p_fref[i] = pl[i].buf.ctypes.data + off
                    m.p_fref_w = pl[0].buf.ctypes.data + off
                    m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
                    m.fenc_stride, m.stride = st, st
                    m.weight = OrcWeight(0, 0, 0, 0)
                    m.mvp[0], m.mvp[1] = int(mvp[0]), int(mvp[1])
                    m.mv[0], m.mv[1] = int(mv[0]), int(mv[1])
                    ms.append(m)
                o.orc_me_refine_bidir_satd(C.byref(c), C.byref(ms[0]), C.byref(ms[1]), weight)
                got = (ms[0].mv[0], ms[0].mv[1], ms[1].mv[0], ms[1].mv[1])
                assert got == (a0[0], a0[1], a1[0], a1[1]), (kind, ip, bx, by, tuple(mv0), tuple(mv1), weight)
                moved += got != (mv0[0], mv0[1], mv1[0], mv1[1])
        assert moved > 20
    finally:
        r.xref_close(hnd)


@pytest.mark.parametrize("subme_param", [1, 7])
@pytest.mark.parametrize("kind", ["texture", "flat"])
def test_refine_qpel_matches_reference(subme_param, kind):
    """x264_me_refine_qpel / x264_me_refine_qpel_refdupe (me.c:800-814): the final refinement of a stored vector, every h->mb
    sub-pel level (level 1 takes the simplified SAD quarter-pel diamond, me.c:965-985), weights, half-pel thresholds"""
    _libs._bind_me()
    o, r = oracle(), ref()
    r.xref_me_refine_qpel.argtypes = [C.c_void_p, C.POINTER(XrefMeArgs), C.c_int, C.c_int, C.c_void_p, C.c_ssize_t] + [C.c_void_p] * 4 + [C.c_ssize_t]
    r.xref_me_refine_qpel.restype = None
    o.orc_me_refine_qpel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    o.orc_me_refine_qpel.restype = None
    hnd = r.xref_open(W, H, b"medium", ("subme=%d" % subme_param).encode(), 0)
    assert hnd
    try:
        n = 2 * 4 * r.xref_param(hnd, b"mvrange")
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table_qp(hnd, 12, tab, n)
        rng = np.random.default_rng(9 * subme_param + len(kind))
        moved = 0
        for rep in range(4):
            fenc_l, ref_l = _content(kind, rng)
            planes = make_ref_planes(ref_l)
            st = planes[0].stride
            fenc = PaddedPlane(W, H, stride=st)
            fenc.inner()[:] = fenc_l
            for _ in range(60):
                ip = int(rng.integers(0, 7))
                bw, bh = PIXEL_W[ip], PIXEL_H[ip]
                bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
                by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
                mode = int(rng.integers(0, 2))
                subpel = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 11]))
                mvr = 4 * 64
                lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
                lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
                spread = int(rng.choice([4, 20, 90]))
                mv = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
                mvp = mv + rng.integers(-6, 7, 2)
                cost = int(rng.integers(50, 6000))
                ref_cost = int(rng.integers(0, 5))
                wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.25 else (0, 0, 0, 0)
                use_thresh = mode == 1 and rng.random() < 0.4
                thresh = int(rng.integers(50, 6000))
                off = planes[0].off(bx, by)
                a = XrefMeArgs()
                a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = ip, 1, subpel, 16, 12
                for i in range(2):
                    a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i], a.mv[i] = lim_min[i], lim_max[i], int(mvp[i]), int(mv[i])
                a.cost = cost
                a.wt_en, a.wt_scale, a.wt_denom, a.wt_offset = wt
                a.use_thresh, a.halfpel_thresh = int(use_thresh), thresh
                r.xref_me_refine_qpel(hnd, C.byref(a), mode, ref_cost, ptr(fenc.buf, fenc.off(bx, by)), st, *[ptr(p.buf, off) for p in planes], st)
                c = OrcMeCtx()
                c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, subpel, 16, int(subme_param > 1)
                for i in range(2):
                    c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
                m = OrcMe()
                m.i_pixel = ip
                m.p_cost_mv = tab.ctypes.data + 2 * n
                for i in range(4):
                    m.p_fref[i] = planes[i].buf.ctypes.data + off
                m.p_fref_w = planes[0].buf.ctypes.data + off
                m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
                m.fenc_stride, m.stride = st, st
                m.weight = OrcWeight(*wt)
                m.mvp[0], m.mvp[1], m.mv[0], m.mv[1], m.cost = int(mvp[0]), int(mvp[1]), int(mv[0]), int(mv[1]), cost
                th = C.c_int(thresh)
                o.orc_me_refine_qpel(C.byref(c), C.byref(m), mode, ref_cost, C.byref(th) if use_thresh else None)
                key = (kind, ip, mode, subpel, tuple(mv), tuple(mvp), cost, wt, use_thresh, thresh)
                assert (m.mv[0], m.mv[1], m.cost) == (a.mv[0], a.mv[1], a.cost), key
                if use_thresh:
                    assert th.value == a.thresh_out, key
                moved += (m.mv[0], m.mv[1]) != (int(mv[0]), int(mv[1]))
        assert moved > 20
    finally:
        r.xref_close(hnd)
