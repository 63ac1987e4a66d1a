"""Runs the golden cases (tests/golden/cases.py) through one of three implementations and returns plain numpy arrays:
  "ref"    the compiled reference (oracle/_ref/libx264ref.so) -- used by make_golden.py only
  "oracle" the CPU restatement (oracle/liboracle.so)
  "cuda"   the product, through the C ABI of libx264_b200.so (needs a B200)
Test infrastructure: nothing under x264_b200/ imports this."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402
from _libs import (oracle, ref, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe, XrefMeArgs, OrcLaParams, make_ref_planes,
                   PIXEL_W, PIXEL_H, PAD, cand_dtype)  # noqa: E402
from . import cases as G  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------------
def run_pixel(backend, pattern, ctx=None):
    a, b, fo, ro = G.pixel_case(pattern)
    n = len(fo)
    out = np.zeros((len(G.PIX_COMBOS), n), np.int32)
    if backend == "cuda":
        import x264_b200 as x
        cand = np.zeros(n, x.cand_dtype)
    else:
        cand = np.zeros(n, cand_dtype)
    cand["fenc_off"], cand["ref_off"] = fo, ro
    for k, (metric, ip) in enumerate(G.PIX_COMBOS):
        if backend == "cuda":
            out[k] = ctx.pixel_cmp_batch_host(metric, ip, a, G.PIX_STRIDE, b, G.PIX_STRIDE, cand)
        else:
            f = ref().xref_pixel_cmp_batch if backend == "ref" else oracle().orc_pixel_cmp_batch
            f(metric, ip, a, G.PIX_STRIDE, b, G.PIX_STRIDE, cand, n, out[k])
    return out, G.digest(a, b, fo, ro)


# ---------------------------------------------------------------------------------------------------------------------
def _expand16(luma):
    h, w = luma.shape
    W16, H16 = (w + 15) // 16 * 16, (h + 15) // 16 * 16
    src = np.zeros((H16, W16), np.uint8)
    src[:h, :w] = luma
    src[:h, w:] = luma[:, w - 1:w]
    src[h:, :] = src[h - 1:h, :]
    return src


def run_lowres(backend, ctx=None):
    _libs._bind_mc()
    luma = G.lowres_case()
    h, w = luma.shape
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    wl, ll = mbw * 8, mbh * 8
    if backend == "ref":
        r = ref()
        hnd = r.xref_open(w, h, b"medium", b"", 0)
        st = r.xref_param(hnd, b"stride_lowres")
        pb = st * (ll + 2 * PAD)
        out = np.zeros(4 * pb, np.uint8)
        assert r.xref_frame_lowres(hnd, ptr(luma), w, ptr(out)) == 0
        r.xref_close(hnd)
        planes = np.stack([out[i * pb:(i + 1) * pb].reshape(ll + 2 * PAD, st)[:, :wl + 2 * PAD] for i in range(4)])
    elif backend == "oracle":
        pl = [PaddedPlane(wl, ll) for _ in range(4)]
        arr = (C.c_void_p * 4)(*[p.buf.ctypes.data + p.origin for p in pl])
        src = _expand16(luma)
        oracle().orc_frame_init_lowres(ptr(src), src.shape[1], src.shape[1], src.shape[0], arr, pl[0].stride, wl, ll)
        planes = np.stack([p.view()[:, :wl + 2 * PAD] for p in pl])
    else:
        pl = PaddedPlane(wl, ll)
        st, pb = pl.stride, pl.buf.size
        stride_src = (w + 63) // 64 * 64
        host = np.zeros((h, stride_src), np.uint8)
        host[:, :w] = luma
        d_src = ctx.upload(host)
        d_planes = ctx.malloc(4 * pb + 256)
        ctx.check(ctx.L.x264cu_memset(ctx.h, d_planes, 0, 4 * pb))
        darr = (C.c_void_p * 4)(*[d_planes + i * pb + pl.origin for i in range(4)])
        ctx.check(ctx.L.x264cu_frame_init_lowres(ctx.h, d_src, stride_src, w, h, darr, st))
        got = ctx.download(d_planes, (4, ll + 2 * PAD, st), np.uint8)
        planes = np.ascontiguousarray(got[:, :, :wl + 2 * PAD])
        ctx.free(d_src)
        ctx.free(d_planes)
    return np.ascontiguousarray(planes), G.digest(luma)


def run_hpel(backend, ctx=None):
    _libs._bind_mc()
    luma = G.hpel_case()
    h, w = luma.shape
    if backend == "ref":
        r = ref()
        hnd = r.xref_open(w, h, b"medium", b"", 0)
        st = r.xref_param(hnd, b"stride")
        pb = st * (h + 2 * PAD)
        out = np.zeros(3 * pb, np.uint8)
        assert r.xref_frame_hpel(hnd, ptr(luma), w, ptr(out), None) == 0
        r.xref_close(hnd)
        planes = np.stack([out[i * pb:(i + 1) * pb].reshape(h + 2 * PAD, st)[:, :w + 2 * PAD] for i in range(3)])
    elif backend == "oracle":
        pl = make_ref_planes(luma)
        planes = np.stack([p.view()[:, :w + 2 * PAD] for p in pl[1:]])
    else:
        src = PaddedPlane(w, h)
        src.inner()[:] = luma
        st, nbytes, org = src.stride, src.buf.size, src.origin
        d = [ctx.upload(src.buf)] + [ctx.malloc(nbytes + 256) for _ in range(3)]
        ctx.check(ctx.L.x264cu_hpel_filter(ctx.h, d[0] + org, st, w, h, d[1] + org, d[2] + org, d[3] + org, 1))
        planes = np.stack([ctx.download(d[i], (h + 2 * PAD, st), np.uint8)[:, :w + 2 * PAD] for i in range(1, 4)])
        for p in d:
            ctx.free(p)
    return np.ascontiguousarray(planes), G.digest(luma)


# ---------------------------------------------------------------------------------------------------------------------
def run_me(backend, gi, ctx=None):
    """-> int32 [ME_JOBS, 4] = (mvx, mvy, cost, halfpel threshold after the call or -1)"""
    _libs._bind_me()
    method, subpel, me_range, satd, wt = G.ME_GROUPS[gi]
    fenc_l, ref_l = G.me_content(gi)
    planes = make_ref_planes(ref_l)
    st = planes[0].stride
    fenc = PaddedPlane(G.ME_W, G.ME_H, stride=st)
    fenc.inner()[:] = fenc_l
    if wt[0]:
        wplane = PaddedPlane(G.ME_W, G.ME_H, stride=st)
        ow = OrcWeight(*wt)
        oracle().orc_weight_scale_plane(ptr(wplane.buf), st, ptr(planes[0].buf), st, st, G.ME_H + 2 * PAD, C.byref(ow))
    else:
        wplane = planes[0]
    jobs = G.me_jobs(gi)
    dig = G.digest(fenc.buf, *[p.buf for p in planes], wplane.buf)
    out = np.zeros((len(jobs), 4), np.int32)
    n = 2 * 4 * G.ME_MV_RANGE
    if backend == "ref":
        r = ref()
        esa = method >= 3
        r.xref_me_search_frame.argtypes = [C.c_void_p, C.POINTER(XrefMeArgs), C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
        hnd = r.xref_open(G.ME_W, G.ME_H, b"medium", (b"subme=7" if satd else b"subme=1") + (b":me=%s:merange=32:partitions=all" % (b"tesa" if method == 4 else b"esa") if esa else b""), 0)
        assert hnd
        ref_l = np.ascontiguousarray(ref_l)
        for k, j in enumerate(jobs):
            off = planes[0].off(j["bx"], j["by"])
            a = XrefMeArgs()
            a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = j["ip"], method, subpel, me_range, 12
            for i in range(2):
                a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i] = j["lim_min"][i], j["lim_max"][i], j["mvp"][i]
            a.i_mvc = j["i_mvc"]
            for i in range(8):
                a.mvc[i][0], a.mvc[i][1] = int(j["mvc"][i][0]), int(j["mvc"][i][1])
            a.wt_en, a.wt_scale, a.wt_denom, a.wt_offset = wt
            a.use_thresh, a.halfpel_thresh = int(j["use_thresh"]), j["thresh"]
            if esa:     # the reference builds the frame (and its integral image) itself from the same luma
                assert r.xref_me_search_frame(hnd, C.byref(a), ptr(fenc.buf, fenc.off(j["bx"], j["by"])), st, ptr(ref_l), G.ME_W, j["bx"], j["by"]) == 0
            else:
                r.xref_me_search(hnd, C.byref(a), ptr(fenc.buf, fenc.off(j["bx"], j["by"])), st,
                                 *[ptr(p.buf, off) for p in planes], ptr(wplane.buf, off), st)
            out[k] = (a.mv[0], a.mv[1], a.cost, a.thresh_out if j["use_thresh"] else -1)
        r.xref_close(hnd)
    elif backend == "oracle":
        o = oracle()
        tab = np.zeros(2 * n + 1, np.uint16)
        o.orc_cost_mv_table(tab, n, 1)
        for k, j in enumerate(jobs):
            off = planes[0].off(j["bx"], j["by"])
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = method, subpel, me_range, satd
            for i in range(2):
                c.mv_min_spel[i], c.mv_max_spel[i] = j["lim_min"][i], j["lim_max"][i]
                c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = j["lim_min"][i] >> 2, j["lim_max"][i] >> 2
            m = OrcMe()
            m.i_pixel = j["ip"]
            m.p_cost_mv = tab.ctypes.data + 2 * n
            for i in range(4):
                m.p_fref[i] = planes[i].buf.ctypes.data + off
            m.p_fref_w = wplane.buf.ctypes.data + off
            m.p_fenc = fenc.buf.ctypes.data + fenc.off(j["bx"], j["by"])
            m.fenc_stride, m.stride = st, st
            m.weight = OrcWeight(*wt)
            m.mvp[0], m.mvp[1] = j["mvp"]
            mvc_arr = np.ascontiguousarray(j["mvc"])
            th = C.c_int(j["thresh"])
            o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), j["i_mvc"], C.byref(th) if j["use_thresh"] else None)
            out[k] = (m.mv[0], m.mv[1], m.cost, th.value if j["use_thresh"] else -1)
    else:
        import x264_b200 as x
        ja = np.zeros(len(jobs), x.me_job_dtype)
        for k, j in enumerate(jobs):
            e = ja[k]
            e["i_pixel"], e["fenc_off"], e["ref_off"] = j["ip"], fenc.off(j["bx"], j["by"]), planes[0].off(j["bx"], j["by"])
            e["mvp"], e["i_mvc"] = j["mvp"], j["i_mvc"]
            e["mvc"][:8] = j["mvc"]
            e["mv_min_spel"], e["mv_max_spel"] = j["lim_min"], j["lim_max"]
            e["halfpel_thresh"] = j["thresh"] if j["use_thresh"] else -1
        d_fenc = ctx.upload(fenc.buf)
        d_pl = [ctx.upload(p.buf) for p in planes]
        d_w = ctx.upload(wplane.buf) if wt[0] else d_pl[0]
        params = x.MeParams(method, subpel, me_range, satd, 1, G.ME_MV_RANGE, *wt)
        res = x.me_search_batch(ctx, params, d_fenc, st, d_pl, d_w, st, ja)
        for p in [d_fenc] + d_pl + ([d_w] if wt[0] else []):
            ctx.free(p)
        out[:, 0], out[:, 1], out[:, 2] = res["mv"][:, 0], res["mv"][:, 1], res["cost"]
        out[:, 3] = [int(res[k]["halfpel_thresh"]) if jobs[k]["use_thresh"] else -1 for k in range(len(jobs))]
    return out, dig


# ---------------------------------------------------------------------------------------------------------------------
def run_refine(backend, ci, ctx=None):
    """-> int32 [REFINE_JOBS, 4] = (mvx, mvy, cost, half-pel threshold after the call or -1) of x264_me_refine_qpel(_refdupe)"""
    _libs._bind_me()
    refdupe, satd, subpel, wt = G.REFINE_CASES[ci]
    fenc_l, ref_l, jobs = G.refine_case(ci)
    planes = make_ref_planes(np.ascontiguousarray(ref_l))
    st = planes[0].stride
    fenc = PaddedPlane(G.ME_W, G.ME_H, stride=st)
    fenc.inner()[:] = fenc_l
    dig = G.digest(fenc.buf, *[p.buf for p in planes])
    out = np.zeros((len(jobs), 4), np.int32)
    n = 2 * 4 * G.ME_MV_RANGE
    if backend == "ref":
        r = ref()
        r.xref_me_refine_qpel.argtypes = [C.c_void_p, C.POINTER(XrefMeArgs), C.c_int, C.c_int, C.c_void_p, C.c_ssize_t] + [C.c_void_p] * 4 + [C.c_ssize_t]
        r.xref_me_refine_qpel.restype = None
        hnd = r.xref_open(G.ME_W, G.ME_H, b"medium", b"subme=7" if satd else b"subme=1", 0)
        assert hnd
        for k, j in enumerate(jobs):
            off = planes[0].off(j["bx"], j["by"])
            a = XrefMeArgs()
            a.i_pixel, a.me_method, a.subpel_refine, a.me_range, a.qp = j["ip"], 1, subpel, 16, 12
            for i in range(2):
                a.mv_min_spel[i], a.mv_max_spel[i], a.mvp[i], a.mv[i] = j["lim_min"][i], j["lim_max"][i], j["mvp"][i], j["mv"][i]
            a.cost = j["cost"]
            a.wt_en, a.wt_scale, a.wt_denom, a.wt_offset = wt
            a.use_thresh, a.halfpel_thresh = int(j["use_thresh"]), j["thresh"]
            r.xref_me_refine_qpel(hnd, C.byref(a), refdupe, j["ref_cost"], ptr(fenc.buf, fenc.off(j["bx"], j["by"])), st, *[ptr(p.buf, off) for p in planes], st)
            out[k] = (a.mv[0], a.mv[1], a.cost, a.thresh_out if j["use_thresh"] else -1)
        r.xref_close(hnd)
    elif backend == "oracle":
        o = oracle()
        o.orc_me_refine_qpel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        o.orc_me_refine_qpel.restype = None
        tab = np.zeros(2 * n + 1, np.uint16)
        o.orc_cost_mv_table(tab, n, 1)
        for k, j in enumerate(jobs):
            off = planes[0].off(j["bx"], j["by"])
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, subpel, 16, satd
            for i in range(2):
                c.mv_min_spel[i], c.mv_max_spel[i] = j["lim_min"][i], j["lim_max"][i]
            m = OrcMe()
            m.i_pixel = j["ip"]
            m.p_cost_mv = tab.ctypes.data + 2 * n
            for i in range(4):
                m.p_fref[i] = planes[i].buf.ctypes.data + off
            m.p_fref_w = planes[0].buf.ctypes.data + off
            m.p_fenc = fenc.buf.ctypes.data + fenc.off(j["bx"], j["by"])
            m.fenc_stride, m.stride = st, st
            m.weight = OrcWeight(*wt)
            m.mvp[0], m.mvp[1], m.mv[0], m.mv[1], m.cost = j["mvp"][0], j["mvp"][1], j["mv"][0], j["mv"][1], j["cost"]
            th = C.c_int(j["thresh"])
            o.orc_me_refine_qpel(C.byref(c), C.byref(m), refdupe, j["ref_cost"], C.byref(th) if j["use_thresh"] else None)
            out[k] = (m.mv[0], m.mv[1], m.cost, th.value if j["use_thresh"] else -1)
    else:
        import x264_b200 as x
        ja = np.zeros(len(jobs), x.me_refine_job_dtype)
        for k, j in enumerate(jobs):
            e = ja[k]
            e["i_pixel"], e["fenc_off"], e["ref_off"] = j["ip"], fenc.off(j["bx"], j["by"]), planes[0].off(j["bx"], j["by"])
            e["mvp"], e["mv"], e["cost"], e["i_ref_cost"] = j["mvp"], j["mv"], j["cost"], j["ref_cost"]
            e["mv_min_spel"], e["mv_max_spel"] = j["lim_min"], j["lim_max"]
            e["halfpel_thresh"] = j["thresh"] if j["use_thresh"] else -1
        d_fenc = ctx.upload(fenc.buf)
        d_pl = [ctx.upload(p.buf) for p in planes]
        params = x.MeParams(1, subpel, 16, satd, 1, G.ME_MV_RANGE, *wt)
        res = x.me_refine_qpel_batch(ctx, params, refdupe, d_fenc, st, d_pl, st, ja)
        for p in [d_fenc] + d_pl:
            ctx.free(p)
        out[:, 0], out[:, 1], out[:, 2] = res["mv"][:, 0], res["mv"][:, 1], res["cost"]
        out[:, 3] = [int(res[k]["halfpel_thresh"]) if jobs[k]["use_thresh"] else -1 for k in range(len(jobs))]
    return out, dig


# ---------------------------------------------------------------------------------------------------------------------
def run_bidir(backend, ci, ctx=None):
    """-> int16 [BIDIR_JOBS, 4] = the refined (m0x, m0y, m1x, m1y) of x264_me_refine_bidir_satd"""
    _libs._bind_me()
    satd, _ = G.BIDIR_CASES[ci]
    fenc_l, ref0_l, ref1_l, jobs = G.bidir_case(ci)
    pl0, pl1 = make_ref_planes(np.ascontiguousarray(ref0_l)), make_ref_planes(np.ascontiguousarray(ref1_l))
    st = pl0[0].stride
    fenc = PaddedPlane(G.ME_W, G.ME_H, stride=st)
    fenc.inner()[:] = fenc_l
    dig = G.digest(fenc.buf, *[p.buf for p in pl0], *[p.buf for p in pl1])
    out = np.zeros((len(jobs), 4), np.int16)
    n = 2 * 4 * G.ME_MV_RANGE
    if backend == "ref":
        r = ref()
        r.xref_me_refine_bidir_satd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        r.xref_me_refine_bidir_satd.restype = None
        hnd = r.xref_open(G.ME_W, G.ME_H, b"medium", b"subme=7" if satd else b"subme=1", 0)
        assert hnd
        for k, j in enumerate(jobs):
            off = pl0[0].off(j["bx"], j["by"])
            f0 = (C.c_void_p * 4)(*[p.buf.ctypes.data + off for p in pl0])
            f1 = (C.c_void_p * 4)(*[p.buf.ctypes.data + off for p in pl1])
            mv = np.array(j["mv"], np.int16)
            mvp = np.array(j["mvp"], np.int16)
            a0, a1 = mv[:2].copy(), mv[2:].copy()
            p0, p1 = mvp[:2].copy(), mvp[2:].copy()
            lmin, lmax = np.array(j["lim_min"], np.int32), np.array(j["lim_max"], np.int32)
            r.xref_me_refine_bidir_satd(hnd, j["ip"], 12, ptr(fenc.buf, fenc.off(j["bx"], j["by"])), st, f0, f1, st, ptr(a0), ptr(p0), ptr(a1), ptr(p1),
                                        j["weight"], ptr(lmin), ptr(lmax))
            out[k] = (a0[0], a0[1], a1[0], a1[1])
        r.xref_close(hnd)
    elif backend == "oracle":
        o = oracle()
        o.orc_me_refine_bidir_satd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        o.orc_me_refine_bidir_satd.restype = None
        tab = np.zeros(2 * n + 1, np.uint16)
        o.orc_cost_mv_table(tab, n, 1)
        for k, j in enumerate(jobs):
            off = pl0[0].off(j["bx"], j["by"])
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, 7, 16, satd
            for i in range(2):
                c.mv_min_spel[i], c.mv_max_spel[i] = j["lim_min"][i], j["lim_max"][i]
            ms = []
            for li, pl in enumerate((pl0, pl1)):
                m = OrcMe()
                m.i_pixel = j["ip"]
                m.p_cost_mv = tab.ctypes.data + 2 * n
                for i in range(4):
                    m.p_fref[i] = pl[i].buf.ctypes.data + off
                m.p_fref_w = pl[0].buf.ctypes.data + off
                m.p_fenc = fenc.buf.ctypes.data + fenc.off(j["bx"], j["by"])
                m.fenc_stride, m.stride = st, st
                m.weight = OrcWeight(0, 0, 0, 0)
                m.mvp[0], m.mvp[1] = j["mvp"][2 * li], j["mvp"][2 * li + 1]
                m.mv[0], m.mv[1] = j["mv"][2 * li], j["mv"][2 * li + 1]
                ms.append(m)
            o.orc_me_refine_bidir_satd(C.byref(c), C.byref(ms[0]), C.byref(ms[1]), j["weight"])
            out[k] = (ms[0].mv[0], ms[0].mv[1], ms[1].mv[0], ms[1].mv[1])
    else:
        import x264_b200 as x
        ja = np.zeros(len(jobs), x.bidir_job_dtype)
        for k, j in enumerate(jobs):
            e = ja[k]
            off = pl0[0].off(j["bx"], j["by"])
            e["i_pixel"], e["fenc_off"], e["ref0_off"], e["ref1_off"] = j["ip"], fenc.off(j["bx"], j["by"]), off, off
            e["mv"], e["mvp"], e["mv_min_spel"], e["mv_max_spel"], e["i_weight"] = j["mv"], j["mvp"], j["lim_min"], j["lim_max"], j["weight"]
        d_fenc = ctx.upload(fenc.buf)
        d0, d1 = [ctx.upload(p.buf) for p in pl0], [ctx.upload(p.buf) for p in pl1]
        params = x.MeParams(1, 7, 16, satd, 1, G.ME_MV_RANGE, 0, 0, 0, 0)
        res = x.me_refine_bidir_batch(ctx, params, d_fenc, st, d0, d1, st, ja)
        for p in [d_fenc] + d0 + d1:
            ctx.free(p)
        out[:] = res["mv"]
    return out, dig


# ---------------------------------------------------------------------------------------------------------------------
def run_la(backend, ci, params_bytes=None, ctx=None):
    """-> dict of arrays: everything slicetype_frame_cost leaves behind for LA_REQUESTS (entries the reference never writes
    are zeroed on every side); params_bytes: the OrcLaParams of the case (produced by the ref run)"""
    _libs._bind_la()
    preset, opts, (w, h) = G.LA_CASES[ci]
    if backend == "ref":
        r = ref()
        hnd = r.xref_open(w, h, preset.encode(), opts.encode(), 0)
        assert hnd
        p = _libs.la_params_from_ref(hnd, w, h)
    else:
        p = OrcLaParams.from_buffer_copy(bytes(params_bytes))
    frames, qs = G.la_case(ci, p.weighted_pred)
    nfr, B, n = G.LA_NFR, p.bframes, p.mb_width * p.mb_height
    mask = np.ones((p.mb_height, p.mb_width), bool)
    if not (p.do_edges or p.mb_width <= 2 or p.mb_height <= 2):
        mask[:] = False
        mask[1:-1, 1:-1] = True
    mask = mask.reshape(-1)
    reqs = [q for q in G.LA_REQUESTS if q[1] < nfr and q[1] - q[0] <= B + 1]
    res = dict(scores=np.zeros(len(reqs), np.int32), weights=np.zeros((len(reqs), 4), np.int32),
               mvs=np.zeros((nfr, 2, B + 1, n, 2), np.int16), mv_costs=np.zeros((nfr, 2, B + 1, n), np.int32),
               intra=np.zeros((nfr, n), np.int32), cost_est=np.full((nfr, B + 2, B + 2, 2), -1, np.int32),
               intra_mbs=np.full((nfr, B + 2), -1, np.int32), lowres_costs=np.zeros((nfr, B + 2, B + 2, n), np.uint16))
    written = {i: set() for i in range(nfr)}
    nt = 2 * 4 * p.mv_range
    if backend == "ref":
        la = r.xref_la_new(hnd, nfr)
        for i, f in enumerate(frames):
            assert r.xref_la_set_frame(la, i, ptr(f), w, ptr(qs[i])) == 0
        cost = lambda p0, p1, b: r.xref_la_frame_cost(la, p0, p1, b)
        get = lambda idx, what, i, j, arr: r.xref_la_get(la, idx, what, i, j, ptr(arr))
    elif backend == "oracle":
        o = oracle()
        tab = np.zeros(2 * nt + 1, np.uint16)
        o.orc_cost_mv_table(tab, nt, 1)
        ofr = (C.c_void_p * (nfr + 2))()
        for i, f in enumerate(frames):
            ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
            o.orc_la_frame_set_qscale(ofr[i], qs[i])
        cost = lambda p0, p1, b: o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * nt, ofr, p0, p1, b)
        get = lambda idx, what, i, j, arr: o.orc_la_frame_get(ofr[idx], what, i, j, ptr(arr))
    else:
        import x264_b200 as x
        la = x.Lookahead(ctx, w, h, subpel_refine=p.subpel_refine, me_method=min(p.me_method, 2), me_range=p.me_range,
                         mv_range=p.mv_range, bframes=B, bframe_bias=p.bframe_bias, weighted_bipred=p.weighted_bipred,
                         aq_mode=p.aq_mode, mb_tree=int(p.do_edges and not p.vbv), vbv=p.vbv, n_slots=nfr,
                         weighted_pred=p.weighted_pred)
        for i, f in enumerate(frames):
            la.frame_put(i, f, qs[i])
        slots = list(range(nfr))
        cost = lambda p0, p1, b: la.frame_cost(slots, p0, p1, b)
    for k, (p0, p1, b) in enumerate(reqs):
        if not (p0 == p1 and written[b]):         # an I request after any other request is a memo hit: nothing is written
            written[b].add((b - p0, p1 - b))
        res["scores"][k] = cost(p0, p1, b)
        if p.weighted_pred and b == p1 and p0 != p1:
            if backend == "cuda":
                res["weights"][k] = la.get_weight(b)
            else:
                w4 = np.zeros(4, np.int32)
                get(b, 6, 0, 0, w4)
                res["weights"][k] = w4
            if not res["weights"][k][0]:
                res["weights"][k] = 0             # a disabled weight's other fields are not part of the contract
    for idx in range(nfr):
        for l in range(2 if B else 1):
            for d in range(B + 1):
                if backend == "cuda":
                    mv, co = la.get_mvs(idx, l, d)
                else:
                    mv = np.zeros((n, 2), np.int16)
                    co = np.zeros(n, np.int32)
                    get(idx, 0, l, d, mv)
                    if mv[0, 0] != 0x7FFF:
                        get(idx, 1, l, d, co)
                if mv[0, 0] == 0x7FFF:
                    res["mvs"][idx, l, d, 0, 0] = 0x7FFF          # "never searched": only the sentinel is defined
                else:
                    res["mvs"][idx, l, d] = mv
                    res["mv_costs"][idx, l, d][mask] = co[mask]
        for (i, j) in sorted(written[idx]):
            if backend == "cuda":
                ce, cea, imb = la.get_cost_est(idx, i, j)
                lc = la.get_costs(idx, i, j)
            else:
                e = np.zeros(3, np.int32)
                get(idx, 4, i, j, e)
                ce, cea, imb = int(e[0]), int(e[1]), int(e[2])
                lc = np.zeros(n, np.uint16)
                get(idx, 2, i, j, lc)
            res["cost_est"][idx, i, j] = (ce, cea)
            if j == 0:
                res["intra_mbs"][idx, i] = imb
            res["lowres_costs"][idx, i, j][mask] = lc[mask]
        if written[idx]:
            if backend == "cuda":
                ic = la.get_intra(idx)
            else:
                ic = np.zeros(n, np.int32)
                get(idx, 3, 0, 0, ic)
            res["intra"][idx][mask] = ic[mask]
    if backend == "ref":
        r.xref_la_free(la)
        r.xref_close(hnd)
    elif backend == "oracle":
        for i in range(nfr):
            oracle().orc_la_frame_delete(ofr[i])
    else:
        la.close()
    dig = G.digest(*frames, *qs)
    return res, np.frombuffer(bytes(p), np.uint8).copy(), dig


# ---------------------------------------------------------------------------------------------------------------------
def run_st(backend, ci, params_bytes=None, ctx=None):
    """-> int32 [n, 2] = (display index, X264_TYPE_*) in coded order; params_bytes: SlicetypeParams of the case"""
    import test_slicetype_host as host
    from x264_b200.binding_ext import SlicetypeParams
    preset, opts, (w, h), n, cut = G.ST_CASES[ci]
    frames = G.st_case(ci)
    dig = G.digest(*frames)
    if backend == "ref":
        p, types = host.reference_types(preset, opts, w, h, frames)
    else:
        p = SlicetypeParams.from_buffer_copy(bytes(params_bytes))
        # the fixture was written when 0 meant "the reference's default" for these two; a negative value says that now
        if p.qcompress == 0:
            p.qcompress = -1.0
        if p.aq_strength == 0:
            p.aq_strength = -1.0
        if backend == "oracle":
            types = host.decide_with(_libs.slicetype_oracle_lib(), p, frames)
        else:
            import x264_b200 as x
            st = x.Slicetype.from_params(ctx, p)
            try:
                types = st.decide(frames)
            finally:
                st.close()
    return np.array(types, np.int32).reshape(-1, 2), np.frombuffer(bytes(p), np.uint8).copy(), dig


# ---------------------------------------------------------------------------------------------------------------------
def run_aq(backend, ci, ctx=None):
    """-> (f_qp_offset_aq float32[mb], i_inv_qscale_factor u16[mb], stats u64[6])"""
    (w, h), mode, strength = G.AQ_CASES[ci]
    luma, cb, cr = G.aq_case(ci)
    nmb = ((w + 15) // 16) * ((h + 15) // 16)
    q, iq, st = np.zeros(nmb, np.float32), np.zeros(nmb, np.uint16), np.zeros(6, np.uint64)
    if backend == "ref":
        r = ref()
        r.xref_aq_frame.argtypes = [C.c_void_p] * 7
        hnd = r.xref_open(w, h, b"medium", ("aq-mode=%d:aq-strength=%g" % (mode, strength)).encode(), 0)
        assert hnd and r.xref_aq_frame(hnd, ptr(luma), ptr(cb), ptr(cr), ptr(q), ptr(iq), ptr(st)) == 0
        r.xref_close(hnd)
    elif backend == "oracle":
        o = oracle()
        o.orc_adaptive_quant_frame.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
        o.orc_adaptive_quant_frame(ptr(luma), w, ptr(cb), ptr(cr), cb.shape[1], w, h, mode, strength, ptr(q), ptr(iq), ptr(st))
    else:
        ctx.L.x264cu_adaptive_quant_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int,
                                                      C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        d_l, d_b, d_r = ctx.upload(luma), ctx.upload(cb), ctx.upload(cr)
        d_q, d_i = ctx.malloc(nmb * 4), ctx.malloc(nmb * 2)
        ctx.check(ctx.L.x264cu_adaptive_quant_frame(ctx.h, d_l, w, d_b, d_r, cb.shape[1], w, h, mode, strength, d_q, d_i, st.ctypes.data))
        q, iq = ctx.download(d_q, (nmb,), np.float32), ctx.download(d_i, (nmb,), np.uint16)
        for p in (d_l, d_b, d_r, d_q, d_i):
            ctx.free(p)
    return (q, iq, st), G.digest(luma, cb, cr)


def run_mbtree(backend, params_bytes=None, ctx=None):
    """-> (types int32[n,2], qp float32[k, mb] for the non-B pictures in coded order, SlicetypeParams bytes, digest)"""
    import test_slicetype_host as host
    from x264_b200.binding_ext import SlicetypeParams
    preset, opts, (w, h), n, cut = G.MBTREE_CASE
    frames = G.mbtree_case()
    qp = {}
    if backend == "ref":
        p, types = host.reference_types(preset, opts, w, h, frames, qp)
    else:
        p = SlicetypeParams.from_buffer_copy(bytes(params_bytes))
        # the fixture was written when 0 meant "the reference's default" for these two; a negative value says that now
        if p.qcompress == 0:
            p.qcompress = -1.0
        if p.aq_strength == 0:
            p.aq_strength = -1.0
        if backend == "oracle":
            types = host.decide_with(_libs.slicetype_oracle_lib(), p, frames, qp)
        else:
            import x264_b200 as x
            st = x.Slicetype.from_params(ctx, p)
            try:
                types = st.decide(frames, qp)
            finally:
                st.close()
    nonb = [f for f, t in types if t not in (4, 5)]
    qarr = np.stack([np.asarray(qp[f], np.float32) for f in nonb])
    return np.array(types, np.int32).reshape(-1, 2), qarr, np.frombuffer(bytes(p), np.uint8).copy(), G.digest(*frames)
