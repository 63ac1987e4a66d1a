"""GPU parity of x264cu_me_search_batch (the batched x264_me_search_ref twin) against the oracle, which tests/test_oracle_me.py
pins to the compiled reference: every partition size, DIA / HEX / UMH, every sub-pel level, weighted references, half-pel
thresholds, random predictors and windows, on textured, flat (tie-heavy) and noisy content.  (mv, cost) bit-exact."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ptr, PaddedPlane, OrcWeight, OrcMeCtx, OrcMe, make_ref_planes, PIXEL_W, PIXEL_H
from test_oracle_me import _content, chroma_pair, W, H

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
        i_mvc = int(rng.integers(0, 10))
        spread = int(rng.choice([2, 12, 50]))
        mvp = rng.integers(-spread, spread + 1, 2)
        mvcs = rng.integers(-spread, spread + 1, (9, 2))
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
@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_me_search_batch_matches_oracle(ctx, kind, method):
    """method 3 / 4 = ESA / TESA (me.c:618-771): with a weighted reference the ADS prefilter (sums of the unweighted plane) decides
    which positions are measured; under TESA fpelcmp is SATD when mbcmp is"""
    rng = np.random.default_rng(17 * method + len(kind))
    for subpel in (0, 1, 2, 3, 4, 5, 6, 7, 9):
        me_range = int(rng.choice([16, 24, 32] if method == 2 else [4, 8, 16, 24] if method == 4 else [4, 8, 16]))
        satd = int(subpel > 1 and rng.random() < 0.8)
        wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.3 else (0, 0, 0, 0)
        run_group(ctx, kind, method, subpel, me_range, satd, wt, rng)


@pytest.mark.parametrize("kind", ["texture", "flat", "noise"])
@pytest.mark.parametrize("method", [1, 2])
def test_me_search_frame_chroma_multiref_matches_oracle(ctx, kind, method):
    """x264cu_me_search_frame: one launch over jobs that name their reference picture (three of them, weighted luma / chroma
    planes) and their lambda, with chroma ME (me.c:826-857, mc.c:251-283) -- against the oracle's chroma ME, which
    tests/test_oracle_me.py pins to the compiled reference"""
    o = oracle()
    rng = np.random.default_rng(911 * method + len(kind))
    mv_range = 64
    n = 2 * 4 * mv_range
    lambdas = [1, 3, 7, 12]
    tabs = []
    for lam in lambdas:
        t = np.zeros(2 * n + 1, np.uint16)
        o.orc_cost_mv_table(t, n, lam)
        tabs.append(t)
    for subpel in (5, 7, 9, 4):
        me_range = int(rng.choice([16, 24] if method == 2 else [8, 16]))
        fenc_l, _ = _content(kind, rng)
        st = PaddedPlane(W, H).stride
        fenc = PaddedPlane(W, H, stride=st)
        fenc.inner()[:] = fenc_l
        refs = []
        for _ in range(3):
            _, ref_l = _content(kind, rng)
            planes = make_ref_planes(ref_l, stride=st)
            wt = [(1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.4 else (0, 0, 0, 0)] + \
                 [(1, int(rng.integers(40, 90)), int(rng.integers(0, 7)), int(rng.integers(-6, 7))) if rng.random() < 0.4 else (0, 0, 0, 0)
                  for _ in range(2)]
            if wt[0][0]:
                wplane = PaddedPlane(W, H, stride=st)
                ow = OrcWeight(*wt[0])
                o.orc_weight_scale_plane(ptr(wplane.buf), st, ptr(planes[0].buf), st, st, H + 64, C.byref(ow))
            else:
                wplane = planes[0]
            fc, rc = chroma_pair(kind, rng, st)
            refs.append((planes, wplane, rc, wt))
        fenc_c, _ = chroma_pair(kind, rng, st)
        n_jobs = 160
        jobs = np.zeros(n_jobs, x.me_frame_job_dtype)
        want = []
        for k in range(n_jobs):
            ip = int(rng.integers(0, 4 if rng.random() < 0.8 else 7))
            bw, bh = PIXEL_W[ip], PIXEL_H[ip]
            bx = int(rng.integers(0, (W - bw) // 8 + 1)) * 8 if ip < 4 else int(rng.integers(0, (W - bw) // 4 + 1)) * 4
            by = int(rng.integers(0, (H - bh) // 8 + 1)) * 8 if ip < 4 else int(rng.integers(0, (H - bh) // 4 + 1)) * 4
            mvr = 4 * mv_range
            lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
            lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
            i_mvc = int(rng.integers(0, 10))
            spread = int(rng.choice([2, 12, 50]))
            mvp = rng.integers(-spread, spread + 1, 2)
            mvcs = rng.integers(-spread, spread + 1, (9, 2))
            use_thresh = rng.random() < 0.2
            thresh = int(rng.integers(50, 3000)) if use_thresh else -1
            i_ref, i_lam = int(rng.integers(0, 3)), int(rng.integers(0, len(lambdas)))
            planes, wplane, rc, wt = refs[i_ref]
            j = jobs[k]
            j["job"]["i_pixel"], j["job"]["fenc_off"], j["job"]["ref_off"] = ip, by * st + bx, by * st + bx
            j["job"]["mvp"], j["job"]["mvc"], j["job"]["i_mvc"] = mvp, mvcs, i_mvc
            j["job"]["mv_min_spel"], j["job"]["mv_max_spel"], j["job"]["halfpel_thresh"] = lim_min, lim_max, thresh
            j["i_ref"], j["i_lambda"] = i_ref, i_lam
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd, c.chroma_me = method, subpel, me_range, 1, 1
            for i in range(2):
                c.mv_min_spel[i], c.mv_max_spel[i] = lim_min[i], lim_max[i]
                c.mv_limit_fpel[0][i], c.mv_limit_fpel[1][i] = lim_min[i] >> 2, lim_max[i] >> 2
            m = OrcMe()
            m.i_pixel = ip
            m.p_cost_mv = tabs[i_lam].ctypes.data + 2 * n
            off = planes[0].off(bx, by)
            for i in range(4):
                m.p_fref[i] = planes[i].buf.ctypes.data + off
            m.p_fref_w = wplane.buf.ctypes.data + off
            m.p_fenc = fenc.buf.ctypes.data + fenc.off(bx, by)
            m.fenc_stride, m.stride = st, st
            m.weight = OrcWeight(*wt[0])
            m.mvp[0], m.mvp[1] = int(mvp[0]), int(mvp[1])
            m.p_fref_uv, m.stride_uv = rc.buf.ctypes.data + rc.off(bx & ~1, by // 2), st
            m.p_fenc_uv, m.fenc_uv_stride = fenc_c.buf.ctypes.data + fenc_c.off(bx & ~1, by // 2), st
            m.weight_uv[0], m.weight_uv[1] = OrcWeight(*wt[1]), OrcWeight(*wt[2])
            mvc_arr = np.ascontiguousarray(mvcs.astype(np.int16))
            th = C.c_int(thresh)
            o.orc_me_search_ref(C.byref(c), C.byref(m), ptr(mvc_arr), i_mvc, C.byref(th) if use_thresh else None)
            want.append((m.mv[0], m.mv[1], m.cost, th.value if use_thresh else -1))
        live = []
        def up(pl):
            d = ctx.upload(pl.buf)
            live.append(d)
            return d + pl.origin
        d_fenc, d_fenc_c = up(fenc), up(fenc_c)
        dev_refs = []
        for planes, wplane, rc, wt in refs:
            d_pl = [up(p) for p in planes]
            dev_refs.append((d_pl, up(wplane) if wt[0][0] else None, up(rc), wt))
        frame, keep = x.make_me_frame(d_fenc, st, d_fenc_c, st, st, st, dev_refs, lambdas, 1)
        params = x.MeParams(method, subpel, me_range, 1, 1, mv_range, 0, 0, 0, 0)
        res = x.me_search_frame(ctx, params, frame, jobs)
        for d in live:
            ctx.free(d)
        for k in range(n_jobs):
            got = (int(res[k]["mv"][0]), int(res[k]["mv"][1]), int(res[k]["cost"]), int(res[k]["halfpel_thresh"]))
            assert got == want[k], (kind, method, subpel, k, jobs[k], got, want[k])


def test_me_search_batch_tesa_many_jobs(ctx):
    """more TESA jobs than warps in flight share the bounded candidate-list scratch (each warp walks several jobs)"""
    rng = np.random.default_rng(5)
    run_group(ctx, "flat", 4, 7, 16, 1, (0, 0, 0, 0), rng, n_jobs=700)
    run_group(ctx, "texture", 4, 2, 32, 1, (1, 70, 6, -2), rng, n_jobs=64)


def test_me_search_batch_large_batches_are_walked_by_partition_size(ctx):
    """from 4 096 jobs on the kernel walks the jobs in the order of a device counting sort by i_pixel (me.cu: me_order_*_kernel;
    one instruction stream per SM at a time): every result must still land in its caller's slot -- mixed sizes, two chunk
    boundaries (1 024 jobs per chunk), a ragged tail"""
    rng = np.random.default_rng(4096)
    run_group(ctx, "texture", 1, 7, 16, 1, (0, 0, 0, 0), rng, n_jobs=4096 + 1024 + 37)
    run_group(ctx, "noise", 2, 5, 24, 1, (1, 70, 6, -2), rng, n_jobs=4100)


@pytest.mark.parametrize("kind", ["texture", "flat"])
@pytest.mark.parametrize("satd", [0, 1])
def test_me_refine_bidir_batch_matches_oracle(ctx, kind, satd):
    """x264cu_me_refine_bidir_batch against the oracle's x264_me_refine_bidir_satd (pinned to the reference in
    tests/test_oracle_me.py): every partition size, bipred weights 32 / 21 / 43 / -10 / 64, pairs near the window edge"""
    o = oracle()
    o.orc_me_refine_bidir_satd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    o.orc_me_refine_bidir_satd.restype = None
    rng = np.random.default_rng(3 + satd + len(kind))
    fenc_l, ref0_l = _content(kind, rng)
    _, ref1_l = _content(kind, rng)
    if kind == "texture":
        ref1_l = np.ascontiguousarray(np.roll(ref0_l, (3, -2), (0, 1)))
    pl0, pl1 = make_ref_planes(np.ascontiguousarray(ref0_l)), make_ref_planes(np.ascontiguousarray(ref1_l))
    st = pl0[0].stride
    fenc = PaddedPlane(W, H, stride=st)
    fenc.inner()[:] = fenc_l
    mv_range, lam = 64, 2
    n = 2 * 4 * mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, lam)
    n_jobs = 300
    jobs = np.zeros(n_jobs, x.bidir_job_dtype)
    want = []
    for k in range(n_jobs):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
        mvr = 4 * mv_range
        lim_min = np.array([max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)])
        lim_max = np.array([min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)])
        spread = int(rng.choice([6, 30, 120]))
        mv0 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
        mv1 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
        mvp0, mvp1 = mv0 + rng.integers(-6, 7, 2), mv1 + rng.integers(-6, 7, 2)
        weight = int(rng.choice([32, 32, 21, 43, -10, 64]))
        off = pl0[0].off(bx, by)
        j = jobs[k]
        j["i_pixel"], j["fenc_off"], j["ref0_off"], j["ref1_off"] = ip, fenc.off(bx, by), off, off
        j["mv"], j["mvp"] = np.concatenate([mv0, mv1]), np.concatenate([mvp0, mvp1])
        j["mv_min_spel"], j["mv_max_spel"], j["i_weight"] = lim_min, lim_max, weight
        c = OrcMeCtx()
        c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, 7, 16, satd
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
        want.append((ms[0].mv[0], ms[0].mv[1], ms[1].mv[0], ms[1].mv[1], ms[0].cost))
    d_fenc = ctx.upload(fenc.buf)
    d0, d1 = [ctx.upload(p.buf) for p in pl0], [ctx.upload(p.buf) for p in pl1]
    params = x.MeParams(1, 7, 16, satd, lam, mv_range, 0, 0, 0, 0)
    res = x.me_refine_bidir_batch(ctx, params, d_fenc, st, d0, d1, st, jobs)
    for p in [d_fenc] + d0 + d1:
        ctx.free(p)
    moved = 0
    for k in range(n_jobs):
        got = tuple(int(v) for v in res[k]["mv"]) + (int(res[k]["cost"]),)
        assert got == want[k], (kind, satd, k, jobs[k], got, want[k])
        moved += got[:4] != tuple(int(v) for v in jobs[k]["mv"])
    assert moved > 50


@pytest.mark.parametrize("kind", ["texture", "flat"])
@pytest.mark.parametrize("refdupe", [0, 1])
def test_me_refine_qpel_batch_matches_oracle(ctx, kind, refdupe):
    """x264cu_me_refine_qpel_batch against the oracle's x264_me_refine_qpel / _refdupe (pinned to the reference in
    tests/test_oracle_me.py): every sub-pel level incl. the simplified level 1, weights, half-pel thresholds"""
    o = oracle()
    o.orc_me_refine_qpel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    o.orc_me_refine_qpel.restype = None
    rng = np.random.default_rng(11 + refdupe + len(kind))
    mv_range, lam = 64, 1
    n = 2 * 4 * mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, lam)
    moved = 0
    for subpel in (1, 2, 3, 4, 5, 6, 7, 9):
        satd = int(subpel > 1 and rng.random() < 0.8)
        wt = (1, int(rng.integers(40, 90)), 6, int(rng.integers(-4, 5))) if rng.random() < 0.3 else (0, 0, 0, 0)
        fenc_l, ref_l = _content(kind, rng)
        planes = make_ref_planes(ref_l)
        st = planes[0].stride
        fenc = PaddedPlane(W, H, stride=st)
        fenc.inner()[:] = fenc_l
        n_jobs = 64
        jobs = np.zeros(n_jobs, x.me_refine_job_dtype)
        want = []
        for k in range(n_jobs):
            ip = int(rng.integers(0, 7))
            bw, bh = PIXEL_W[ip], PIXEL_H[ip]
            bx = int(rng.integers(0, (W - bw) // 4 + 1)) * 4
            by = int(rng.integers(0, (H - bh) // 4 + 1)) * 4
            mvr = 4 * mv_range
            lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
            lim_max = [min(4 * (W - bx - bw + 24), mvr - 1), min(4 * (H - by - bh + 24), mvr - 1)]
            spread = int(rng.choice([4, 20, 90]))
            mv = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
            mvp = mv + rng.integers(-6, 7, 2)
            cost, ref_cost = int(rng.integers(50, 6000)), int(rng.integers(0, 5))
            use_thresh = bool(refdupe and rng.random() < 0.4)
            thresh = int(rng.integers(50, 6000)) if use_thresh else -1
            off = planes[0].off(bx, by)
            j = jobs[k]
            j["i_pixel"], j["fenc_off"], j["ref_off"], j["mvp"], j["mv"] = ip, fenc.off(bx, by), off, mvp, mv
            j["cost"], j["i_ref_cost"], j["mv_min_spel"], j["mv_max_spel"], j["halfpel_thresh"] = cost, ref_cost, lim_min, lim_max, thresh
            c = OrcMeCtx()
            c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd = 1, subpel, 16, satd
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
            o.orc_me_refine_qpel(C.byref(c), C.byref(m), refdupe, ref_cost, C.byref(th) if use_thresh else None)
            want.append((m.mv[0], m.mv[1], m.cost, th.value if use_thresh else -1))
            moved += (m.mv[0], m.mv[1]) != (int(mv[0]), int(mv[1]))
        d_fenc = ctx.upload(fenc.buf)
        d_pl = [ctx.upload(p.buf) for p in planes]
        params = x.MeParams(1, subpel, 16, satd, lam, mv_range, *wt)
        res = x.me_refine_qpel_batch(ctx, params, refdupe, d_fenc, st, d_pl, st, jobs)
        for p in [d_fenc] + d_pl:
            ctx.free(p)
        for k in range(n_jobs):
            got = (int(res[k]["mv"][0]), int(res[k]["mv"][1]), int(res[k]["cost"]), int(res[k]["halfpel_thresh"]))
            assert got == want[k], (kind, refdupe, subpel, satd, wt, k, jobs[k], got, want[k])
    assert moved > 30


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
    jobs["mvc"][:, :8] = rng.integers(-30, 31, (mbw * mbh, 8, 2))
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
