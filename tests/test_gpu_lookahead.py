"""GPU parity of the lowres lookahead (x264cu_lookahead_frame_cost == slicetype_frame_cost, encoder/slicetype.c:836-995)
against the oracle, and against the compiled reference where it travelled with the snapshot.  Every per-MB output and
every frame-level sum, bit-exact, for I / P / B requests in an order that exercises the memo and the availability of the
temporal-direct vectors (slicetype.c:629-642)."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ref, have_ref, ptr, OrcLaParams, synth_sequence

pytestmark = pytest.mark.gpu

REQUESTS = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (0, 2, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 3, 2), (1, 2, 2),
            (2, 4, 3), (2, 4, 4), (1, 4, 4), (1, 4, 2), (1, 4, 3), (4, 4, 4), (3, 5, 4), (3, 5, 5), (2, 5, 5)]

# (width, height, subme, me, merange, bframes, weightb, aq, mbtree, vbv)
CONFIGS = [
    (112, 80, 7, 1, 16, 3, 1, 1, 1, 0),
    (112, 80, 1, 1, 16, 3, 1, 1, 0, 0),
    (96, 96, 7, 2, 24, 4, 1, 0, 1, 0),
    (100, 60, 7, 0, 16, 3, 0, 1, 0, 0),
    (64, 48, 0, 0, 16, 2, 1, 0, 0, 0),
    (80, 64, 7, 1, 16, 3, 1, 1, 1, 1),
    (352, 288, 7, 1, 16, 3, 1, 1, 1, 0),
    # an 11th field = weighted_pred: the lookahead weight analysis (slicetype.c:284-501) on a fade
    (112, 80, 7, 1, 16, 3, 1, 1, 1, 0, 1),
    (96, 64, 2, 1, 16, 2, 1, 0, 0, 0, 1),
    (176, 144, 7, 2, 16, 3, 1, 1, 1, 0, 1),
]


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_la()
    c = x.Context(0)
    yield c
    c.close()


def oracle_params(cfg, mv_range=512):
    w, h, subme, me, merange, bframes, weightb, aq, mbtree, vbv = cfg[:10]
    p = OrcLaParams()
    p.weighted_pred = cfg[10] if len(cfg) > 10 else 0
    p.width, p.height = w, h
    p.mb_width, p.mb_height = (w + 15) // 16, (h + 15) // 16
    p.subpel_refine, p.me_method, p.me_range, p.mv_range = subme, me, merange, mv_range
    p.bframes, p.bframe_bias, p.weighted_bipred = bframes, 0, weightb
    p.aq_mode, p.vbv, p.do_edges = aq, vbv, int(mbtree or vbv)
    return p


@pytest.mark.parametrize("cfg", CONFIGS)
def test_frame_cost_matches_oracle(ctx, cfg):
    w, h, subme, me, merange, bframes, weightb, aq, mbtree, vbv = cfg[:10]
    o = oracle()
    p = oracle_params(cfg)
    nfr = 6
    frames = synth_sequence(w, h, nfr, seed=w + h, cut_at=4)
    if p.weighted_pred:
        frames = [np.clip(f.astype(np.float32) * (0.55 + 0.09 * i) + 3 * i, 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
    weights_seen = []
    n = 2 * 4 * p.mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, 1)
    rng = np.random.default_rng(1)
    la = x.Lookahead(ctx, w, h, subpel_refine=subme, me_method=me, me_range=merange, mv_range=p.mv_range, bframes=bframes,
                     weighted_bipred=weightb, aq_mode=aq, mb_tree=mbtree, vbv=vbv, n_slots=nfr,
                     weighted_pred=p.weighted_pred)
    ofr = (C.c_void_p * (nfr + 2))()
    try:
        for i, f in enumerate(frames):
            q = rng.integers(180, 400, p.mb_width * p.mb_height).astype(np.uint16)
            la.frame_put(i, f, q)
            ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
            o.orc_la_frame_set_qscale(ofr[i], q)
            # lowres planes first: everything else depends on them
            for pl in range(4):
                got = la.get_lowres_plane(i, pl)
                st = o.orc_la_frame_stride(ofr[i])
                rows = p.mb_height * 8 + 64
                want = np.ctypeslib.as_array((C.c_uint8 * (rows * st)).from_address(o.orc_la_frame_plane(ofr[i], pl))).reshape(rows, st)
                wl = p.mb_width * 8 + 64
                assert got.shape[1] == st and np.array_equal(got[:, :wl], want[:, :wl]), ("lowres", i, pl)
        reqs = [r for r in REQUESTS if r[1] < nfr and r[1] - r[0] <= bframes + 1]
        nmb = p.mb_width * p.mb_height
        mask = np.ones((p.mb_height, p.mb_width), bool)
        if not p.do_edges:
            mask[:] = False
            mask[1:-1, 1:-1] = True
        mask = mask.reshape(-1)
        slots = list(range(nfr))
        for (p0, p1, b) in reqs:
            s_gpu = la.frame_cost(slots, p0, p1, b)
            s_orc = o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * n, ofr, p0, p1, b)
            tag = (cfg, p0, p1, b)
            if p.weighted_pred and b == p1 and p0 != p1:
                w2 = np.zeros(4, np.int32)
                o.orc_la_frame_get(ofr[b], 6, 0, 0, ptr(w2))
                assert la.get_weight(b) == tuple(w2), (tag, "weights", la.get_weight(b), w2)
                if w2[0]:
                    weights_seen.append(tuple(w2))
            # vectors / vector costs of every searched list
            for l in range(2 if bframes else 1):
                for d in range(bframes + 1):
                    a = np.zeros((nmb, 2), np.int16)
                    o.orc_la_frame_get(ofr[b], 0, l, d, ptr(a))
                    mv, co = la.get_mvs(b, l, d)
                    assert np.array_equal(mv, a), (tag, "mvs", l, d, np.argwhere(mv != a)[:5], mv[mv != a][:6], a[mv != a][:6])
                    if a[0, 0] != 0x7FFF:
                        c2 = np.zeros(nmb, np.int32)
                        o.orc_la_frame_get(ofr[b], 1, l, d, ptr(c2))
                        assert np.array_equal(co[mask], c2[mask]), (tag, "mv_costs", l, d)
            a = np.zeros(nmb, np.int32)
            o.orc_la_frame_get(ofr[b], 3, 0, 0, ptr(a))
            assert np.array_equal(la.get_intra(b)[mask], a[mask]), (tag, "intra")
            e = np.zeros(3, np.int32)
            o.orc_la_frame_get(ofr[b], 4, b - p0, p1 - b, ptr(e))
            ce, cea, imb = la.get_cost_est(b, b - p0, p1 - b)
            assert (ce, cea) == (e[0], e[1]), (tag, "cost_est", (ce, cea), e)
            if b == p1:
                assert imb == e[2], (tag, "intra_mbs")
            o.orc_la_frame_get(ofr[b], 4, 0, 0, ptr(e))
            ce, cea, _ = la.get_cost_est(b, 0, 0)
            assert (ce, cea) == (e[0], e[1]), (tag, "intra cost_est")
            c2 = np.zeros(nmb, np.uint16)
            o.orc_la_frame_get(ofr[b], 2, b - p0, p1 - b, ptr(c2))
            got = la.get_costs(b, b - p0, p1 - b)
            if not (p0 == p1 and (b - p0, p1 - b) == (0, 0) and s_gpu == s_orc and c2.max() == 0):
                assert np.array_equal(got[mask], c2[mask]), (tag, "lowres_costs", np.argwhere(got != c2)[:5])
            if vbv:
                r2 = np.zeros(p.mb_height, np.int32)
                o.orc_la_frame_get(ofr[b], 5, b - p0, p1 - b, ptr(r2))
                assert np.array_equal(la.get_row_satds(b, b - p0, p1 - b), r2), (tag, "row_satds")
            assert s_gpu == s_orc, (tag, "score", s_gpu, s_orc)
        if p.weighted_pred:
            assert weights_seen, "the fade should have produced at least one weighted P search"
    finally:
        for i in range(nfr):
            if ofr[i]:
                o.orc_la_frame_delete(ofr[i])
        la.close()


def test_search_batch_equals_on_demand(ctx):
    """x264cu_lookahead_search_batch (the slicetype_prep analogue) must not change any frame_cost result"""
    cfg = (176, 144, 7, 1, 16, 3, 1, 0, 1, 0)
    w, h = cfg[0], cfg[1]
    frames = synth_sequence(w, h, 6, seed=77)
    res = []
    for batch in (False, True):
        la = x.Lookahead(ctx, w, h, bframes=3, aq_mode=0, n_slots=6)
        for i, f in enumerate(frames):
            la.frame_put(i, f)
        if batch:
            jobs = [(b, b - d, 0, d) for b in range(6) for d in range(1, 5) if b - d >= 0]
            jobs += [(b, b + d, 1, d) for b in range(6) for d in range(1, 5) if b + d < 6]
            la.search_batch(jobs)
        out = [la.frame_cost(list(range(6)), p0, p1, b) for (p0, p1, b) in REQUESTS if p1 < 6]
        res.append(out)
        la.close()
    assert res[0] == res[1]
