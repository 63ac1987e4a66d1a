"""The BASELINE.json configurations as GPU parity cases (sizes the CPU checkers finish in seconds to a minute):
  configs[1]  1920x1080, preset medium, rc-lookahead 40: slice-type decisions of the CUDA lookahead == the same host logic over
              the CPU oracle (which tests/test_slicetype_host.py pins to the reference encoder), MB-tree offsets included
  configs[3]  7680x4320, rc-lookahead 250, bframes 16, b-adapt 2: at full size the oracle is too slow for a whole sequence, so
              (a) a handful of slicetype_frame_cost requests against the oracle, every per-MB array, and (b) the size-independent
              property that prefetching / run-ahead / on-demand scheduling give identical decisions on the device
  odd sizes   width / height not multiples of 16 (1080 = 67.5 MB rows) and the smallest pictures the reference supports"""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ptr, OrcLaParams, synth_sequence, slicetype_oracle_lib
import test_slicetype_host as host
from x264_b200.binding_ext import SlicetypeParams, LookaheadParams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_la()
    c = x.Context(0)
    yield c
    c.close()


def big_sequence(w, h, n, seed, cut_at=None):
    """cheap large synthetic sequence: a low-pass texture translated by a per-picture global motion, +-2 noise, optional cut"""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (h // 8 + 40, w // 8 + 40)).astype(np.float32)
    k = np.ones(5, np.float32) / 5
    small = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 0, small)
    small = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 1, small)
    master = np.kron(small, np.ones((8, 8), np.float32))
    master2 = master[::-1, ::-1].copy()
    out = []
    px = py = 100
    for i in range(n):
        m = master2 if cut_at is not None and i >= cut_at else master
        px = int(np.clip(px + rng.integers(-5, 6), 0, 300))
        py = int(np.clip(py + rng.integers(-3, 4), 0, 300))
        out.append(np.clip(m[py:py + h, px:px + w] + rng.integers(-2, 3, (h, w)), 0, 255).astype(np.uint8))
    return out


def slicetype_params(w, h, bframes=3, b_adapt=1, rc_lookahead=40, mb_tree=1, weighted_pred=0, aq=0, keyint_max=250):
    la = LookaheadParams(w, h, 7, 1, 16, 512, bframes, 0, 1, aq, mb_tree, 0, 0, weighted_pred)
    return SlicetypeParams(la, keyint_max, 25, 40, b_adapt, 2, rc_lookahead, 0, 3, 0)


def gpu_decide(ctx, p, frames, qp=None, prefetch=None, run_ahead=None):
    st = x.Slicetype.from_params(ctx, p)
    try:
        if prefetch is not None:
            st.set_prefetch(prefetch)
        if run_ahead is not None:
            st.set_run_ahead(run_ahead)
        return st.decide(frames, qp)
    finally:
        st.close()


@pytest.mark.parametrize("weighted_pred", [0, 1])
def test_config1_1080p_medium_lookahead40(ctx, weighted_pred):
    w, h, n = 1920, 1080, 64
    frames = big_sequence(w, h, n, seed=1080, cut_at=37)
    if weighted_pred:
        for i in range(8):                    # a fade-in: the weight analysis picks weights
            frames[i] = np.clip(frames[i].astype(np.float32) * (0.4 + 0.07 * i) + 2 * i, 0, 255).astype(np.uint8)
    p = slicetype_params(w, h, weighted_pred=weighted_pred)
    qp_gpu, qp_orc = {}, {}
    got = gpu_decide(ctx, p, frames, qp_gpu)
    want = host.decide_with(slicetype_oracle_lib(), p, frames, qp_orc)
    assert got == want, [z for z in zip(got, want) if z[0] != z[1]][:6]
    assert sorted(f for f, _ in got) == list(range(n))
    assert any(t == 1 for f, t in got if f == 37) or any(t in (1, 2) for f, t in got if f == 37)      # the cut became a keyframe
    for fr in qp_orc:
        assert np.array_equal(qp_gpu[fr], qp_orc[fr]), ("f_qp_offset", fr)


def test_config3_8k_frame_costs_match_oracle(ctx):
    w, h, bframes = 7680, 4320, 16
    frames = big_sequence(w, h, 4, seed=4320)
    o = oracle()
    p = OrcLaParams()
    p.width, p.height, p.mb_width, p.mb_height = w, h, w // 16, h // 16
    p.subpel_refine, p.me_method, p.me_range, p.mv_range = 7, 1, 16, 512
    p.bframes, p.bframe_bias, p.weighted_bipred, p.aq_mode, p.vbv, p.do_edges, p.weighted_pred = bframes, 0, 1, 0, 0, 1, 0
    n = 2 * 4 * p.mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, 1)
    nmb = p.mb_width * p.mb_height
    la = x.Lookahead(ctx, w, h, bframes=bframes, aq_mode=0, mb_tree=1, n_slots=4)
    ofr = (C.c_void_p * 6)()
    try:
        for i, f in enumerate(frames):
            la.frame_put(i, f)
            ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
            o.orc_la_frame_set_qscale(ofr[i], np.full(nmb, 256, np.uint16))
        for (p0, p1, b) in [(0, 1, 1), (0, 3, 3), (0, 3, 1), (1, 3, 2)]:
            assert la.frame_cost([0, 1, 2, 3], p0, p1, b) == o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * n, ofr, p0, p1, b), (p0, p1, b)
            for l in range(2):
                for d in range(3):
                    a = np.zeros((nmb, 2), np.int16)
                    o.orc_la_frame_get(ofr[b], 0, l, d, ptr(a))
                    mv, co = la.get_mvs(b, l, d)
                    assert np.array_equal(mv, a), ("mvs", p0, p1, b, l, d, int((mv != a).any(axis=1).sum()))
                    if a[0, 0] != 0x7FFF:
                        c2 = np.zeros(nmb, np.int32)
                        o.orc_la_frame_get(ofr[b], 1, l, d, ptr(c2))
                        assert np.array_equal(co, c2), ("mv_costs", p0, p1, b, l, d)
            lc = np.zeros(nmb, np.uint16)
            o.orc_la_frame_get(ofr[b], 2, b - p0, p1 - b, ptr(lc))
            assert np.array_equal(la.get_costs(b, b - p0, p1 - b), lc), ("lowres_costs", p0, p1, b)
    finally:
        la.close()
        for i in range(4):
            o.orc_la_frame_delete(ofr[i])


def test_config3_8k_bframes16_trellis_scheduling_invariance(ctx):
    """rc-lookahead 250 / bframes 16 / b-adapt 2 at 7680x4320: the decisions of the prefetching, run-ahead schedule equal the
    on-demand ones (every search launched inside the cost request that needs it, in the reference's order)"""
    w, h, n = 7680, 4320, 40
    frames = big_sequence(w, h, n, seed=8, cut_at=23)
    p = slicetype_params(w, h, bframes=16, b_adapt=2, rc_lookahead=250, keyint_max=250)
    a = gpu_decide(ctx, p, frames, prefetch=0, run_ahead=0)
    b = gpu_decide(ctx, p, frames)
    assert a == b and sorted(f for f, _ in a) == list(range(n))


@pytest.mark.parametrize("wh", [(1920, 1080), (1366, 768), (34, 18), (16, 16), (48, 32)])
def test_odd_and_tiny_sizes_match_oracle_harness(ctx, wh):
    w, h = wh
    n = 24
    frames = synth_sequence(w, h, n, seed=w + h, cut_at=13) if w < 512 else big_sequence(w, h, n, seed=w, cut_at=13)
    p = slicetype_params(w, h, rc_lookahead=10, bframes=2)
    got = gpu_decide(ctx, p, frames)
    want = host.decide_with(slicetype_oracle_lib(), p, frames)
    assert got == want, [z for z in zip(got, want) if z[0] != z[1]][:6]


@pytest.mark.parametrize("cfg", [dict(bframes=3, b_adapt=1, rc_lookahead=40), dict(bframes=5, b_adapt=2, rc_lookahead=30),
                                 dict(bframes=3, b_adapt=1, rc_lookahead=40, weighted_pred=1), dict(bframes=8, b_adapt=2, rc_lookahead=60)])
def test_speculative_cost_requests_do_not_change_anything(ctx, cfg):
    """x264cu_lookahead_finalize_batch answers cost requests ahead of time in the variant the reference's order usually yields;
    a request of the other variant is computed on demand.  Decisions and MB-tree offsets with and without it, and against the
    oracle-backed host logic, must be identical -- and the speculation must actually be used."""
    w, h, n = 640, 368, 90
    frames = big_sequence(w, h, n, seed=77, cut_at=41)
    if cfg.get("weighted_pred"):
        for i in range(10):
            frames[i] = np.clip(frames[i].astype(np.float32) * (0.4 + 0.06 * i) + 2 * i, 0, 255).astype(np.uint8)
    p = slicetype_params(w, h, **cfg)
    res = {}
    for spec in (1, 0):
        st = x.Slicetype.from_params(ctx, p)
        try:
            st.set_speculation(spec)
            qp = {}
            res[spec] = (st.decide(frames, qp), qp, st.speculation_stats())
        finally:
            st.close()
    assert res[1][0] == res[0][0]
    assert res[1][1].keys() == res[0][1].keys() and all(np.array_equal(res[1][1][k], res[0][1][k]) for k in res[0][1])
    launched, hits, misses = res[1][2]
    assert launched > 0 and hits > misses, res[1][2]
    assert res[0][2][0] == 0 and res[0][2][1] == 0
    want = host.decide_with(slicetype_oracle_lib(), p, frames, {})
    assert res[1][0] == want
