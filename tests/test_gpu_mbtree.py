"""GPU parity of MB-tree (x264cu_lookahead_mbtree_* = macroblock_tree_propagate / _finish, encoder/slicetype.c:1029-1184,
common/mc.c:511-598): (1) macroblock_tree's call sequence replayed on the device and on the oracle (which
tests/test_oracle_mbtree.py pins to the compiled reference): i_propagate_cost exact, f_qp_offset bit-exact; (2) end to end:
the f_qp_offset of every non-B picture out of x264cu_slicetype_step versus the reference ENCODER's own (where it travelled)
and versus the same host logic over the oracle."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, have_ref, ptr, OrcLaParams, synth_sequence, slicetype_oracle_lib
import test_slicetype_host as host
from test_oracle_mbtree import replay_macroblock_tree, T_P, T_B, T_BREF

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_la()
    c = x.Context(0)
    yield c
    c.close()


# (w, h, bframes, weighted_pred, aq, types of frames 1..n, b_pyramid)
REPLAY = [
    (112, 80, 3, 0, 0, [T_B, T_B, T_P, T_B, T_P], False),
    (112, 80, 3, -1, 1, [T_P, T_B, T_BREF, T_B, T_P], True),
    (352, 288, 2, 1, 1, [T_B, T_P, T_B, T_B, T_P], False),
    # static content, long chain: i_propagate_cost saturates at MC_CLIP_ADD's (1<<15)-1 (common/mc.h:29); wp = "static" marker
    (112, 80, 2, "static", 0, [T_P, T_B, T_P] * 12, False),
]


@pytest.mark.parametrize("cfg", REPLAY)
def test_mbtree_replay_matches_oracle(ctx, cfg):
    w, h, bframes, wp, aq_on, types, b_pyramid = cfg
    static = wp == "static"
    wp = 0 if static else wp
    o = oracle()
    p = OrcLaParams()
    p.width, p.height, p.mb_width, p.mb_height = w, h, (w + 15) // 16, (h + 15) // 16
    p.subpel_refine, p.me_method, p.me_range, p.mv_range = 7, 1, 16, 512
    p.bframes, p.bframe_bias, p.weighted_bipred, p.aq_mode, p.vbv, p.do_edges, p.weighted_pred = bframes, 0, 1, aq_on, 0, 1, wp
    nfr = len(types) + 1
    frames = synth_sequence(w, h, nfr, seed=w + 7, cut_at=None)
    if static:
        noise = np.random.default_rng(11)
        frames = [np.clip(frames[0].astype(np.int16) + noise.integers(-1, 2, frames[0].shape), 0, 255).astype(np.uint8) for _ in frames]
    if wp:
        frames = [np.clip(f.astype(np.float32) * (0.55 + 0.09 * i) + 3 * i, 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
    n = 2 * 4 * p.mv_range
    tab = np.zeros(2 * n + 1, np.uint16)
    o.orc_cost_mv_table(tab, n, 1)
    nmb = p.mb_width * p.mb_height
    la = x.Lookahead(ctx, w, h, bframes=bframes, aq_mode=aq_on, mb_tree=1, n_slots=nfr, weighted_pred=wp)
    ofr = (C.c_void_p * (nfr + 2))()
    rng = np.random.default_rng(3)
    for i, f in enumerate(frames):
        q = rng.integers(180, 400, nmb).astype(np.uint16) if aq_on else np.full(nmb, 256, np.uint16)
        aq = (rng.normal(0, 1.5, nmb) if aq_on else np.zeros(nmb)).astype(np.float32)
        la.frame_put(i, f, q)
        la.set_qp_offset_aq(i, aq)
        ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
        o.orc_la_frame_set_qscale(ofr[i], q)
        o.orc_la_frame_set_qp_offset_aq(ofr[i], ptr(aq))
    slots = list(range(nfr))
    fps_prop, fps_fin, strength = np.float32(0.5 / 256), 512, np.float32(5.0) * (np.float32(1.0) - np.float32(0.6))

    requested = []

    def cost(p0, p1, b):
        assert la.frame_cost(slots, p0, p1, b) == o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * n, ofr, p0, p1, b), (p0, p1, b)
        requested.append((p0, p1, b))

    def reset(i):
        la.mbtree_reset(i)
        o.orc_la_mbtree_reset(ofr[i])

    def propagate(p0, p1, b, referenced):
        la.mbtree_propagate(slots, p0, p1, b, referenced, fps_prop)
        o.orc_la_mbtree_propagate(C.byref(p), ofr, p0, p1, b, referenced, fps_prop)

    def finish(i, dist):
        la.mbtree_finish(i, fps_fin, dist, strength)
        o.orc_la_mbtree_finish(ofr[i], fps_fin, dist, strength)

    try:
        # as in the encoder, every non-B picture has been costed against the previous one before the tree is built (the
        # reference computes a picture's intra costs inside its first cost request, this backend when the picture is queued)
        all_types = [T_P] + types
        prev = 0
        for k in range(1, nfr):
            if all_types[k] == T_P:
                cost(prev, k, k)
                prev = k
        touched = replay_macroblock_tree(all_types, b_pyramid, cost, reset, propagate, finish)
        seen = False
        peak = 0
        for k in sorted(touched):
            want = np.zeros(nmb, np.uint16)
            o.orc_la_frame_get_mbtree(ofr[k], 2, 0, ptr(want))
            got = la.get_propagate_cost(k)
            assert np.array_equal(got, want), ("propagate_cost", k, np.argwhere(got != want)[:5])
            seen |= bool(want.any())
            peak = max(peak, int(got.max()))
        assert seen
        assert peak == 32767 if static else peak <= 32767
        for k in range(1, nfr):
            want = np.zeros(nmb, np.float32)
            o.orc_la_frame_get_mbtree(ofr[k], 0, 0, ptr(want))
            got = la.get_qp_offset(k)
            assert np.array_equal(got, want), ("qp_offset", k, float(np.abs(got - want).max()))
            for d in range(bframes + 1):
                wd = C.c_float()
                o.orc_la_frame_get_mbtree(ofr[k], 3, d, C.byref(wd))
                assert la.get_weighted_cost_delta(k, d) == wd.value, ("weighted_cost_delta", k, d)
        # slicetype_frame_cost_recalculate (slicetype.c:999-1024) of every requested cost
        o.orc_la_frame_cost_recalculate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        for p0, p1, b in requested:
            is_b = int(all_types[b] in (T_B, T_BREF))
            want = o.orc_la_frame_cost_recalculate(C.byref(p), ofr, p0, p1, b, is_b)
            rows = np.zeros(p.mb_height, np.int32)
            o.orc_la_frame_get(ofr[b], 5, b - p0, p1 - b, ptr(rows))
            got, got_rows = la.frame_cost_recalculate(b, b - p0, p1 - b, is_b)
            assert got == want and np.array_equal(got_rows, rows), ("recalculate", p0, p1, b, got, want)
            assert np.array_equal(la.get_row_satds(b, b - p0, p1 - b), rows)
    finally:
        la.close()
        for k in range(nfr):
            o.orc_la_frame_delete(ofr[k])


@pytest.mark.parametrize("case", host.MBTREE_CASES)
def test_mbtree_qp_offsets_end_to_end(ctx, case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w, cut_at=cut)
    qp_ref, qp_orc, qp_gpu = {}, {}, {}
    if have_ref():
        p, want = host.reference_types(preset, opts, w, h, frames, qp_ref)
    else:
        pytest.skip("compiled reference did not travel (its option parsing provides the parameters)")
    want_orc = host.decide_with(slicetype_oracle_lib(), p, frames, qp_orc)
    st = x.Slicetype.from_params(ctx, p)
    try:
        got = st.decide(frames, qp_gpu)
    finally:
        st.close()
    assert got == want == want_orc
    compared = 0
    for fr, ty in want:
        if ty in (4, 5):
            continue
        assert np.array_equal(qp_gpu[fr], qp_orc[fr]), ("vs oracle", fr, float(np.abs(qp_gpu[fr] - qp_orc[fr]).max()))
        assert np.array_equal(qp_gpu[fr], qp_ref[fr]), ("vs reference encoder", fr, float(np.abs(qp_gpu[fr] - qp_ref[fr]).max()))
        compared += 1
    assert compared >= 5 and any(np.abs(q).max() > 0.5 for q in qp_gpu.values())


@pytest.mark.parametrize("case", host.DEFAULT_PATH_CASES)
def test_default_path_i420_end_to_end(ctx, case):
    """preset medium as it is (aq-mode 1, weightp 2, mb-tree, psy) and two variants: I420 pictures into x264cu_slicetype_step_i420
    -- adaptive quantisation, weight analysis, lookahead, slice-type decision and MB-tree all on the device -- against the
    reference ENCODER's frame types and f_qp_offset, and against the same host logic over the oracle"""
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w + 5, cut_at=cut)
    for i in range(8):
        frames[i] = np.clip(frames[i].astype(np.float32) * (0.4 + 0.07 * i) + 2 * i, 0, 255).astype(np.uint8)
    cb = np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8)
    if not have_ref():
        pytest.skip("compiled reference did not travel (its option parsing provides the parameters)")
    qp_ref, qp_orc, qp_gpu = {}, {}, {}
    p, want = host.reference_types(preset, opts, w, h, frames, qp_ref)
    p.aq_strength = float(opts.split("aq-strength=")[1].split(":")[0]) if "aq-strength" in opts else 1.0
    want_orc = host.decide_with(slicetype_oracle_lib(), p, frames, qp_orc, chroma=(cb, cb))
    st = x.Slicetype.from_params(ctx, p)
    try:
        got = st.decide(frames, qp_gpu, chroma=[(cb, cb)] * n)
    finally:
        st.close()
    assert got == want == want_orc
    compared = 0
    for fr, ty in want:
        if ty in (4, 5):
            continue
        assert np.array_equal(qp_gpu[fr], qp_orc[fr]), ("vs oracle", fr, float(np.abs(qp_gpu[fr] - qp_orc[fr]).max()))
        assert np.array_equal(qp_gpu[fr], qp_ref[fr]), ("vs reference encoder", fr, float(np.abs(qp_gpu[fr] - qp_ref[fr]).max()))
        compared += 1
    assert compared >= 5
