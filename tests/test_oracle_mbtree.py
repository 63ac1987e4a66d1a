"""Pins the oracle's MB-tree restatement (oracle/oracle_lookahead.c: orc_la_mbtree_*) against the compiled reference:
macroblock_tree_propagate + mbtree_propagate_cost/_list (slicetype.c:1050-1089, mc.c:511-598), macroblock_tree_finish
(slicetype.c:1029-1048) and x264_log2 (base.h:226-230), replaying macroblock_tree's call sequence for two GOP shapes.
i_propagate_cost is integer and compared exactly; f_qp_offset is float (the reference builds with -ffast-math, and its own
checkasm accepts +-1 / 1e-4 on the propagate amounts, tools/checkasm.c:1798-1805): tolerance 1e-4, exact match reported."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import oracle, ref, have_ref, ptr, la_params_from_ref, synth_sequence

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

T_P, T_BREF, T_B, T_I = 3, 4, 5, 2

# (reference options, (w, h), types of frames 1..n (frame 0 is the last non-B of the previous GOP), fade)
CASES = [
    ("weightp=0:no-psy=1:bframes=3:aq-mode=0:b-pyramid=none", (112, 80), [T_B, T_B, T_P, T_B, T_P], False),
    ("weightp=0:bframes=3:aq-mode=1", (112, 80), [T_P, T_B, T_BREF, T_B, T_P], True),       # fake weights + pyramid + AQ
    ("weightp=2:bframes=2:aq-mode=1:b-pyramid=none", (96, 64), [T_B, T_P, T_B, T_B, T_P], True),
    # static content over a long chain: i_propagate_cost reaches MC_CLIP_ADD's ceiling (1<<15)-1 (common/mc.h:29)
    ("weightp=0:bframes=2:aq-mode=0:b-pyramid=none", (112, 80), [T_P, T_B, T_P] * 12, "static"),
]


def replay_macroblock_tree(all_types, b_pyramid, cost, reset, propagate, finish):
    """macroblock_tree( frames, num_frames = len(all_types) - 1, b_intra = 0 ) with rc-lookahead > 0 (slicetype.c:1091-1184) as a
    sequence of calls of the four callbacks; returns the set of frames whose i_propagate_cost was reset (= is defined)"""
    is_b = lambda t: t in (T_B, T_BREF)
    touched = set()

    def rst(i):
        touched.add(i)
        reset(i)

    nfr = len(all_types)
    i = nfr - 1
    while i > 0 and is_b(all_types[i]):
        i -= 1
    last_nonb = i
    rst(last_nonb)
    bframes = 0
    while i > 1:
        i -= 1
        cur_nonb = i
        while is_b(all_types[cur_nonb]) and cur_nonb > 0:
            cur_nonb -= 1
        if cur_nonb < 1:
            break
        cost(cur_nonb, last_nonb, last_nonb)
        rst(cur_nonb)
        bframes = last_nonb - cur_nonb - 1
        if b_pyramid and bframes > 1:
            middle = (bframes + 1) // 2 + cur_nonb
            cost(cur_nonb, last_nonb, middle)
            rst(middle)
            while i > cur_nonb:
                p0 = middle if i > middle else cur_nonb
                p1 = middle if i < middle else last_nonb
                if i != middle:
                    cost(p0, p1, i)
                    propagate(p0, p1, i, 0)
                i -= 1
            propagate(cur_nonb, last_nonb, middle, 1)
        else:
            while i > cur_nonb:
                cost(cur_nonb, last_nonb, i)
                propagate(cur_nonb, last_nonb, i, 0)
                i -= 1
        propagate(cur_nonb, last_nonb, last_nonb, 1)
        last_nonb = cur_nonb
    finish(last_nonb, last_nonb)
    if b_pyramid and bframes > 1:
        finish(last_nonb + (bframes + 1) // 2, 0)
    return touched


def test_log2_table():
    _libs._bind_la()
    lut = (C.c_float * 128).in_dll(ref(), "x264_log2_lut")
    for x in [(128 + i) << k for i in range(128) for k in (0, 9, 24)] + [1, 2, 3, 255, 256, 65535, 1 << 20, 12345678]:
        lz = 32 - x.bit_length()
        want = np.float32(lut[((x << lz) >> 24) & 0x7f]) + np.float32(31 - lz)
        assert oracle().orc_log2(x) == want, x


@pytest.mark.parametrize("case", CASES)
def test_mbtree_matches_reference(case):
    opts, (w, h), types, fade = case
    _libs._bind_la()
    o, r = oracle(), ref()
    hnd = r.xref_open(w, h, b"medium", opts.encode(), 0)
    assert hnd
    try:
        p = la_params_from_ref(hnd, w, h)
        assert p.do_edges
        nfr = len(types) + 1
        frames = synth_sequence(w, h, nfr, seed=w + 7, cut_at=None)
        if fade == "static":
            noise = np.random.default_rng(11)
            frames = [np.clip(frames[0].astype(np.int16) + noise.integers(-1, 2, frames[0].shape), 0, 255).astype(np.uint8) for _ in frames]
        elif fade:
            frames = [np.clip(f.astype(np.float32) * (0.55 + 0.09 * i) + 3 * i, 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
        n = 2 * 4 * p.mv_range
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table(hnd, tab, n)
        nmb = p.mb_width * p.mb_height
        la = r.xref_la_new(hnd, nfr)
        ofr = (C.c_void_p * (nfr + 2))()
        rng = np.random.default_rng(3)
        dur = 1 / 25.
        for i, f in enumerate(frames):
            q = rng.integers(180, 400, nmb).astype(np.uint16) if p.aq_mode else np.full(nmb, 256, np.uint16)
            assert r.xref_la_set_frame(la, i, ptr(f), w, ptr(q)) == 0
            aq = (rng.normal(0, 1.5, nmb) if p.aq_mode else np.zeros(nmb)).astype(np.float32)
            r.xref_la_set_qp_offset_aq(la, i, ptr(aq))
            r.xref_la_set_type(la, i, T_P if i == 0 else types[i - 1], dur)
            ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
            o.orc_la_frame_set_qscale(ofr[i], q)
            o.orc_la_frame_set_qp_offset_aq(ofr[i], ptr(aq))
        all_types = [T_P] + types
        requested = []
        def cost(p0, p1, b):
            s1, s2 = r.xref_la_frame_cost(la, p0, p1, b), o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * n, ofr, p0, p1, b)
            assert s1 == s2, (p0, p1, b)
            requested.append((p0, p1, b))
        fps_prop = np.float32(dur) / (np.float32(dur) * np.float32(256.0)) * np.float32(0.5)
        fps_fin = int(round(dur / dur * 256 / 0.5))
        strength = np.float32(5.0) * (np.float32(1.0) - np.float32(0.6))

        def propagate(p0, p1, b, referenced):
            r.xref_la_mbtree_propagate(la, dur, p0, p1, b, referenced)
            o.orc_la_mbtree_propagate(C.byref(p), ofr, p0, p1, b, referenced, fps_prop)

        def reset(i):
            r.xref_la_mbtree_reset(la, i)
            o.orc_la_mbtree_reset(ofr[i])

        def finish(i, dist):
            r.xref_la_mbtree_finish(la, i, dur, dist)
            o.orc_la_mbtree_finish(ofr[i], fps_fin, dist, strength)

        touched = replay_macroblock_tree(all_types, "b-pyramid=none" not in opts, cost, reset, propagate, finish)
        is_b = lambda t: t in (T_B, T_BREF)
        exact = total = 0
        checked_nonzero = False
        peak = 0
        for k in range(1, nfr):          # frame 0 (the previous GOP's last non-B) is never touched with b_intra = 0: uninitialised in the reference
            a = np.zeros(nmb, np.uint16); b = np.zeros(nmb, np.uint16)
            r.xref_la_get_mbtree(la, k, 2, 0, ptr(a)); o.orc_la_frame_get_mbtree(ofr[k], 2, 0, ptr(b))
            if k in touched:
                assert np.array_equal(a, b), ("propagate_cost", k, np.argwhere(a != b)[:5], a[a != b][:5], b[a != b][:5])
                checked_nonzero |= bool(a.any())
                peak = max(peak, int(a.max()))
            qa = np.zeros(nmb, np.float32); qb = np.zeros(nmb, np.float32)
            r.xref_la_get_mbtree(la, k, 0, 0, ptr(qa)); o.orc_la_frame_get_mbtree(ofr[k], 0, 0, ptr(qb))
            assert np.allclose(qa, qb, atol=1e-4, rtol=0), ("qp_offset", k, np.abs(qa - qb).max())
            exact += int((qa == qb).sum()); total += nmb
            for d in range(p.bframes + 1):
                wa = C.c_float(); wb = C.c_float()
                r.xref_la_get_mbtree(la, k, 3, d, C.byref(wa)); o.orc_la_frame_get_mbtree(ofr[k], 3, d, C.byref(wb))
                assert wa.value == wb.value, ("weighted_cost_delta", k, d, wa.value, wb.value)
        assert checked_nonzero
        assert peak <= 32767
        if fade == "static":
            assert peak == 32767, "the static case is meant to saturate i_propagate_cost (peak %d)" % peak
        assert exact == total, "f_qp_offset bit-exact on %d of %d macroblocks" % (exact, total)
        # slicetype_frame_cost_recalculate (slicetype.c:999-1024) of every requested cost: MB-tree's offsets for P / I, AQ's for B
        r.xref_la_frame_cost_recalculate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_frame_cost_recalculate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        changed = 0
        for p0, p1, b in requested:
            before = r.xref_la_frame_cost(la, p0, p1, b)
            s1 = r.xref_la_frame_cost_recalculate(la, p0, p1, b)
            s2 = o.orc_la_frame_cost_recalculate(C.byref(p), ofr, p0, p1, b, int(is_b(all_types[b])))
            assert s1 == s2, ("recalculate", p0, p1, b, s1, s2)
            ra = np.zeros(p.mb_height, np.int32); rb = np.zeros(p.mb_height, np.int32)
            r.xref_la_get(la, b, 5, b - p0, p1 - b, ptr(ra)); o.orc_la_frame_get(ofr[b], 5, b - p0, p1 - b, ptr(rb))
            assert np.array_equal(ra, rb), ("recalculate rows", p0, p1, b)
            changed += s1 != before
        assert changed or not p.aq_mode
        for k in range(nfr):
            o.orc_la_frame_delete(ofr[k])
        r.xref_la_free(la)
    finally:
        r.xref_close(hnd)
