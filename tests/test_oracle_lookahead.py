"""Pins oracle/oracle_lookahead.c against the compiled reference's slicetype_frame_cost (encoder/slicetype.c:836-995):
every per-MB output (lowres_mvs, lowres_mv_costs, lowres_costs, intra costs) and every frame-level sum, for I / P / B
requests in several orders (the memo / temporal-predictor availability is order dependent, slicetype.c:629-642)."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import oracle, ref, have_ref, ptr, OrcLaParams, la_params_from_ref, synth_sequence

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

REQUESTS = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (0, 2, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 3, 2), (1, 2, 2),
            (2, 4, 3), (2, 4, 4), (1, 4, 4), (1, 4, 2), (1, 4, 3), (4, 4, 4), (3, 5, 4), (3, 5, 5), (2, 5, 5)]


def compare_frame(r, la, of, idx, p, tag, written):
    """written: set of (b-p0, p1-b) cost slots requested so far for this frame (other slots are uninitialised memory
    in the reference); searched vector slots are recognised by their cleared 0x7FFF sentinel"""
    n = p.mb_width * p.mb_height
    B = p.bframes
    # without mbtree / vbv the edge MBs are never visited (slicetype.c:823-833): their slots are uninitialised memory
    mask = np.ones((p.mb_height, p.mb_width), bool)
    if not (p.do_edges or p.mb_width <= 2 or p.mb_height <= 2):
        mask[:] = False
        mask[1:-1, 1:-1] = True
    mask = mask.reshape(-1)
    for l in range(2 if B else 1):
        for d in range(B + 1):
            a = np.zeros((n, 2), np.int16); b = np.zeros((n, 2), np.int16)
            r.xref_la_get(la, idx, 0, l, d, ptr(a)); oracle().orc_la_frame_get(of, 0, l, d, ptr(b))
            assert np.array_equal(a, b), (tag, "mvs", idx, l, d, np.argwhere(a != b)[:4])
            if a[0, 0] != 0x7FFF:
                a = np.zeros(n, np.int32); b = np.zeros(n, np.int32)
                r.xref_la_get(la, idx, 1, l, d, ptr(a)); oracle().orc_la_frame_get(of, 1, l, d, ptr(b))
                assert np.array_equal(a[mask], b[mask]), (tag, "mv_costs", idx, l, d, np.argwhere(a != b)[:4])
    for (i, j) in written:
        e1 = np.zeros(3, np.int32); e2 = np.zeros(3, np.int32)
        r.xref_la_get(la, idx, 4, i, j, ptr(e1)); oracle().orc_la_frame_get(of, 4, i, j, ptr(e2))
        assert np.array_equal(e1[:2], e2[:2]), (tag, "cost_est", idx, i, j, e1, e2)
        if j == 0:
            assert e1[2] == e2[2], (tag, "intra_mbs", idx, i, e1, e2)
        a = np.zeros(n, np.uint16); b = np.zeros(n, np.uint16)
        r.xref_la_get(la, idx, 2, i, j, ptr(a)); oracle().orc_la_frame_get(of, 2, i, j, ptr(b))
        assert np.array_equal(a[mask], b[mask]), (tag, "lowres_costs", idx, i, j, np.argwhere(a != b)[:4])
    e1 = np.zeros(3, np.int32); e2 = np.zeros(3, np.int32)
    r.xref_la_get(la, idx, 4, 0, 0, ptr(e1)); oracle().orc_la_frame_get(of, 4, 0, 0, ptr(e2))
    assert np.array_equal(e1[:2], e2[:2]), (tag, "intra cost_est", idx, e1, e2)
    a = np.zeros(n, np.int32); b = np.zeros(n, np.int32)
    r.xref_la_get(la, idx, 3, 0, 0, ptr(a)); oracle().orc_la_frame_get(of, 3, 0, 0, ptr(b))
    assert np.array_equal(a[mask], b[mask]), (tag, "intra", idx)


CONFIGS = [
    ("medium", "bframes=3", (112, 80)),                                  # default weightp=2 + mb-tree + psy
    ("medium", "weightp=1:bframes=2:subme=1", (96, 64)),
    ("medium", "weightp=0:bframes=3", (112, 80)),                        # X264_WEIGHTP_FAKE (encoder.c:1316-1317)
    ("medium", "weightp=0:bframes=3:subme=1:no-mbtree=1", (112, 80)),
    ("medium", "weightp=0:bframes=4:me=umh:merange=24:aq-mode=0", (96, 96)),
    ("medium", "weightp=0:bframes=3:me=dia:weightb=0:no-mbtree=1", (100, 60)),
    ("ultrafast", "bframes=2", (64, 48)),
    ("medium", "no-mbtree=1:bframes=2:subme=2:aq-mode=0", (96, 64)),      # weights analysis with border MBs never costed
    ("medium", "weightp=0:bframes=3:vbv-maxrate=1000:vbv-bufsize=1000", (80, 64)),
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_frame_cost_matches_reference(cfg):
    preset, opts, (w, h) = cfg
    _libs._bind_la()
    o, r = oracle(), ref()
    hnd = r.xref_open(w, h, preset.encode(), opts.encode(), 0)
    assert hnd
    try:
        p = la_params_from_ref(hnd, w, h)
        nfr = 6
        frames = synth_sequence(w, h, nfr, seed=w + h, cut_at=4)
        if p.weighted_pred:
            # a fade: the lookahead weight analysis (slicetype.c:284-501) must find (and both sides agree on) weights
            frames = [np.clip(f.astype(np.float32) * (0.55 + 0.09 * i) + 3 * i, 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
        n = 2 * 4 * p.mv_range
        tab = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table(hnd, tab, n)
        rng = np.random.default_rng(1)
        la = r.xref_la_new(hnd, nfr)
        ofr = (C.c_void_p * (nfr + 2))()
        for i, f in enumerate(frames):
            q = rng.integers(180, 400, p.mb_width * p.mb_height).astype(np.uint16)
            assert r.xref_la_set_frame(la, i, ptr(f), w, ptr(q)) == 0
            ofr[i] = o.orc_la_frame_new(C.byref(p), ptr(f), w)
            o.orc_la_frame_set_qscale(ofr[i], q)
        reqs = [q for q in REQUESTS if q[1] < nfr and q[1] - q[0] <= p.bframes + 1]
        written = {i: set() for i in range(nfr)}
        weights_seen = []
        for (p0, p1, b) in reqs:
            if not (p0 == p1 and written[b]):      # an I request after any other request is a memo hit: nothing is written
                written[b].add((b - p0, p1 - b))
            s1 = r.xref_la_frame_cost(la, p0, p1, b)
            s2 = o.orc_la_frame_cost(C.byref(p), tab.ctypes.data + 2 * n, ofr, p0, p1, b)
            if p.weighted_pred and b == p1 and p0 != p1:
                w1 = np.zeros(4, np.int32); w2 = np.zeros(4, np.int32)
                r.xref_la_get(la, b, 6, 0, 0, ptr(w1)); o.orc_la_frame_get(ofr[b], 6, 0, 0, ptr(w2))
                if w1[0] or w2[0]:
                    weights_seen.append(tuple(w1))
                    assert np.array_equal(w1, w2), (cfg, "weights", p0, p1, b, w1, w2)
            assert s1 == s2, (cfg, p0, p1, b, s1, s2)
            compare_frame(r, la, ofr[b], b, p, (cfg, p0, p1, b), written[b])
            if p.vbv:
                a = np.zeros(p.mb_height, np.int32); bb = np.zeros(p.mb_height, np.int32)
                r.xref_la_get(la, b, 5, b - p0, p1 - b, ptr(a)); o.orc_la_frame_get(ofr[b], 5, b - p0, p1 - b, ptr(bb))
                assert np.array_equal(a, bb), (cfg, "row_satds", p0, p1, b)
        if p.weighted_pred and 'weightp=0' not in opts:
            assert weights_seen, 'the fade should have produced at least one weighted P search'
        for i in range(nfr):
            o.orc_la_frame_delete(ofr[i])
        r.xref_la_free(la)
    finally:
        r.xref_close(hnd)
