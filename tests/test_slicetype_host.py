"""The host-side slice-type logic (x264_b200/csrc/slicetype.c: decide / analyse / scenecut / trellis path / MB-tree's
request order / the synchronous frame queue) against the reference ENCODER's own frame-type output, on CPU: the five
lookahead entry points it calls are served by the oracle here (tests/csrc/slicetype_oracle_glue.c); on the GPU box
tests/test_gpu_slicetype.py runs the same comparison with the CUDA lookahead underneath."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import ref, have_ref, slicetype_oracle_lib, synth_sequence
from x264_b200.binding_ext import SlicetypeParams, LookaheadParams

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

# (preset, options, (w,h), n_frames, cut_at)
# NB "weightp=0" alone does not switch the lookahead weight analysis off: with mb-tree AND psy the encoder turns it back on
# as X264_WEIGHTP_FAKE (encoder.c:1316-1317); no-psy / no-mbtree cases run without it, the last four with it.
CASES = [
    ("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 40, 17),
    ("medium", "weightp=0:no-psy=1:bframes=3:b-adapt=2:rc-lookahead=12:keyint=40", (96, 64), 36, 20),
    ("medium", "weightp=0:bframes=2:b-adapt=0:no-mbtree=1:rc-lookahead=0:scenecut=0:keyint=12", (64, 48), 30, None),
    ("medium", "weightp=0:no-psy=1:bframes=4:b-pyramid=none:rc-lookahead=8:subme=1", (80, 64), 30, 9),
    ("ultrafast", "keyint=20", (64, 48), 45, 21),
    ("medium", "weightp=0:no-psy=1:bframes=0:rc-lookahead=5:keyint=25", (64, 64), 30, 11),
    ("slower", "weightp=0:no-psy=1:bframes=3:rc-lookahead=16:keyint=50:me=umh", (96, 80), 30, 13),
    ("medium", "weightp=0:no-mbtree=1:bframes=3:b-adapt=2:rc-lookahead=20:keyint=60", (128, 96), 48, 25),
    # lookahead weight analysis on (x264_weights_analyse, slicetype.c:284-501): the default preset, fake weights, weightp 1
    ("medium", "bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 40, 17),
    ("medium", "weightp=0:bframes=3:rc-lookahead=12", (96, 64), 36, 20),
    ("medium", "weightp=1:bframes=4:b-adapt=2:rc-lookahead=14", (96, 80), 36, 14),
    ("slow", "rc-lookahead=16:keyint=50", (128, 96), 40, 19),
    # open GOP (keyframes become I, counted in display order) and periodic intra refresh (no keyframes but the first)
    ("medium", "weightp=0:no-psy=1:open-gop=1:bframes=3:rc-lookahead=12:keyint=16:min-keyint=4", (96, 64), 60, 41),
    ("medium", "open-gop=1:bframes=3:b-adapt=2:rc-lookahead=14:keyint=20", (96, 64), 56, 23),
    ("medium", "weightp=0:no-psy=1:intra-refresh=1:bframes=2:rc-lookahead=10:keyint=24", (96, 64), 60, 31),
    ("medium", "intra-refresh=1:rc-lookahead=12:keyint=30", (112, 80), 50, 17),
]


def params_from_ref(hnd, w, h):
    r = ref()
    g = lambda n: r.xref_param(hnd, n.encode())
    la = LookaheadParams(w, h, g("subme"), min(g("me"), 2), g("merange"), g("mvrange"), g("bframes"), g("b_bias"), g("weightb"),
                         g("aq_mode"), g("mbtree"), g("vbv"), 0, -1 if g("weightp") < 0 else int(g("weightp") != 0))
    p = SlicetypeParams(la, g("keyint_max"), g("keyint_min"), g("scenecut"), g("b_adapt"), g("b_pyramid"), g("lookahead"),
                        g("psy"), g("ref"), 0)
    p.open_gop, p.intra_refresh = g("open_gop"), g("intra_refresh")
    return p


def decide_with(lib, p, frames, qp_out=None, chroma=None, forced=None, rc_out=None, stats=None):
    """qp_out: dict filled with frame -> f_qp_offset (MB-tree's output) for every non-B picture, read when it is returned;
    chroma: (cb, cr) planes fed with every picture through x264cu_slicetype_step_i420 (adaptive quantisation inside)"""
    lib.x264cu_slicetype_open.argtypes = [C.c_void_p, C.POINTER(SlicetypeParams), C.POINTER(C.c_void_p)]
    lib.x264cu_slicetype_step.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.x264cu_slicetype_close.argtypes = [C.c_void_p]
    st = C.c_void_p()
    assert lib.x264cu_slicetype_open(C.c_void_p(1), C.byref(p), C.byref(st)) == 0
    out = []
    fr, ty = C.c_int(), C.c_int()
    lib.x264cu_slicetype_get_qp_offset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    nmb = ((p.la.width + 15) // 16) * ((p.la.height + 15) // 16)

    mbh = (p.la.height + 15) // 16
    lib.x264cu_slicetype_rc_analyse_slice.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    lib.x264cu_slicetype_get_planned.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]

    def note():
        out.append((fr.value, ty.value))
        if rc_out is not None:
            # x264_rc_analyse_slice of the picture just returned: (cost, rows, planned types, planned costs)
            cost, rows = C.c_int(-1), np.zeros(mbh, np.int32)
            if ty.value not in (4, 5) or p.la.vbv:
                assert lib.x264cu_slicetype_rc_analyse_slice(st, fr.value, C.byref(cost), rows.ctypes.data, None) == 0
            pt, ps = np.zeros(32, np.int32), np.zeros(32, np.int32)
            k = lib.x264cu_slicetype_get_planned(st, fr.value, pt.ctypes.data, ps.ctypes.data, 32) if p.la.vbv and p.rc_lookahead and ty.value not in (4, 5) else 0
            assert k >= 0
            rc_out[fr.value] = (cost.value, rows, list(pt[:k]), list(ps[:k]))
        if qp_out is not None and ty.value not in (4, 5):
            q = np.zeros(nmb, np.float32)
            assert lib.x264cu_slicetype_get_qp_offset(st, fr.value, q.ctypes.data) == 0
            qp_out[fr.value] = q

    lib.x264cu_slicetype_step_i420.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.x264cu_slicetype_set_next_type.argtypes = [C.c_void_p, C.c_int]
    for i, f in enumerate(frames):
        if forced is not None and forced[i]:
            assert lib.x264cu_slicetype_set_next_type(st, int(forced[i])) == 0
        if chroma is None:
            assert lib.x264cu_slicetype_step(st, f.ctypes.data, f.shape[1], None, C.byref(fr), C.byref(ty)) == 0
        else:
            assert lib.x264cu_slicetype_step_i420(st, f.ctypes.data, f.shape[1], chroma[0].ctypes.data, chroma[1].ctypes.data, chroma[0].shape[1],
                                                  C.byref(fr), C.byref(ty)) == 0
        if fr.value >= 0:
            note()
    while True:
        assert lib.x264cu_slicetype_step(st, None, 0, None, C.byref(fr), C.byref(ty)) == 0
        if fr.value < 0:
            break
        note()
    if stats is not None:
        lib.x264cu_slicetype_farthest_list1.argtypes = [C.c_void_p]
        stats["farthest_list1"] = lib.x264cu_slicetype_farthest_list1(st)
    lib.x264cu_slicetype_close(st)
    return out


def reference_types(preset, opts, w, h, frames, qp_out=None, forced=None, rc_out=None):
    """qp_out: dict filled with frame -> the f_qp_offset array the reference encoder used for it"""
    r = ref()
    r.xref_encode_types.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    r.xref_set_qp_capture.argtypes = [C.c_void_p]
    hnd = r.xref_open(w, h, preset.encode(), opts.encode(), 0)
    assert hnd
    try:
        p = params_from_ref(hnd, w, h)
        n = len(frames)
        luma = np.ascontiguousarray(np.stack(frames))
        idx = (C.c_int * (n + 8))()
        typ = (C.c_int * (n + 8))()
        nmb = ((w + 15) // 16) * ((h + 15) // 16)
        cap = np.zeros((n + 8, nmb), np.float32)
        r.xref_set_qp_capture(cap.ctypes.data if qp_out is not None else None)
        mbh = (h + 15) // 16
        stride = 2 + mbh + 64
        rc_cap = np.zeros((n + 8, stride), np.int32)
        r.xref_set_rc_capture.argtypes = [C.c_void_p, C.c_int]
        r.xref_set_rc_capture(rc_cap.ctypes.data if rc_out is not None else None, stride)
        r.xref_set_forced_types.argtypes = [C.c_void_p]
        ft = (C.c_int * n)(*[int(t) for t in forced]) if forced is not None else None
        r.xref_set_forced_types(ft)
        k = r.xref_encode_types(hnd, luma.ctypes.data, n, idx, typ)
        r.xref_set_forced_types(None)
        r.xref_set_qp_capture(None)
        r.xref_set_rc_capture(None, 0)
        assert k == n
        if rc_out is not None:
            for i in range(k):
                c = rc_cap[i]
                np_ = int(c[1 + mbh])
                rc_out[idx[i]] = (int(c[0]), c[1:1 + mbh].copy(), list(c[2 + mbh:2 + mbh + np_]), list(c[2 + mbh + 32:2 + mbh + 32 + np_]))
        if qp_out is not None:
            for i in range(k):
                qp_out[idx[i]] = cap[i].copy()
        return p, [(idx[i], typ[i]) for i in range(k)]
    finally:
        r.xref_close(hnd)


@pytest.mark.parametrize("case", CASES)
def test_frame_types_match_reference_encoder(case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w, cut_at=cut)
    if cut is not None and n > cut + 9:          # a two-frame flash, which must not become a scene cut
        frames[cut + 7] = np.full_like(frames[0], 235)
        frames[cut + 8] = np.full_like(frames[0], 235)
    if cut is not None:                          # a fade-in over the first pictures: exercises the weight analysis
        for i in range(min(10, cut)):
            frames[i] = np.clip(frames[i].astype(np.float32) * (0.35 + 0.065 * i) + 2 * i, 0, 255).astype(np.uint8)
    p, want = reference_types(preset, opts, w, h, frames)
    got = decide_with(slicetype_oracle_lib(), p, frames)
    assert got == want, (case, [x for x in zip(got, want) if x[0] != x[1]][:6])


# MB-tree end to end: the f_qp_offset the reference ENCODER used for every non-B picture versus macroblock_tree in the
# product's host logic (the device entry points served by the oracle here).  AQ off, so that f_qp_offset_aq is zero on both sides.
MBTREE_CASES = [
    ("medium", "aq-mode=0:weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 60, 37),
    ("medium", "aq-mode=0:weightp=0:bframes=3:b-adapt=2:rc-lookahead=12", (96, 64), 60, 41),          # fake weights (psy + mb-tree)
    ("medium", "aq-mode=0:weightp=0:no-psy=1:bframes=2:b-pyramid=none:rc-lookahead=8", (96, 64), 50, 31),
]


def mbtree_compare(want_types, got_types, qp_ref, qp_got, n_frames, lookahead):
    """frame types must agree everywhere; f_qp_offset is compared for every non-B picture (B pictures are coded with
    f_qp_offset_aq).  Returns (pictures compared, pictures bit-exact, worst absolute difference)."""
    assert got_types == want_types
    compared = exact = 0
    worst = 0.0
    for fr, ty in want_types:
        if ty in (4, 5):
            continue
        a, b = qp_ref[fr], qp_got[fr]
        compared += 1
        exact += int(np.array_equal(a, b))
        worst = max(worst, float(np.abs(a - b).max()))
    return compared, exact, worst


@pytest.mark.parametrize("case", MBTREE_CASES)
def test_mbtree_qp_offsets_match_reference_encoder(case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w, cut_at=cut)
    qp_ref, qp_got = {}, {}
    p, want = reference_types(preset, opts, w, h, frames, qp_ref)
    got = decide_with(slicetype_oracle_lib(), p, frames, qp_got)
    compared, exact, worst = mbtree_compare(want, got, qp_ref, qp_got, n, p.rc_lookahead)
    assert compared >= 5 and any(np.abs(q).max() > 0.5 for q in qp_got.values())
    assert exact == compared, "f_qp_offset bit-exact on %d of %d pictures, worst |diff| %.3g" % (exact, compared, worst)


# The whole default path: preset medium as it is (aq-mode 1, weightp 2, mb-tree, psy) and two variants, I420 pictures in, adaptive
# quantisation + weight analysis + lookahead + MB-tree inside: frame types and f_qp_offset of every non-B picture against the
# reference ENCODER.  (The reference shim feeds flat chroma, so does this test.)
DEFAULT_PATH_CASES = [
    ("medium", "bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 60, 37),
    ("medium", "aq-mode=2:rc-lookahead=12:b-adapt=2", (96, 64), 50, 31),
    ("medium", "aq-mode=3:aq-strength=1.3:weightp=1:bframes=2:rc-lookahead=8", (96, 64), 50, 27),
]


@pytest.mark.parametrize("case", DEFAULT_PATH_CASES)
def test_default_presets_with_adaptive_quant_match_reference_encoder(case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w + 5, cut_at=cut)
    for i in range(8):                                       # a fade-in: weights are chosen
        frames[i] = np.clip(frames[i].astype(np.float32) * (0.4 + 0.07 * i) + 2 * i, 0, 255).astype(np.uint8)
    flat = (np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8), np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8))
    qp_ref, qp_got = {}, {}
    p, want = reference_types(preset, opts, w, h, frames, qp_ref)
    p.aq_strength = float(opts.split("aq-strength=")[1].split(":")[0]) if "aq-strength" in opts else 1.0
    got = decide_with(slicetype_oracle_lib(), p, frames, qp_got, chroma=flat)
    compared, exact, worst = mbtree_compare(want, got, qp_ref, qp_got, n, p.rc_lookahead)
    assert compared >= 5 and exact == compared, "f_qp_offset bit-exact on %d of %d pictures, worst |diff| %.3g" % (exact, compared, worst)


# forced frame types (pic_in.i_type: what a qpfile or an application's keyframe request sets): 1 IDR, 2 I, 3 P, 4 BREF, 5 B, 6 KEYFRAME
FORCED_CASES = [
    ("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=60", (96, 64), 50, 27, {9: 1, 20: 6, 33: 2, 41: 3}),
    ("medium", "bframes=3:b-adapt=2:rc-lookahead=12", (96, 64), 48, 30, {5: 5, 6: 5, 7: 3, 16: 6, 24: 4, 25: 5, 37: 1}),
    ("medium", "weightp=0:no-psy=1:open-gop=1:bframes=2:rc-lookahead=8:keyint=20", (96, 64), 50, 35, {11: 6, 12: 5, 30: 2}),
    ("medium", "weightp=0:no-mbtree=1:bframes=3:b-adapt=0:rc-lookahead=0:scenecut=0", (64, 48), 40, None, {3: 3, 10: 1, 11: 5, 12: 5, 13: 5, 14: 5, 20: 6}),
]


@pytest.mark.parametrize("case", FORCED_CASES)
def test_forced_frame_types_match_reference_encoder(case):
    preset, opts, (w, h), n, cut, forced_at = case
    frames = synth_sequence(w, h, n, seed=n + w + 9, cut_at=cut)
    forced = [forced_at.get(i, 0) for i in range(n)]
    p, want = reference_types(preset, opts, w, h, frames, forced=forced)
    got = decide_with(slicetype_oracle_lib(), p, frames, forced=forced)
    assert got == want, (case, [x for x in zip(got, want) if x[0] != x[1]][:6])
    assert all(dict(want)[i] in ((1, 2) if t == 6 else (t,)) for i, t in forced_at.items() if t in (1, 6)), "forced keyframes were honoured"


# x264_rc_analyse_slice (slicetype.c:1976-2030) and the VBV lookahead (vbv_lookahead, slicetype.c:1225-1286): the cost and row
# SATDs the reference ENCODER's rate control was given for every picture it called it for, and the planned types / costs it left
# with every non-B picture, against the product's host logic.  With a VBV the B pictures are analysed too, the delay follows
# rc-lookahead, the whole lookahead is analysed every time and MB-tree finishes every referenced picture.
RC_CASES = [
    ("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 44, 17),                       # ABR/CRF, no VBV
    ("medium", "weightp=0:no-psy=1:no-mbtree=1:aq-mode=0:bframes=2:rc-lookahead=8", (96, 64), 40, 21),                        # plain cost
    ("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:vbv-maxrate=400:vbv-bufsize=300", (112, 80), 44, 17),   # VBV + MB-tree
    ("medium", "weightp=0:no-psy=1:no-mbtree=1:bframes=3:b-adapt=2:rc-lookahead=12:vbv-maxrate=400:vbv-bufsize=300", (96, 64), 40, 23),
    ("medium", "bframes=3:rc-lookahead=12:keyint=40:vbv-maxrate=500:vbv-bufsize=400", (112, 80), 44, 19),                       # the default path + VBV
    ("medium", "weightp=0:no-psy=1:bframes=0:rc-lookahead=6:vbv-maxrate=300:vbv-bufsize=200", (64, 64), 30, 11),
]


# more shapes, on the CPU harness only (the device entry points underneath are the same ones the cases above exercise on the GPU)
RC_CASES_HOST = [
    ("medium", "weightp=0:no-psy=1:bframes=3:b-pyramid=strict:rc-lookahead=10:keyint=30:vbv-maxrate=400:vbv-bufsize=300", (112, 80), 44, 17),
    ("medium", "weightp=0:no-psy=1:bframes=4:b-pyramid=normal:b-adapt=2:rc-lookahead=14:vbv-maxrate=400:vbv-bufsize=300", (96, 64), 44, 21),
    ("medium", "weightp=0:no-psy=1:no-mbtree=1:bframes=2:rc-lookahead=0:vbv-maxrate=400:vbv-bufsize=300", (96, 64), 36, 15),   # VBV without a lookahead
    ("medium", "open-gop=1:bframes=3:rc-lookahead=12:keyint=20:vbv-maxrate=500:vbv-bufsize=400", (96, 64), 50, 23),
    ("medium", "weightp=0:bframes=3:rc-lookahead=10:aq-mode=2:vbv-maxrate=500:vbv-bufsize=400", (112, 80), 40, 13),
    ("slow", "rc-lookahead=20:vbv-maxrate=500:vbv-bufsize=400:keyint=40", (96, 64), 50, 27),
    ("medium", "bframes=16:b-adapt=2:rc-lookahead=30:keyint=60:vbv-maxrate=500:vbv-bufsize=400", (64, 48), 70, 33),
]


def rc_compare(want, got, rc_ref, rc_got, vbv):
    assert got == want, [x for x in zip(got, want) if x[0] != x[1]][:6]
    analysed = planned = 0
    for fr, ty in want:
        a, b = rc_ref[fr], rc_got[fr]
        if a[0] >= 0:
            assert a[0] == b[0], ("cost", fr, ty, a[0], b[0])
            analysed += 1
            if vbv:
                assert np.array_equal(a[1], b[1]), ("row satds", fr, ty)
        if vbv and ty not in (4, 5):
            assert a[2] == b[2] and a[3] == b[3], ("planned", fr, ty, a[2], b[2], a[3], b[3])
            planned += len(a[2])
    return analysed, planned


@pytest.mark.parametrize("case", RC_CASES + RC_CASES_HOST)
def test_rc_analyse_slice_and_vbv_lookahead_match_reference_encoder(case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w + 5, cut_at=cut)
    rc_ref, rc_got = {}, {}
    p, want = reference_types(preset, opts, w, h, frames, rc_out=rc_ref)
    # with adaptive quantisation on the pictures go in as I420 (chroma 128, as the reference harness feeds it): AQ runs inside
    chroma = (np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8),) * 2 if p.la.aq_mode else None
    got = decide_with(slicetype_oracle_lib(), p, frames, rc_out=rc_got, chroma=chroma)
    analysed, planned = rc_compare(want, got, rc_ref, rc_got, p.la.vbv)
    assert analysed >= 10 and (planned > 20 or not (p.la.vbv and p.rc_lookahead))


@pytest.mark.parametrize("b_adapt", [0, 1, 2])
@pytest.mark.parametrize("bframes", [2, 3, 5, 8, 16])
def test_nobody_asks_for_list1_beyond_half_a_minigop_under_a_b_pyramid(b_adapt, bframes):
    """what the prefetcher relies on (slicetype.c: note_searches_of): with a B pyramid and no VBV every consumer -- the b-adapt 1 loop,
    the trellis, MB-tree, the rate control's costs -- prices a B picture against the middle picture or the nearer anchor, so no
    request names a later reference further than (bframes+1) - (bframes+1)/2 pictures away.  With the rate control's reads included
    (x264cu_slicetype_rc_analyse_slice of every picture), on sequences with a cut, a still stretch and noise."""
    lib = slicetype_oracle_lib()
    w, h, n = 64, 48, 56 if bframes < 16 else 72
    worst = 0
    for seed, mbtree in ((3, 1), (11, 0)):
        frames = synth_sequence(w, h, n, seed=seed + bframes, cut_at=n // 2 + 3)
        for i in range(8, 16):                       # a still stretch: long runs of B pictures
            frames[i] = frames[8]
        la = LookaheadParams(w, h, 2, 1, 16, 512, bframes, 0, 1, 0, mbtree, 0, 0, 0)
        p = SlicetypeParams(la, 250, 25, 40, b_adapt, 2, 30, 0, 3, 0)
        stats, rc = {}, {}
        types = decide_with(lib, p, frames, rc_out=rc, stats=stats)
        assert sorted(f for f, _ in types) == list(range(n))
        assert any(t in (4, 5) for _, t in types)
        worst = max(worst, stats["farthest_list1"])
    span = bframes + 1
    assert 1 <= worst <= span - span // 2, (worst, span)
