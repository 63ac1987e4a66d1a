"""End-to-end parity of the slice-type decisions: raw luma -> x264cu_slicetype_step (CUDA lookahead underneath) versus the
frame types the reference ENCODER itself outputs for the same pictures (oracle/_ref travels with the snapshot), or, if it
did not travel, versus the same host logic running on the CPU oracle.  Identical decisions, frame by frame, coded order."""
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import have_ref, synth_sequence, slicetype_oracle_lib
import test_slicetype_host as host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = x.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", host.CASES + [("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=20", (640, 360), 48, 23)])
def test_gpu_frame_types(ctx, case):
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w, cut_at=cut)
    if cut is not None and n > cut + 9:
        frames[cut + 7] = np.full_like(frames[0], 235)
        frames[cut + 8] = np.full_like(frames[0], 235)
    if cut is not None:
        for i in range(min(10, cut)):
            frames[i] = np.clip(frames[i].astype(np.float32) * (0.35 + 0.065 * i) + 2 * i, 0, 255).astype(np.uint8)
    if not have_ref():
        pytest.skip("compiled reference did not travel")
    p, want = host.reference_types(preset, opts, w, h, frames)
    st = x.Slicetype.from_params(ctx, p)
    try:
        got = st.decide(frames)
    finally:
        st.close()
    assert got == want, (case, [z for z in zip(got, want) if z[0] != z[1]][:6])


@pytest.mark.parametrize("mode", [(0, 0), (1, 0), (1, 5), (0, 16), (1, 24), (1, 32)])
def test_prefetch_and_run_ahead_do_not_change_decisions(ctx, mode):
    prefetch, run_ahead = mode
    w, h, n = 320, 192, 60
    frames = synth_sequence(w, h, n, seed=5, cut_at=31)
    outs = []
    for pf, ra in ((0, 0), (prefetch, run_ahead)):
        st = x.Slicetype(ctx, w, h, rc_lookahead=20, psy=0, aq_mode=0)
        st.set_prefetch(pf)
        st.set_run_ahead(ra)
        outs.append(st.decide(frames))
        st.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("async_upload", [0, 1])
def test_page_locked_pictures_are_read_in_place(ctx, async_upload):
    """x264cu_lookahead_frame_put from page-locked memory (copied by the DMA engine on the upload stream, no staging copy;
    with async_upload the call does not wait for it) gives the decisions of the staged path"""
    w, h, n = 320, 192, 48
    frames = synth_sequence(w, h, n, seed=9, cut_at=20)
    st = x.Slicetype(ctx, w, h, rc_lookahead=20, psy=0, aq_mode=0)
    want = st.decide(frames)
    st.close()
    pinned = ctx.malloc_host(n * w * h).reshape(n, h, w)
    for i in range(n):
        pinned[i] = frames[i]
    st = x.Slicetype(ctx, w, h, rc_lookahead=20, psy=0, aq_mode=0)
    st.set_async_upload(async_upload)
    got = st.decide([pinned[i] for i in range(n)])
    st.close()
    assert got == want


@pytest.mark.parametrize("mode", [(1, 0), (1, 24)])
def test_prefetch_with_weighted_prediction(ctx, mode):
    """with the lookahead's weight analysis on, list-0 searches are run ahead of time only for pairs whose analysis ends at
    its early exit (x264cu_lookahead_weight_trivial); a fade (weights chosen) and a cut in the sequence: same decisions and
    same per-picture weights as the on-demand path"""
    prefetch, run_ahead = mode
    w, h, n = 320, 192, 60
    frames = synth_sequence(w, h, n, seed=11, cut_at=40)
    for i in range(12):                      # fade-in: weights are chosen here
        frames[i] = np.clip(frames[i].astype(np.float32) * (0.3 + 0.06 * i) + 3 * i, 0, 255).astype(np.uint8)
    outs = []
    for pf, ra in ((0, 0), (prefetch, run_ahead)):
        st = x.Slicetype(ctx, w, h, rc_lookahead=20, psy=1, aq_mode=0, weighted_pred=1)
        st.set_prefetch(pf)
        st.set_run_ahead(ra)
        outs.append(st.decide(frames))
        st.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("case", host.FORCED_CASES)
def test_gpu_forced_frame_types(ctx, case):
    """pic_in.i_type forced for some pictures (IDR / I / P / BREF / B / KEYFRAME): the decisions of the reference encoder"""
    preset, opts, (w, h), n, cut, forced_at = case
    frames = synth_sequence(w, h, n, seed=n + w + 9, cut_at=cut)
    forced = [forced_at.get(i, 0) for i in range(n)]
    if not have_ref():
        pytest.skip("compiled reference did not travel")
    p, want = host.reference_types(preset, opts, w, h, frames, forced=forced)
    st = x.Slicetype.from_params(ctx, p)
    try:
        got = st.decide(frames, forced=forced)
    finally:
        st.close()
    assert got == want, (case, [z for z in zip(got, want) if z[0] != z[1]][:6])


@pytest.mark.parametrize("case", host.RC_CASES)
def test_gpu_rc_analyse_slice_and_vbv_lookahead(ctx, case):
    """x264_rc_analyse_slice's cost / row SATDs (slicetype_frame_cost_recalculate with MB-tree) and the VBV lookahead's planned
    types / costs: what the reference ENCODER's rate control was given, picture by picture"""
    preset, opts, (w, h), n, cut = case
    frames = synth_sequence(w, h, n, seed=n + w + 5, cut_at=cut)
    if not have_ref():
        pytest.skip("compiled reference did not travel")
    rc_ref, rc_got = {}, {}
    p, want = host.reference_types(preset, opts, w, h, frames, rc_out=rc_ref)
    chroma = [(np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8),) * 2] * n if p.la.aq_mode else None
    st = x.Slicetype.from_params(ctx, p)
    try:
        got = st.decide(frames, chroma=chroma, rc_out=rc_got, vbv=bool(p.la.vbv))
    finally:
        st.close()
    analysed, planned = host.rc_compare(want, got, rc_ref, rc_got, p.la.vbv)
    assert analysed >= 10 and (planned > 20 or not p.la.vbv)
