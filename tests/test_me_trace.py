"""BASELINE config 3: the x264_me_t stream of the reference ENCODER -- every call its analysis makes to
x264_me_search_ref (encoder/analyse.c:1287, :1392-1785, :1938, :2231-2491: all partition sizes, every reference of both lists,
weighted duplicates, per-macroblock lambdas, chroma ME at --preset slower) -- recorded by oracle/ref_shim.c and replayed
(1) through the oracle on the CPU, which pins oracle/oracle_me.c to the real stream, and (2) through
x264cu_me_search_frame on the GPU: 100 % of the searches must give the reference's (mv, cost, cost_mv, threshold)."""
import numpy as np
import pytest
import _libs
from _libs import have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

CIF_OPTS = b"me=umh:merange=32:ref=3:bframes=2:threads=1"


def check(got, t, T, what):
    want, ee = T.expected(t), T.early_exit(t)
    bad = (got[:, [0, 1, 2, 4]] != want[:, [0, 1, 2, 4]]).any(1) | ((got[:, 3] != want[:, 3]) & ~ee)
    if bad.any():
        i = int(np.argmax(bad))
        raise AssertionError("%s: %d of %d searches of coded picture %d differ; first: %r got %r want %r"
                             % (what, int(bad.sum()), len(got), t.coded, t.recs[i], got[i], want[i]))


def test_oracle_reproduces_the_encoder_me_stream():
    import _me_trace as T
    frames = T.record(352, 288, 6, CIF_OPTS, max_frames=5, skip=1)
    assert len(frames) == 5
    assert {t.slice_type for t in frames} == {0, 1}                      # P and B pictures
    assert all(t.chroma_me and t.subpel == 9 and t.me_method == 2 for t in frames)
    assert max(t.n_refs for t in frames) >= 3
    assert any(rf["weighted"] for t in frames for rf in t.refs) and any(rf["weight"][1][0] for t in frames for rf in t.refs)
    seen = set()
    for t in frames:
        seen |= set(int(v) for v in np.unique(t.recs["i_pixel"]))
        check(T.replay_oracle(t), t, T, "oracle")
    assert seen == set(range(7))                                          # every partition size of the analysis


@pytest.mark.gpu
def test_gpu_replays_the_encoder_me_stream_cif():
    import x264_b200 as x
    import _me_trace as T
    frames = T.record(352, 288, 6, CIF_OPTS, max_frames=5, skip=1)
    with x.Context(0) as ctx:
        for t in frames:
            d = T.DeviceTrace(ctx, t, x)
            try:
                d.launch()
                check(d.results(), t, T, "x264cu_me_search_frame")
            finally:
                d.close()


@pytest.mark.gpu
def test_gpu_replays_the_encoder_me_stream_4k_slower():
    """3840x2160 --preset slower --me umh --merange 64 (BASELINE configs[2]): P and B pictures, every search of each"""
    import x264_b200 as x
    import _me_trace as T
    frames = T.record(3840, 2160, 8, b"me=umh:merange=64:threads=1", max_frames=4, skip=1)
    assert len(frames) == 4
    assert sum(t.slice_type == 0 for t in frames) >= 2, [t.slice_type for t in frames]
    total = 0
    with x.Context(0) as ctx:
        for t in frames:
            d = T.DeviceTrace(ctx, t, x)
            try:
                d.launch()
                check(d.results(), t, T, "x264cu_me_search_frame")
                total += d.n
            finally:
                d.close()
    assert total > 200000
