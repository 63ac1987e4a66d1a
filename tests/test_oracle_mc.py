"""Pins oracle/oracle_mc.c against the compiled reference: lowres init (mc.c:458-507), hpel planes
(mc.c:172-196 + frame.c expand), mc_luma/get_ref (mc.c:198-249), avg (mc.c:77-111), cost_mv (analyse.c:143-188)."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import oracle, ref, have_ref, ptr, PaddedPlane, OrcWeight, synth_luma, PIXEL_W, PIXEL_H, PAD

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")


@pytest.fixture(scope="module")
def libs():
    _libs._bind_mc()
    return oracle(), ref()


@pytest.mark.parametrize("wh", [(64, 48), (100, 52), (176, 144)])
def test_lowres_planes(libs, wh):
    o, r = libs
    w, h = wh
    luma = synth_luma(w, h, seed=w * h, kind="noise")
    hnd = r.xref_open(w, h, b"medium", b"", 0)
    assert hnd
    try:
        st = r.xref_param(hnd, b"stride_lowres")
        wl, ll = r.xref_param(hnd, b"width_lowres"), r.xref_param(hnd, b"lines_lowres")
        mbw, mbh = r.xref_param(hnd, b"mb_width"), r.xref_param(hnd, b"mb_height")
        assert (wl, ll) == (mbw * 8, mbh * 8)
        plane_bytes = st * (ll + 2 * PAD)
        out = np.zeros(4 * plane_bytes, np.uint8)
        assert r.xref_frame_lowres(hnd, ptr(luma), w, ptr(out)) == 0
        # oracle: source = mod16-expanded picture (frame.c:640-665)
        W16, H16 = mbw * 16, mbh * 16
        src = np.zeros((H16, W16), np.uint8)
        src[:h, :w] = luma
        src[:h, w:] = luma[:, w - 1:w]
        src[h:, :] = src[h - 1:h, :]
        planes = [PaddedPlane(wl, ll, stride=st) for _ in range(4)]
        arr = (C.c_void_p * 4)(*[p.buf.ctypes.data + p.origin for p in planes])
        o.orc_frame_init_lowres(ptr(src), W16, W16, H16, arr, st, wl, ll)
        for i in range(4):
            got = planes[i].view()[:, :wl + 2 * PAD]
            want = out[i * plane_bytes:(i + 1) * plane_bytes].reshape(ll + 2 * PAD, st)[:, :wl + 2 * PAD]
            assert np.array_equal(got, want), i
    finally:
        r.xref_close(hnd)


@pytest.mark.parametrize("wh", [(64, 48), (96, 80)])
def test_hpel_planes(libs, wh):
    o, r = libs
    w, h = wh
    luma = synth_luma(w, h, seed=7 + w, kind="noise")
    hnd = r.xref_open(w, h, b"medium", b"", 0)
    try:
        st = r.xref_param(hnd, b"stride")
        plane_bytes = st * (h + 2 * PAD)
        out = np.zeros(3 * plane_bytes, np.uint8)
        assert r.xref_frame_hpel(hnd, ptr(luma), w, ptr(out), None) == 0
        planes = [PaddedPlane(w, h, stride=st) for _ in range(3)]
        o.orc_hpel_filter_plane(ptr(luma), w, w, h, *[ptr(p.buf, p.origin) for p in planes], st, PAD)
        for i in range(3):
            got = planes[i].view()[:, :w + 2 * PAD]
            want = out[i * plane_bytes:(i + 1) * plane_bytes].reshape(h + 2 * PAD, st)[:, :w + 2 * PAD]
            assert np.array_equal(got, want), "HVC"[i]
    finally:
        r.xref_close(hnd)


def test_mc_luma_and_avg(libs):
    o, r = libs
    rng = np.random.default_rng(5)
    st = 128
    planes = [rng.integers(0, 256, st * 96, dtype=np.uint8) for _ in range(4)]
    org = 40 * st + 48
    srcs = (C.c_void_p * 4)(*[p.ctypes.data + org for p in planes])
    for ip in range(7):
        w, h = PIXEL_W[ip], PIXEL_H[ip]
        for mvx in range(-9, 10):
            for mvy in range(-7, 8):
                for wt in (None, (1, 70, 6, -3), (1, 33, 0, 2)):
                    got = np.zeros(32 * 32, np.uint8)
                    want = np.zeros(32 * 32, np.uint8)
                    want2 = np.zeros(32 * 32, np.uint8)
                    wts = OrcWeight(*wt) if wt else OrcWeight(0, 0, 0, 0)
                    o.orc_mc_luma(ptr(got), 32, srcs, st, mvx, mvy, w, h, C.byref(wts))
                    a = [C.c_void_p(p.ctypes.data + org) for p in planes]
                    wa = wt or (0, 0, 0, 0)
                    r.xref_mc_luma(ptr(want), 32, *a, st, mvx, mvy, w, h, *wa)
                    r.xref_get_ref(ptr(want2), 32, *a, st, mvx, mvy, w, h, *wa)
                    assert np.array_equal(got, want) and np.array_equal(got, want2), (ip, mvx, mvy, wt)
    a = rng.integers(0, 256, 32 * 32, dtype=np.uint8)
    b = rng.integers(0, 256, 32 * 32, dtype=np.uint8)
    for ip in range(8):
        for weight in (32, 0, 13, 40, 64, -10, 80):
            got = np.zeros(32 * 32, np.uint8)
            want = np.zeros(32 * 32, np.uint8)
            o.orc_pixel_avg(ptr(got), 32, ptr(a), 32, ptr(b), 32, PIXEL_W[ip], PIXEL_H[ip], weight)
            r.xref_avg(ip, ptr(want), 32, ptr(a), 32, ptr(b), 32, weight)
            assert np.array_equal(got, want), (ip, weight)


def test_cost_mv_table(libs):
    o, r = libs
    hnd = r.xref_open(64, 48, b"medium", b"", 0)
    try:
        mvr = r.xref_param(hnd, b"mvrange")
        n = 2 * 4 * mvr
        want = np.zeros(2 * n + 1, np.uint16)
        got = np.zeros(2 * n + 1, np.uint16)
        r.xref_cost_mv_table(hnd, want, n)
        o.orc_cost_mv_table(got, n, 1)
        assert np.array_equal(got, want)
    finally:
        r.xref_close(hnd)
