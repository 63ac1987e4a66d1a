"""GPU parity of the frame-preparation twins (x264cu_frame_init_lowres, x264cu_hpel_filter) against the oracle
(pinned to mc.c:458-507 / mc.c:172-196 + frame.c border expansion in tests/test_oracle_mc.py).  Bit-exact."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ptr, PaddedPlane, synth_luma, PAD

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_mc()
    c = x.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("wh", [(64, 48), (100, 52), (176, 144), (1920, 1080), (3840, 2160)])
def test_frame_init_lowres(ctx, wh):
    w, h = wh
    luma = synth_luma(w, h, seed=w + h, kind="noise")
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    wl, ll = mbw * 8, mbh * 8
    planes = [PaddedPlane(wl, ll) for _ in range(4)]
    st = planes[0].stride
    # oracle on the mod-16 expanded picture
    W16, H16 = mbw * 16, mbh * 16
    src = np.zeros((H16, W16), np.uint8)
    src[:h, :w] = luma
    src[:h, w:] = luma[:, w - 1:w]
    src[h:, :] = src[h - 1:h, :]
    arr = (C.c_void_p * 4)(*[p.buf.ctypes.data + p.origin for p in planes])
    oracle().orc_frame_init_lowres(ptr(src), W16, W16, H16, arr, st, wl, ll)
    # device
    stride_src = (w + 63) // 64 * 64
    host = np.zeros((h, stride_src), np.uint8)
    host[:, :w] = luma
    d_src = ctx.upload(host)
    plane_bytes = planes[0].buf.size
    d_planes = ctx.malloc(4 * plane_bytes + 256)
    ctx.check(ctx.L.x264cu_memset(ctx.h, d_planes, 0, 4 * plane_bytes))
    darr = (C.c_void_p * 4)(*[d_planes + i * plane_bytes + planes[0].origin for i in range(4)])
    ctx.check(ctx.L.x264cu_frame_init_lowres(ctx.h, d_src, stride_src, w, h, darr, st))
    got = ctx.download(d_planes, (4, ll + 2 * PAD, st), np.uint8)
    for i in range(4):
        assert np.array_equal(got[i][:, :wl + 2 * PAD], planes[i].view()[:, :wl + 2 * PAD]), "FHVC"[i]
    ctx.free(d_src)
    ctx.free(d_planes)


@pytest.mark.parametrize("wh", [(64, 48), (96, 80), (200, 120)])
def test_hpel_filter(ctx, wh):
    w, h = wh
    luma = synth_luma(w, h, seed=3 * w + h, kind="noise")
    ref_planes = _libs.make_ref_planes(luma)          # F (border filled), H, V, C from the oracle
    st = ref_planes[0].stride
    nbytes = ref_planes[0].buf.size
    src = PaddedPlane(w, h, stride=st)
    src.inner()[:] = luma                              # border left at zero: the kernel must fill it
    d = [ctx.upload(src.buf)] + [ctx.malloc(nbytes + 256) for _ in range(3)]
    org = src.origin
    ctx.check(ctx.L.x264cu_hpel_filter(ctx.h, d[0] + org, st, w, h, d[1] + org, d[2] + org, d[3] + org, 1))
    for i in range(4):
        got = ctx.download(d[i], (h + 2 * PAD, st), np.uint8)
        assert np.array_equal(got[:, :w + 2 * PAD], ref_planes[i].view()[:, :w + 2 * PAD]), "FHVC"[i]
    for p in d:
        ctx.free(p)
