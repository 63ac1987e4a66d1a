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


@pytest.mark.parametrize("wh", [(64, 48), (96, 80), (200, 120), (98, 50), (704, 368), (1920, 1080), (3840, 2160)])
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


def test_frame_kernels_over_a_stack_of_pictures(ctx):
    """x264cu_hpel_filter_batch / x264cu_frame_init_lowres_batch: several pictures per launch give each picture's own result"""
    vp, ss, ci = C.c_void_p, C.c_ssize_t, C.c_int
    ctx.L.x264cu_frame_init_lowres_batch.argtypes = [vp, vp, ss, ss, ci, ci, ci, C.POINTER(vp), ss, ss]
    ctx.L.x264cu_hpel_filter_batch.argtypes = [vp, vp, ss, ss, ci, ci, ci, vp, vp, vp, ci]
    w, h, n = 704, 368, 3
    lumas = [synth_luma(w, h, seed=50 + i, kind="noise") for i in range(n)]
    # hpel
    refs = [_libs.make_ref_planes(l) for l in lumas]
    st, nbytes = refs[0][0].stride, refs[0][0].buf.size
    pitch = (nbytes + 255) & ~255
    d = [ctx.malloc(n * pitch + 256) for _ in range(4)]
    for i, l in enumerate(lumas):
        src = PaddedPlane(w, h, stride=st)
        src.inner()[:] = l
        ctx.h2d(d[0] + i * pitch, src.buf)
    org = refs[0][0].origin
    ctx.check(ctx.L.x264cu_hpel_filter_batch(ctx.h, d[0] + org, st, pitch, n, w, h, d[1] + org, d[2] + org, d[3] + org, 1))
    for i in range(n):
        for k in range(4):
            got = ctx.download(d[k] + i * pitch, (h + 2 * PAD, st), np.uint8)
            assert np.array_equal(got[:, :w + 2 * PAD], refs[i][k].view()[:, :w + 2 * PAD]), (i, "FHVC"[k])
    for p_ in d:
        ctx.free(p_)
    # lowres
    wl, ll = w // 2, h // 2
    pl = PaddedPlane(wl, ll)
    plane_bytes = (pl.buf.size + 255) & ~255
    pitch_dst, pitch_src = 4 * plane_bytes, w * h
    d_src, d_pl = ctx.malloc(n * pitch_src + 256), ctx.malloc(n * pitch_dst + 256)
    for i, l in enumerate(lumas):
        ctx.h2d(d_src + i * pitch_src, l)
    darr = (vp * 4)(*[d_pl + k * plane_bytes + pl.origin for k in range(4)])
    ctx.check(ctx.L.x264cu_frame_init_lowres_batch(ctx.h, d_src, w, pitch_src, n, w, h, darr, pl.stride, pitch_dst))
    for i, l in enumerate(lumas):
        want = [PaddedPlane(wl, ll) for _ in range(4)]
        arr = (vp * 4)(*[p_.buf.ctypes.data + p_.origin for p_ in want])
        oracle().orc_frame_init_lowres(ptr(l), w, w, h, arr, pl.stride, wl, ll)
        for k in range(4):
            got = ctx.download(d_pl + i * pitch_dst + k * plane_bytes, (ll + 2 * PAD, pl.stride), np.uint8)
            assert np.array_equal(got[:, :wl + 2 * PAD], want[k].view()[:, :wl + 2 * PAD]), (i, "FHVC"[k])
    ctx.free(d_src)
    ctx.free(d_pl)
