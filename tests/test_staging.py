"""Input staging: x264_frame_copy_picture (common/frame.c:363-480) and the plane-copy entries of the mc table
(common/mc.c:294-339) for the 8-bit 4:2:0 colour spaces.  The oracle is pinned to the compiled reference (CPU); the CUDA twins
(x264cu_frame_copy_picture, x264cu_plane_copy_*) are compared with the oracle (GPU)."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import oracle, ref, have_ref

CSP = {"i420": 2, "yv12": 3, "nv12": 4, "nv21": 5}
VFLIP = 0x1000
SIZES = [(64, 48), (130, 70), (322, 194)]          # even sizes: 4:2:0


def picture(csp, w, h, rng, pad):
    """planes of a random picture in colour space csp, with row padding `pad`; -> (arrays, strides)"""
    cw, ch = w // 2, h // 2
    y = rng.integers(0, 256, (h, w + pad), dtype=np.uint8)
    if csp in (CSP["nv12"], CSP["nv21"]):
        uv = rng.integers(0, 256, (ch, 2 * cw + pad), dtype=np.uint8)
        return [y, uv, None], [w + pad, 2 * cw + pad, 0]
    u = rng.integers(0, 256, (ch, cw + pad), dtype=np.uint8)
    v = rng.integers(0, 256, (ch, cw + pad + 3), dtype=np.uint8)
    return [y, u, v], [w + pad, cw + pad, cw + pad + 3]


def oracle_copy(csp, planes, strides, w, h):
    o = oracle()
    o.orc_frame_copy_picture.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_ssize_t,
                                         C.c_void_p, C.c_ssize_t]
    luma, uv = np.zeros((h, w), np.uint8), np.zeros((h // 2, w), np.uint8)
    pp = (C.c_void_p * 3)(*[p.ctypes.data if p is not None else None for p in planes])
    ss = (C.c_int * 3)(*strides)
    assert o.orc_frame_copy_picture(csp, pp, ss, w, h, luma.ctypes.data, w, uv.ctypes.data, w) == 0
    return luma, uv


@pytest.mark.skipif(not have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("name", sorted(CSP))
@pytest.mark.parametrize("flip", [0, VFLIP])
def test_oracle_frame_copy_picture_matches_reference(name, flip):
    r = ref()
    r.xref_frame_copy_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(len(name) + flip)
    for w, h in SIZES:
        hnd = r.xref_open(w, h, b"medium", b"", 0)
        assert hnd
        try:
            planes, strides = picture(CSP[name], w, h, rng, pad=5)
            luma, uv = np.zeros((h, w), np.uint8), np.zeros((h // 2, w), np.uint8)
            pp = (C.c_void_p * 3)(*[p.ctypes.data if p is not None else None for p in planes])
            ss = (C.c_int * 3)(*strides)
            assert r.xref_frame_copy_picture(hnd, CSP[name] | flip, pp, ss, luma.ctypes.data, uv.ctypes.data) == 0
            ol, ouv = oracle_copy(CSP[name] | flip, planes, strides, w, h)
            assert np.array_equal(ol, luma) and np.array_equal(ouv[:, :2 * (w // 2)], uv[:, :2 * (w // 2)]), (name, flip, w, h)
        finally:
            r.xref_close(hnd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CSP))
@pytest.mark.parametrize("flip", [0, VFLIP])
def test_gpu_frame_copy_picture_matches_oracle(name, flip):
    import x264_b200 as x
    rng = np.random.default_rng(7 + len(name) + flip)
    with x.Context(0) as ctx:
        L = ctx.L
        L.x264cu_frame_copy_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                                                C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
        for (w, h), (sl, sc) in zip(SIZES + [(3840, 2160)], [(64, 64), (192, 160), (384, 328), (3904, 3904)]):
            planes, strides = picture(CSP[name], w, h, rng, pad=5)
            d_l, d_c = ctx.malloc(sl * h + 64), ctx.malloc(sc * (h // 2) + 64)
            pp = (C.c_void_p * 3)(*[p.ctypes.data if p is not None else None for p in planes])
            ss = (C.c_int * 3)(*strides)
            ctx.check(L.x264cu_frame_copy_picture(ctx.h, CSP[name] | flip, pp, ss, w, h, d_l, sl, d_c, sc))
            got_l = ctx.download(d_l, (h, sl), np.uint8)[:, :w]
            got_c = ctx.download(d_c, (h // 2, sc), np.uint8)[:, :2 * (w // 2)]
            ol, ouv = oracle_copy(CSP[name] | flip, planes, strides, w, h)
            assert np.array_equal(got_l, ol) and np.array_equal(got_c, ouv[:, :2 * (w // 2)]), (name, flip, w, h)
            ctx.free(d_l)
            ctx.free(d_c)


@pytest.mark.gpu
def test_gpu_plane_copy_twins_match_oracle():
    import x264_b200 as x
    o = oracle()
    vp, ss, ci = C.c_void_p, C.c_ssize_t, C.c_int
    o.orc_plane_copy_deinterleave.argtypes = [vp, ss, vp, ss, vp, ss, ci, ci]
    rng = np.random.default_rng(5)
    with x.Context(0) as ctx:
        L = ctx.L
        L.x264cu_plane_copy_deinterleave.argtypes = [vp, vp, ss, vp, ss, vp, ss, ci, ci]
        L.x264cu_plane_copy_interleave.argtypes = [vp, vp, ss, vp, ss, vp, ss, ci, ci]
        L.x264cu_plane_copy_swap.argtypes = [vp, vp, ss, vp, ss, ci, ci]
        for w, h, off in [(33, 9, 0), (64, 16, 0), (1920, 540, 0), (50, 7, 1)]:       # off = 1: unaligned planes take the byte path
            st = 2 * w + 24
            src = rng.integers(0, 256, (h, st), dtype=np.uint8)
            d_src = ctx.upload(src) if not off else ctx.upload(np.concatenate([[0], src.reshape(-1)]).astype(np.uint8)) + 1
            d_a, d_b, d_i, d_s = ctx.malloc(st * h + 8), ctx.malloc(st * h + 8), ctx.malloc(st * h + 8), ctx.malloc(st * h + 8)
            ctx.check(L.x264cu_plane_copy_deinterleave(ctx.h, d_a + off, st, d_b + off, st, d_src, st, w, h))
            a, b = np.zeros((h, st), np.uint8), np.zeros((h, st), np.uint8)
            o.orc_plane_copy_deinterleave(a.ctypes.data, st, b.ctypes.data, st, src.ctypes.data, st, w, h)
            ga = ctx.download(d_a + off, (h, st), np.uint8)[:, :w]
            gb = ctx.download(d_b + off, (h, st), np.uint8)[:, :w]
            assert np.array_equal(ga, a[:, :w]) and np.array_equal(gb, b[:, :w]), (w, h, off)
            # interleaving the halves again gives the source back; swapping twice too
            ctx.check(L.x264cu_plane_copy_interleave(ctx.h, d_i, st, d_a + off, st, d_b + off, st, w, h))
            assert np.array_equal(ctx.download(d_i, (h, st), np.uint8)[:, :2 * w], src[:, :2 * w])
            ctx.check(L.x264cu_plane_copy_swap(ctx.h, d_s, st, d_src, st, w, h))
            sw = ctx.download(d_s, (h, st), np.uint8)[:, :2 * w]
            assert np.array_equal(sw[:, 0::2], src[:, 1:2 * w:2]) and np.array_equal(sw[:, 1::2], src[:, 0:2 * w:2])
