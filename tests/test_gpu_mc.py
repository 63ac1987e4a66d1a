"""GPU parity of the stand-alone mc-table twins (x264cu_mc_luma_batch = mc_luma / get_ref, x264cu_pixel_avg_batch = avg[],
x264cu_weight_scale_plane = weight, common/mc.c:49-249) and of x264cu_pixel_ssd_wxh (common/pixel.c:112-151) against the
oracle (pinned to the compiled reference by tests/test_oracle_mc.py) and, where it travelled, the reference itself."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
import _libs
from _libs import oracle, ref, have_ref, ptr, PaddedPlane, OrcWeight, synth_luma, make_ref_planes, PIXEL_W, PIXEL_H, PAD

pytestmark = pytest.mark.gpu

mc_job_dtype = np.dtype([("src_off", np.uint32), ("mvx", np.int16), ("mvy", np.int16)])


@pytest.fixture(scope="module")
def ctx():
    _libs._bind_mc()
    c = x.Context(0)
    L = c.L
    vp, ci, ss = C.c_void_p, C.c_int, C.c_ssize_t
    L.x264cu_mc_luma_batch.argtypes = [vp, C.POINTER(vp), ss, ci, vp, ci, C.POINTER(ci), vp]
    L.x264cu_pixel_avg_batch.argtypes = [vp, ci, vp, vp, ci, ci, vp]
    L.x264cu_weight_scale_plane.argtypes = [vp, vp, vp, ss, ci, ci, C.POINTER(ci)]
    L.x264cu_pixel_ssd_wxh.argtypes = [vp, vp, ss, vp, ss, ci, ci, C.POINTER(C.c_uint64)]
    yield c
    c.close()


@pytest.mark.parametrize("wt", [None, (1, 70, 6, -3), (1, 33, 0, 2)])
def test_mc_luma_batch(ctx, wt):
    w, h = 160, 96
    luma = synth_luma(w, h, seed=21, kind="noise")
    planes = make_ref_planes(luma)                      # F,H,V,C padded, via the oracle
    st = planes[0].stride
    d_pl = [ctx.upload(p.buf) for p in planes]
    rng = np.random.default_rng(3)
    for ip in range(8):
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        n = 300
        jobs = np.zeros(n, mc_job_dtype)
        bx = rng.integers(0, (w - bw) // 4 + 1, n) * 4
        by = rng.integers(0, (h - bh) // 4 + 1, n) * 4
        jobs["src_off"] = planes[0].origin + by * st + bx
        jobs["mvx"] = rng.integers(-4 * 20, 4 * 20 + 1, n)          # stays inside the 32-pixel border
        jobs["mvy"] = rng.integers(-4 * 20, 4 * 20 + 1, n)
        jobs["mvx"] = np.clip(jobs["mvx"], -4 * (bx + 24), 4 * (w - bx - bw + 24))
        jobs["mvy"] = np.clip(jobs["mvy"], -4 * (by + 24), 4 * (h - by - bh + 24))
        d_jobs = ctx.upload(jobs)
        d_dst = ctx.malloc(n * bw * bh)
        arr = (C.c_void_p * 4)(*d_pl)
        wa = (C.c_int * 4)(*wt) if wt else None
        ctx.check(ctx.L.x264cu_mc_luma_batch(ctx.h, arr, st, ip, d_jobs, n, wa, d_dst))
        got = ctx.download(d_dst, (n, bh, bw), np.uint8)
        wts = OrcWeight(*wt) if wt else OrcWeight(0, 0, 0, 0)
        for k in range(n):
            off = int(jobs["src_off"][k])
            srcs = (C.c_void_p * 4)(*[p.buf.ctypes.data + off for p in planes])
            want = np.zeros((bh, bw), np.uint8)
            oracle().orc_mc_luma(ptr(want), bw, srcs, st, int(jobs["mvx"][k]), int(jobs["mvy"][k]), bw, bh, C.byref(wts))
            assert np.array_equal(got[k], want), (ip, k, jobs[k], wt)
            if have_ref() and k < 40:
                r = np.zeros((bh, bw), np.uint8)
                a = [C.c_void_p(p.buf.ctypes.data + off) for p in planes]
                ref().xref_get_ref(ptr(r), bw, *a, st, int(jobs["mvx"][k]), int(jobs["mvy"][k]), bw, bh, *(wt or (0, 0, 0, 0)))
                assert np.array_equal(got[k], r), ("ref", ip, k)
        ctx.free(d_jobs)
        ctx.free(d_dst)
    for p in d_pl:
        ctx.free(p)
    assert ctx.L.x264cu_mc_luma_batch(ctx.h, (C.c_void_p * 4)(*d_pl), st, 0, None, 0, None, None) == 0      # empty batch


def test_pixel_avg_batch(ctx):
    rng = np.random.default_rng(8)
    for ip in range(8):
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        n = 64
        a = rng.integers(0, 256, (n, bh, bw), dtype=np.uint8)
        b = rng.integers(0, 256, (n, bh, bw), dtype=np.uint8)
        d_a, d_b, d_o = ctx.upload(a), ctx.upload(b), ctx.malloc(a.size)
        for weight in (32, 0, 13, 40, 64, -10, 80):
            ctx.check(ctx.L.x264cu_pixel_avg_batch(ctx.h, ip, d_a, d_b, n, weight, d_o))
            got = ctx.download(d_o, (n, bh, bw), np.uint8)
            for k in range(0, n, 7):
                want = np.zeros((bh, bw), np.uint8)
                oracle().orc_pixel_avg(ptr(want), bw, ptr(a[k]), bw, ptr(b[k]), bw, bw, bh, weight)
                assert np.array_equal(got[k], want), (ip, weight, k)
                if have_ref():
                    r = np.zeros((bh, bw), np.uint8)
                    ref().xref_avg(ip, ptr(r), bw, ptr(a[k]), bw, ptr(b[k]), bw, weight)
                    assert np.array_equal(got[k], r), ("ref", ip, weight, k)
        for p in (d_a, d_b, d_o):
            ctx.free(p)


@pytest.mark.parametrize("wt", [(1, 70, 6, -3), (1, 33, 0, 2), (1, 127, 7, -128), (1, 5, 2, 127)])
def test_weight_scale_plane(ctx, wt):
    w, h = 200, 72
    src = PaddedPlane(w, h)
    src.buf[:] = np.random.default_rng(5).integers(0, 256, src.buf.size, dtype=np.uint8)
    st = src.stride
    want = np.zeros_like(src.buf)
    ow = OrcWeight(*wt)
    oracle().orc_weight_scale_plane(ptr(want), st, ptr(src.buf), st, st, h + 2 * PAD, C.byref(ow))
    d_s, d_d = ctx.upload(src.buf), ctx.malloc(src.buf.size)
    ctx.check(ctx.L.x264cu_weight_scale_plane(ctx.h, d_s, d_d, st, st, h + 2 * PAD, (C.c_int * 4)(*wt)))
    got = ctx.download(d_d, src.buf.shape, np.uint8)
    assert np.array_equal(got, want)
    # odd geometry: unaligned start, width not a multiple of 4
    ctx.check(ctx.L.x264cu_memset(ctx.h, d_d, 0, src.buf.size))
    ctx.check(ctx.L.x264cu_weight_scale_plane(ctx.h, d_s + 3, d_d + 3, st, 77, 9, (C.c_int * 4)(*wt)))
    got = ctx.download(d_d, (h + 2 * PAD, st), np.uint8)
    w2 = want.reshape(h + 2 * PAD, st)
    assert np.array_equal(got[:9, 3:80], w2[:9, 3:80]) and not got[:9, 80:].any() and not got[9:].any()
    ctx.free(d_s)
    ctx.free(d_d)


@pytest.mark.parametrize("wh", [(64, 48), (101, 37), (3840, 2160), (1, 1)])
def test_pixel_ssd_wxh(ctx, wh):
    w, h = wh
    rng = np.random.default_rng(w + h)
    sa, sb = w + 13, w + 32
    a = rng.integers(0, 256, (h, sa), dtype=np.uint8)
    b = rng.integers(0, 256, (h, sb), dtype=np.uint8)
    if w == 3840:                      # worst case: every difference 255 -> 3840*2160*65025 needs 40 bits
        a[:] = 255
        b[:] = 0
    want = int(((a[:, :w].astype(np.int64) - b[:, :w].astype(np.int64)) ** 2).sum())
    d_a, d_b = ctx.upload(a), ctx.upload(b)
    out = C.c_uint64()
    ctx.check(ctx.L.x264cu_pixel_ssd_wxh(ctx.h, d_a, sa, d_b, sb, w, h, C.byref(out)))
    assert out.value == want
    ctx.check(ctx.L.x264cu_pixel_ssd_wxh(ctx.h, d_a + 1, sa, d_b + 2, sb, max(w - 2, 0), h, C.byref(out)))
    assert out.value == int(((a[:, 1:w - 1].astype(np.int64) - b[:, 2:w].astype(np.int64)) ** 2).sum())
    ctx.free(d_a)
    ctx.free(d_b)
