"""Successive elimination: pixf.ads[] (x264_pixel_ads1 / 2 / 4, common/pixel.c:759-803) and the integral planes
x264_frame_filter builds with integral_init4h / 8h / 4v / 8v (common/mc.c:424-456, :748-783).  The oracle's restatements are pinned
to the compiled reference (CPU); x264cu_integral_init and x264cu_pixel_ads_batch are compared with the oracle (GPU)."""
import ctypes as C
import numpy as np
import pytest
import _libs
from _libs import oracle, ref, have_ref, PaddedPlane, synth_luma

PAD = 32
ADS_K = {0: 4, 1: 2, 2: 2, 3: 1, 4: 2, 5: 2, 6: 1}          # pixf.ads[i_pixel]: pixel.c:860-862, :1660-1663


def _bind():
    o = oracle()
    vp, ci, ss = C.c_void_p, C.c_int, C.c_ssize_t
    o.orc_pixel_ads.argtypes = [ci, vp, vp, ci, vp, vp, ci, ci]
    o.orc_integral_init.argtypes = [vp, ss, ci, ci, ci, vp, vp]
    return o


def oracle_integral(luma, sub):
    """-> (plane, sum8, sum4 or None) as PaddedPlane-shaped arrays (u16 planes share the pixel plane's stride in elements)"""
    o = _bind()
    h, w = luma.shape
    pl = PaddedPlane(w, h)
    pl.inner()[:] = luma
    pl.fill_border()
    n = pl.buf.size + 16 * pl.stride
    s8 = np.zeros(n, np.uint16)
    s4 = np.zeros(n, np.uint16) if sub else None
    o.orc_integral_init(pl.buf.ctypes.data + pl.origin, pl.stride, w, h, PAD, s8.ctypes.data + 2 * pl.origin,
                        s4.ctypes.data + 2 * pl.origin if sub else None)
    return pl, s8, s4


def box(pl, w, h, n):
    """plain n x n box sums over the padded plane: [y + PAD, x + PAD] = sum of the box with its corner at (x, y)"""
    img = pl.buf[:pl.stride * (h + 2 * PAD)].reshape(h + 2 * PAD, pl.stride)[:, :w + 2 * PAD].astype(np.int64)
    c = np.zeros((img.shape[0] + 1, img.shape[1] + 1), np.int64)
    c[1:, 1:] = img.cumsum(0).cumsum(1)
    return c[n:, n:] - c[:-n, n:] - c[n:, :-n] + c[:-n, :-n]


def region(arr, pl, w, h, n, y_hi):
    """the part of a u16 plane every implementation defines: corners x in [-PAD, w+PAD-n], y in [-PAD, y_hi]"""
    a = arr[:pl.stride * (h + 2 * PAD)].reshape(h + 2 * PAD, pl.stride)
    return a[:y_hi + PAD + 1, :w + 2 * PAD - n + 1]


@pytest.mark.skipif(not have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("sub", [0, 1])
def test_oracle_integral_matches_reference(sub):
    r = ref()
    r.xref_frame_integral.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    w, h = 112, 80
    hnd = r.xref_open(w, h, b"medium", b"me=esa:partitions=all" if sub else b"me=esa:partitions=none", 0)
    assert hnd
    try:
        luma = synth_luma(w, h, seed=5 + sub, kind="noise")
        st_ref = r.xref_param(hnd, b"stride")
        n = st_ref * (h + 2 * PAD)
        r8, r4, st = np.zeros(n, np.uint16), np.zeros(n, np.uint16), C.c_int()
        got_sub = r.xref_frame_integral(hnd, luma.ctypes.data, w, r8.ctypes.data, r4.ctypes.data, C.byref(st))
        assert got_sub == sub and st.value == st_ref
        pl, o8, o4 = oracle_integral(luma, sub)
        # what x264_frame_filter defines (mc.c:748-783): rows -PADV+1 .. height+PADV-17 (the row at -PADV is the zeroed prefix
        # row), columns up to stride-8 counted from -PADH_ALIGN = -64
        y_hi = h + PAD - 17
        x_n = min(w + 2 * PAD - 7, st_ref - 8 - 64 + PAD)
        ref8 = r8.reshape(h + 2 * PAD, st_ref)[1:y_hi + PAD + 1, :x_n]
        assert np.array_equal(region(o8, pl, w, h, 8, y_hi)[1:, :x_n], ref8)
        assert np.array_equal(ref8, box(pl, w, h, 8)[1:y_hi + PAD + 1, :x_n].astype(np.uint16))
        if sub:
            ref4 = r4.reshape(h + 2 * PAD, st_ref)[1:y_hi + PAD + 1, :x_n]
            assert np.array_equal(region(o4, pl, w, h, 4, y_hi)[1:, :x_n], ref4)
            assert np.array_equal(ref4, box(pl, w, h, 4)[1:y_hi + PAD + 1, :x_n].astype(np.uint16))
    finally:
        r.xref_close(hnd)


@pytest.mark.skipif(not have_ref(), reason="compiled reference not present")
def test_oracle_ads_matches_reference():
    o, r = _bind(), ref()
    r.xref_pixel_ads.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    rng = np.random.default_rng(9)
    for ip, k in ADS_K.items():
        for _ in range(40):
            stride = 256
            sums = rng.integers(0, 16321, 20 * stride).astype(np.uint16)
            width = int(rng.integers(1, 17)) * 4
            delta = int(rng.choice([8, 4, 8 * stride, 4 * stride]))
            enc = rng.integers(0, 16321, 4).astype(np.int32)
            cost = rng.integers(0, 200, width + 8).astype(np.uint16)
            thresh = int(rng.integers(100, 30000))
            a, b = np.zeros(width + 8, np.int16), np.zeros(width + 8, np.int16)
            na = r.xref_pixel_ads(ip, enc.ctypes.data, sums.ctypes.data + 2 * 64, delta, cost.ctypes.data, a.ctypes.data, width, thresh)
            nb = o.orc_pixel_ads(k, enc.ctypes.data, sums.ctypes.data + 2 * 64, delta, cost.ctypes.data, b.ctypes.data, width, thresh)
            assert na == nb and np.array_equal(a[:na], b[:nb]), (ip, width, delta, thresh)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(112, 80), (352, 288), (1920, 1088)])
def test_gpu_integral_matches_oracle(size):
    import x264_b200 as x
    w, h = size
    luma = synth_luma(w, h, seed=w, kind="noise")
    pl, o8, o4 = oracle_integral(luma, 1)
    with x.Context(0) as ctx:
        ctx.L.x264cu_integral_init.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        d_pl = ctx.upload(pl.buf)
        n = pl.stride * (h + 2 * PAD)
        d8, d4 = ctx.malloc(2 * n + 64), ctx.malloc(2 * n + 64)
        ctx.check(ctx.L.x264cu_integral_init(ctx.h, d_pl + pl.origin, pl.stride, w, h, d8 + 2 * pl.origin, d4 + 2 * pl.origin))
        g8, g4 = ctx.download(d8, (n,), np.uint16), ctx.download(d4, (n,), np.uint16)
    y_all8, y_all4 = h + PAD - 8, h + PAD - 4
    assert np.array_equal(region(g8, pl, w, h, 8, y_all8), box(pl, w, h, 8).astype(np.uint16))          # every position the box fits
    assert np.array_equal(region(g4, pl, w, h, 4, y_all4), box(pl, w, h, 4).astype(np.uint16))
    y_hi = h + PAD - 17                                             # where the reference's passes define the planes (rows from -PADV+1)
    assert np.array_equal(region(g8, pl, w, h, 8, y_hi)[1:], region(o8, pl, w, h, 8, y_hi)[1:])
    assert np.array_equal(region(g4, pl, w, h, 4, y_hi)[1:, :w + 2 * PAD - 7], region(o4, pl, w, h, 4, y_hi)[1:, :w + 2 * PAD - 7])


@pytest.mark.gpu
def test_gpu_ads_batch_matches_oracle():
    import x264_b200 as x
    o = _bind()
    rng = np.random.default_rng(10)
    job_t = np.dtype([("enc_dc", np.int32, (4,)), ("sums_off", np.uint32), ("delta", np.int32), ("cost_off", np.uint32),
                      ("width", np.int32), ("thresh", np.int32), ("out_off", np.uint32)])
    assert job_t.itemsize == 40
    stride, rows = 512, 160
    sums = rng.integers(0, 16321, rows * stride).astype(np.uint16)
    cost = rng.integers(0, 300, 4096).astype(np.uint16)
    with x.Context(0) as ctx:
        ctx.L.x264cu_pixel_ads_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        d_sums, d_cost = ctx.upload(sums), ctx.upload(cost)
        for ip, k in ADS_K.items():
            n = 300
            jobs = np.zeros(n, job_t)
            out_off = 0
            want = []
            for j in jobs:
                width = int(rng.integers(1, 40)) * 4
                delta = int(rng.choice([8, 4, 8 * stride, 4 * stride]))
                j["enc_dc"] = rng.integers(0, 16321, 4)
                j["sums_off"] = int(rng.integers(0, (rows - 9) * stride - 200))
                j["delta"], j["cost_off"], j["width"] = delta, int(rng.integers(0, 4096 - 200)), width
                j["thresh"], j["out_off"] = int(rng.integers(100, 40000)), out_off
                out_off += width
                mv = np.zeros(width, np.int16)
                enc = np.ascontiguousarray(j["enc_dc"])
                cnt = o.orc_pixel_ads(k, enc.ctypes.data, sums.ctypes.data + 2 * int(j["sums_off"]), delta,
                                      cost.ctypes.data + 2 * int(j["cost_off"]), mv.ctypes.data, width, int(j["thresh"]))
                want.append(mv[:cnt].copy())
            d_jobs, d_cnt, d_mvs = ctx.upload(jobs), ctx.malloc(4 * n), ctx.malloc(2 * out_off + 64)
            ctx.check(ctx.L.x264cu_pixel_ads_batch(ctx.h, ip, d_sums, d_cost, d_jobs, n, d_cnt, d_mvs))
            cnt, mvs = ctx.download(d_cnt, (n,), np.int32), ctx.download(d_mvs, (out_off,), np.int16)
            for i in range(n):
                assert cnt[i] == len(want[i]) and np.array_equal(mvs[int(jobs[i]["out_off"]):int(jobs[i]["out_off"]) + cnt[i]], want[i]), (ip, i)
            assert sum(len(wv) for wv in want) > 0
            for d in (d_jobs, d_cnt, d_mvs):
                ctx.free(d)
