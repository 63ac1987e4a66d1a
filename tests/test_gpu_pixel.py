"""GPU parity of the batched pixel-table twins (x264cu_pixel_cmp_*) against the oracle -- and, where the compiled
reference travelled with the snapshot, against the reference itself.  Bit-exact (integer metrics).
Mirrors tools/checkasm.c check_pixel (:361-514): random + worst-case buffers, every size, sad/ssd/satd/sa8d, x3/x4."""
import numpy as np
import pytest
import x264_b200 as x
from _libs import oracle, ref, have_ref, cand_dtype as ocand, worst_case_pair, PaddedPlane

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = x.Context(0)
    yield c
    c.close()


def _planes(pattern, n, rng):
    if pattern == "random":
        return rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 256, n, dtype=np.uint8)
    if pattern == "worst":
        return worst_case_pair(n, rng)
    return rng.integers(120, 124, n, dtype=np.uint8), rng.integers(120, 124, n, dtype=np.uint8)


METRICS = [(x.SAD, range(8)), (x.SSD, range(8)), (x.SATD, range(8)), (x.SA8D, [0, 3])]


@pytest.mark.parametrize("pattern", ["random", "worst", "flat"])
def test_cmp_batch_matches_oracle(ctx, pattern):
    rng = np.random.default_rng(99)
    stride, rows = 256, 96
    a, b = _planes(pattern, stride * rows, rng)
    for n in (1, 7, 1000):
        cand = np.zeros(n, x.cand_dtype)
        cand["fenc_off"] = rng.integers(0, rows - 16, n) * stride + rng.integers(0, stride - 16, n)
        cand["ref_off"] = rng.integers(0, rows - 16, n) * stride + rng.integers(0, stride - 17, n)
        for metric, sizes in METRICS:
            for ip in sizes:
                got = ctx.pixel_cmp_batch_host(metric, ip, a, stride, b, stride, cand)
                want = np.zeros(n, np.int32)
                oracle().orc_pixel_cmp_batch(metric, ip, a, stride, b, stride, cand.view(ocand), n, want)
                assert np.array_equal(got, want), (pattern, n, metric, ip)
                if have_ref():
                    r = np.zeros(n, np.int32)
                    ref().xref_pixel_cmp_batch(metric, ip, a, stride, b, stride, cand.view(ocand), n, r)
                    assert np.array_equal(got, r), ("ref", pattern, n, metric, ip)


def test_cmp_batch_empty(ctx):
    a = np.zeros(64 * 64, np.uint8)
    got = ctx.pixel_cmp_batch_host(x.SAD, 0, a, 64, a, 64, np.zeros(0, x.cand_dtype))
    assert got.shape == (0,)


def test_x3_x4(ctx):
    rng = np.random.default_rng(3)
    stride, rows = 128, 64
    a = rng.integers(0, 256, stride * rows, dtype=np.uint8)
    b = rng.integers(0, 256, stride * rows, dtype=np.uint8)
    n = 333
    cand = np.zeros(n, x.cand_x4_dtype)
    cand["fenc_off"] = rng.integers(0, rows - 16, n) * stride + rng.integers(0, 7, n) * 16
    cand["ref_off"] = (rng.integers(0, rows - 16, (n, 4)) * stride + rng.integers(0, stride - 17, (n, 4)))
    da, db, dc = ctx.upload(a), ctx.upload(b), ctx.upload(cand)
    dout = ctx.malloc(n * 16)
    for metric in (x.SAD, x.SATD):
        for ip in range(7):
            for nrefs in (3, 4):
                ctx.pixel_cmp_x4_batch(metric, ip, nrefs, da, stride, db, stride, dc, n, dout)
                got = ctx.download(dout, (n, 4), np.int32)
                for j in range(nrefs):
                    c1 = np.zeros(n, ocand)
                    c1["fenc_off"] = cand["fenc_off"]
                    c1["ref_off"] = cand["ref_off"][:, j]
                    want = np.zeros(n, np.int32)
                    oracle().orc_pixel_cmp_batch(metric, ip, a, stride, b, stride, c1, n, want)
                    assert np.array_equal(got[:, j], want), (metric, ip, nrefs, j)


def _mvfield_case(ctx, w, h, n_planes, k, ip, metric, rng, mv_range=16, pattern="random"):
    pl_f = [PaddedPlane(w, h) for _ in range(n_planes)]
    pl_r = [PaddedPlane(w, h) for _ in range(n_planes)]
    stride = pl_f[0].stride
    pitch = pl_f[0].buf.size
    fenc = np.zeros(pitch * n_planes, np.uint8)
    refp = np.zeros(pitch * n_planes, np.uint8)
    for f in range(n_planes):
        a, b = _planes(pattern, w * h, rng)
        pl_f[f].inner()[:] = a.reshape(h, w)
        pl_r[f].inner()[:] = b.reshape(h, w)
        pl_f[f].fill_border()
        pl_r[f].fill_border()
        fenc[f * pitch:(f + 1) * pitch] = pl_f[f].buf
        refp[f * pitch:(f + 1) * pitch] = pl_r[f].buf
    bx, by = w // x.PIXEL_W[ip], h // x.PIXEL_H[ip]
    mv = rng.integers(-mv_range, mv_range + 1, (k, n_planes, by, bx, 2)).astype(np.int16)
    got = ctx.pixel_cmp_mvfield_host(metric, ip, fenc, refp, stride, pitch, w, h, n_planes, k, mv)
    got = got.reshape(k, n_planes, by * bx)
    org = pl_f[0].origin
    for f in range(n_planes):
        want = np.zeros(k * by * bx, np.int32)
        mvf = np.ascontiguousarray(mv[:, f])
        oracle().orc_pixel_cmp_mvfield(metric, ip, fenc.ctypes.data + f * pitch + org, stride,
                                       refp.ctypes.data + f * pitch + org, stride, bx, by, k, mvf.reshape(-1), want)
        assert np.array_equal(got[:, f].reshape(-1), want), (w, h, f, ip, metric)


@pytest.mark.parametrize("metric", [x.SAD, x.SSD, x.SATD])
def test_mvfield_all_sizes(ctx, metric):
    rng = np.random.default_rng(11)
    for ip in range(8):
        _mvfield_case(ctx, 272, 144, 2, 2, ip, metric, rng)          # ragged against the 128x64 tile


def test_mvfield_sa8d(ctx):
    rng = np.random.default_rng(12)
    for ip in (0, 3):
        _mvfield_case(ctx, 272, 144, 1, 1, ip, x.SA8D, rng)


def test_mvfield_worst_case_and_large_vectors(ctx):
    rng = np.random.default_rng(13)
    _mvfield_case(ctx, 256, 128, 1, 1, 0, x.SATD, rng, pattern="worst")
    _mvfield_case(ctx, 256, 128, 1, 3, 3, x.SATD, rng, mv_range=32)   # beyond the staged halo -> global path
    _mvfield_case(ctx, 64, 48, 3, 1, 6, x.SAD, rng)                     # smaller than one tile


def test_mvfield_full_size_property(ctx):
    """BASELINE size (4K): zero motion against an identical plane -> all costs 0; against plane+1 -> SAD = W*H"""
    w, h = 3840, 2160
    p = PaddedPlane(w, h)
    rng = np.random.default_rng(5)
    p.inner()[:] = rng.integers(0, 255, (h, w), dtype=np.uint8)
    p.fill_border()
    q = p.buf + 1
    mv = np.zeros((h // 16) * (w // 16) * 2, np.int16)
    out = ctx.pixel_cmp_mvfield_host(x.SATD, 0, p.buf, p.buf, p.stride, p.buf.size, w, h, 1, 1, mv)
    assert out.shape[0] == 32400 and not out.any()
    out = ctx.pixel_cmp_mvfield_host(x.SAD, 0, p.buf, q, p.stride, p.buf.size, w, h, 1, 1, mv)
    assert (out == 256).all()
    # linearity check of SATD against the oracle on a random subset of blocks with random motion
    mvr = rng.integers(-16, 17, mv.shape).astype(np.int16)
    r = PaddedPlane(w, h)
    r.inner()[:] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    r.fill_border()
    out = ctx.pixel_cmp_mvfield_host(x.SATD, 0, p.buf, r.buf, p.stride, p.buf.size, w, h, 1, 1, mvr)
    want = np.zeros(32400, np.int32)
    oracle().orc_pixel_cmp_mvfield(x.SATD, 0, p.buf.ctypes.data + p.origin, p.stride, r.buf.ctypes.data + r.origin,
                                   r.stride, w // 16, h // 16, 1, mvr, want)
    assert np.array_equal(out, want)
