"""GPU parity of x264cu_adaptive_quant_frame (x264_adaptive_quant_frame, encoder/ratecontrol.c:305-420, aq-mode 0 / 1) against
the oracle (pinned to the compiled reference by tests/test_oracle_aq.py) and, where it travelled, the reference itself:
f_qp_offset_aq bit-exact (float), i_inv_qscale_factor, frame statistics; and that its outputs drive the lookahead to the same
costs as host-provided arrays."""
import ctypes as C
import numpy as np
import pytest
import x264_b200 as x
from _libs import oracle, ref, have_ref, ptr
from test_oracle_aq import make_picture, bind

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = x.Context(0)
    c.L.x264cu_adaptive_quant_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int,
                                                C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    yield c
    c.close()


def gpu_aq(ctx, luma, cb, cr, mode, strength):
    h, w = luma.shape
    nmb = ((w + 15) // 16) * ((h + 15) // 16)
    d_l, d_b, d_r = ctx.upload(luma), ctx.upload(cb), ctx.upload(cr)
    d_q, d_i = ctx.malloc(nmb * 4), ctx.malloc(nmb * 2)
    stats = np.zeros(6, np.uint64)
    ctx.check(ctx.L.x264cu_adaptive_quant_frame(ctx.h, d_l, w, d_b, d_r, cb.shape[1], w, h, mode, strength, d_q, d_i, stats.ctypes.data))
    q, i = ctx.download(d_q, (nmb,), np.float32), ctx.download(d_i, (nmb,), np.uint16)
    for p in (d_l, d_b, d_r, d_q, d_i):
        ctx.free(p)
    return q, i, stats


@pytest.mark.parametrize("cfg", [((112, 80), 1, 1.0), ((100, 52), 1, 1.4), ((96, 64), 0, 1.0), ((1918, 1078), 1, 0.6), ((3840, 2160), 1, 1.0),
                                 ((112, 80), 2, 1.0), ((100, 52), 3, 1.3), ((640, 360), 2, 0.8), ((1920, 1080), 3, 1.0)])
def test_adaptive_quant_frame(ctx, cfg):
    (w, h), mode, strength = cfg
    o, _ = bind() if have_ref() else (oracle(), None)
    if not have_ref():
        o.orc_adaptive_quant_frame.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
    luma, cb, cr = make_picture(w, h, seed=w + h)
    q, i, stats = gpu_aq(ctx, luma, cb, cr, mode, strength)
    nmb = q.size
    qo, io, so = np.zeros(nmb, np.float32), np.zeros(nmb, np.uint16), np.zeros(6, np.uint64)
    o.orc_adaptive_quant_frame(ptr(luma), w, ptr(cb), ptr(cr), cb.shape[1], w, h, mode, strength, ptr(qo), ptr(io), ptr(so))
    assert np.array_equal(stats, so), (stats, so)
    assert np.array_equal(i, io), np.argwhere(i != io)[:5]
    assert np.array_equal(q, qo), ("f_qp_offset_aq", float(np.abs(q - qo).max()))
    if have_ref() and w % 2 == 0 and w < 2000:
        r = ref()
        hnd = r.xref_open(w, h, b"medium", ("aq-mode=%d:aq-strength=%g" % (mode, strength)).encode(), 0)
        try:
            qa, ia, sa = np.zeros(nmb, np.float32), np.zeros(nmb, np.uint16), np.zeros(6, np.uint64)
            assert r.xref_aq_frame(hnd, ptr(luma), ptr(cb), ptr(cr), ptr(qa), ptr(ia), ptr(sa)) == 0
            assert np.array_equal(q, qa) and np.array_equal(i, ia) and np.array_equal(stats, sa)
        finally:
            r.xref_close(hnd)
