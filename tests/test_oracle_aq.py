"""Pins the oracle's adaptive-quantisation restatement (oracle/oracle_aq.c) against the compiled reference's
x264_adaptive_quant_frame (encoder/ratecontrol.c:305-420), aq-mode 0 and 1: per-MB f_qp_offset_aq (float, bit-exact against this
build), i_inv_qscale_factor and the frame statistics the lookahead weight analysis reads."""
import ctypes as C
import numpy as np
import pytest
from _libs import oracle, ref, have_ref, ptr, synth_luma

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")


def make_picture(w, h, seed):
    rng = np.random.default_rng(seed)
    luma = synth_luma(w, h, seed=seed)
    luma[: h // 3] = rng.integers(0, 256, (h // 3, w), dtype=np.uint8)          # a noisy band: large energies
    luma[h // 3: h // 2, : w // 2] = 77                                          # a flat area: energy 0 -> log2(1)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    cb = rng.integers(100, 156, (ch, cw), dtype=np.uint8)
    cr = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    cb[ch // 2:] = 128
    return np.ascontiguousarray(luma), cb, cr


def bind():
    o, r = oracle(), ref()
    o.orc_adaptive_quant_frame.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    r.xref_aq_frame.argtypes = [C.c_void_p] * 7
    return o, r


@pytest.mark.parametrize("cfg", [((112, 80), "aq-mode=1"), ((100, 52), "aq-mode=1:aq-strength=1.4"), ((96, 64), "aq-mode=0"),
                                 ((640, 360), "aq-mode=1:aq-strength=0.6"), ((112, 80), "aq-mode=2"), ((100, 52), "aq-mode=3:aq-strength=1.3"),
                                 ((640, 360), "aq-mode=2:aq-strength=0.8"), ((640, 360), "aq-mode=3")])
def test_adaptive_quant_matches_reference(cfg):
    (w, h), opts = cfg
    o, r = bind()
    luma, cb, cr = make_picture(w, h, seed=w + h)
    hnd = r.xref_open(w, h, b"medium", opts.encode(), 0)
    assert hnd
    try:
        nmb = ((w + 15) // 16) * ((h + 15) // 16)
        qa, qb = np.zeros(nmb, np.float32), np.zeros(nmb, np.float32)
        ia, ib = np.zeros(nmb, np.uint16), np.zeros(nmb, np.uint16)
        sa, sb = np.zeros(6, np.uint64), np.zeros(6, np.uint64)
        assert r.xref_aq_frame(hnd, ptr(luma), ptr(cb), ptr(cr), ptr(qa), ptr(ia), ptr(sa)) == 0
        mode = int(opts.split("aq-mode=")[1][0])
        strength = float(opts.split("aq-strength=")[1]) if "aq-strength" in opts else 1.0
        o.orc_adaptive_quant_frame(ptr(luma), w, ptr(cb), ptr(cr), (w + 1) // 2, w, h, mode, strength, ptr(qb), ptr(ib), ptr(sb))
        assert np.array_equal(sa, sb), (sa, sb)
        assert np.array_equal(ia, ib), np.argwhere(ia != ib)[:5]
        assert np.array_equal(qa, qb), ("f_qp_offset_aq", float(np.abs(qa - qb).max()), int((qa != qb).sum()))
        if mode:
            assert np.ptp(qa) > 1.0
    finally:
        r.xref_close(hnd)
