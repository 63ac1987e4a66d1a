"""bench.py's second CPU row: the reference's C path auto-vectorised for AVX2 (oracle/Makefile.ref, target vec).  It is only a
timing arm, but a timing arm that computed something else would be worthless: its metrics must equal the stock build's."""
import numpy as np
import pytest
import _libs


@pytest.mark.skipif(not _libs.have_ref(), reason="oracle/_ref not built")
def test_autovectorized_reference_build_computes_the_same_metrics():
    v = _libs.ref_vec()
    if v is None:
        pytest.skip("libx264ref_vec.so not built, or this host has no AVX2")
    r = _libs.ref()
    rng = np.random.default_rng(7)
    w = h = st = 192
    f = rng.integers(0, 256, h * st, dtype=np.uint8)
    g = rng.integers(0, 256, h * st, dtype=np.uint8)
    n = 3000
    cand = np.zeros(n, _libs.cand_dtype)
    cand["fenc_off"] = rng.integers(0, h - 16, n) * st + rng.integers(0, w - 16, n)
    cand["ref_off"] = rng.integers(0, h - 16, n) * st + rng.integers(0, w - 16, n)
    for metric in range(3):
        for i_pixel in range(7):
            a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
            r.xref_pixel_cmp_batch(metric, i_pixel, f, st, g, st, cand, n, a)
            v.xref_pixel_cmp_batch(metric, i_pixel, f, st, g, st, cand, n, b)
            assert np.array_equal(a, b), (metric, i_pixel)
