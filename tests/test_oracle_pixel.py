"""Pins oracle/oracle_pixel.c against the compiled reference (common/pixel.c C entries) on checkasm-style
inputs (tools/checkasm.c:381-435): random buffers plus worst-case overflow patterns, all 8 block sizes."""
import numpy as np
import pytest
from _libs import oracle, ref, have_ref, cand_dtype, PIXEL_W, PIXEL_H, SAD, SSD, SATD, SA8D, worst_case_pair

pytestmark = pytest.mark.skipif(not have_ref(), reason="compiled reference not present")

METRICS = [(SAD, range(8)), (SSD, range(8)), (SATD, range(8)), (SA8D, [0, 3])]


@pytest.mark.parametrize("pattern", ["random", "worst", "lowvar"])
def test_pixel_metrics_match_reference(pattern):
    rng = np.random.default_rng(1234)
    stride = 64
    n = stride * 48
    for rep in range(6):
        if pattern == "random":
            a = rng.integers(0, 256, n, dtype=np.uint8)
            b = rng.integers(0, 256, n, dtype=np.uint8)
        elif pattern == "worst":
            a, b = worst_case_pair(n, rng)
        else:
            a = rng.integers(100, 110, n, dtype=np.uint8)
            b = rng.integers(100, 110, n, dtype=np.uint8)
        ncand = 64
        cand = np.zeros(ncand, cand_dtype)
        cand["fenc_off"] = (rng.integers(0, 16, ncand) * stride + rng.integers(0, 3, ncand) * 16).astype(np.uint32)
        cand["ref_off"] = (rng.integers(0, 16, ncand) * stride + rng.integers(0, 40, ncand)).astype(np.uint32)
        for metric, sizes in METRICS:
            for ip in sizes:
                o = np.zeros(ncand, np.int32)
                r = np.zeros(ncand, np.int32)
                oracle().orc_pixel_cmp_batch(metric, ip, a, stride, b, stride, cand, ncand, o)
                ref().xref_pixel_cmp_batch(metric, ip, a, stride, b, stride, cand, ncand, r)
                assert np.array_equal(o, r), (pattern, metric, ip)


def test_known_answers():
    """hand-checkable values: constant difference d over WxH -> SAD = d*W*H, SSD = d^2*W*H,
    SATD = DC only = d*W*H/2 ... per 4x4: |16d|/2 = 8d"""
    a = np.full(64 * 32, 10, np.uint8)
    b = np.full(64 * 32, 13, np.uint8)
    cand = np.zeros(1, cand_dtype)
    for ip in range(8):
        w, h = PIXEL_W[ip], PIXEL_H[ip]
        for metric, want in ((SAD, 3 * w * h), (SSD, 9 * w * h), (SATD, 8 * 3 * (w // 4) * (h // 4))):
            o = np.zeros(1, np.int32)
            oracle().orc_pixel_cmp_batch(metric, ip, a, 64, b, 64, cand, 1, o)
            assert o[0] == want
