"""The drop-in boundary proven inside the reference: oracle/_ref/libx264ref_b200.so is the UNMODIFIED reference
compiled with HAVE_OPENCL=1 and integration/x264_b200_hooks.c in the place of common/opencl.c + encoder/slicetype-cl.c
(oracle/Makefile.ref, target b200).  With `opencl=1` the reference ENCODER runs its lookahead on the B200 through
x264_opencl_lowres_init / _motionsearch / _finalize_cost / _flush / _slicetype_prep; everything else -- slice-type decision,
MB-tree, rate control, analysis, entropy coding -- is the reference's own code.  The coded frame types and the BITSTREAM must be
identical to the run with the reference's CPU lookahead."""
import ctypes as C
import os
import numpy as np
import pytest
import _libs
from _libs import have_ref, ROOT

HOOKED = os.path.join(ROOT, "oracle", "_ref", "libx264ref_b200.so")
pytestmark = pytest.mark.skipif(not (have_ref() and os.path.exists(HOOKED)), reason="compiled reference (hooked variant) not present")


def _bind(lib):
    vp, ci = C.c_void_p, C.c_int
    lib.xref_open.restype = vp
    lib.xref_open.argtypes = [ci, ci, C.c_char_p, C.c_char_p, ci]
    lib.xref_close.argtypes = [vp]
    lib.xref_encode_i420_hash.argtypes = [vp, vp, ci, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), vp, vp]
    lib.xref_offload_active.argtypes = [vp]
    return lib


def encode(lib, w, h, preset, opts, yuv, n):
    hnd = lib.xref_open(w, h, preset, opts, 0)
    assert hnd
    try:
        hv, nb = C.c_uint64(), C.c_int64()
        idx, ty = (C.c_int * (n + 8))(), (C.c_int * (n + 8))()
        k = lib.xref_encode_i420_hash(hnd, yuv.ctypes.data, n, C.byref(hv), C.byref(nb), idx, ty)
        active = lib.xref_offload_active(hnd)
    finally:
        lib.xref_close(hnd)
    assert k == n, k
    return [(idx[i], ty[i]) for i in range(k)], hv.value, nb.value, active


def clip(w, h, n, seed):
    import _me_trace as T
    return T.synth_i420(w, h, n, seed)


def test_hooked_reference_exports_the_seam_and_falls_back_without_a_device():
    """every symbol the reference calls at the seam is provided by the hooks object; without a CUDA device
    x264_opencl_lookahead_init fails and the reference itself switches the offload off (encoder.c:1798-1799)"""
    import subprocess
    syms = subprocess.check_output(["nm", "-D", "--defined-only", HOOKED], text=True)
    for s in ("x264_8_opencl_load_library", "x264_8_opencl_close_library", "x264_8_opencl_lookahead_init", "x264_8_opencl_lookahead_delete",
              "x264_8_opencl_frame_delete", "x264_8_opencl_lowres_init", "x264_8_opencl_motionsearch", "x264_8_opencl_finalize_cost",
              "x264_8_opencl_flush", "x264_8_opencl_slicetype_prep", "x264_8_opencl_slicetype_end", "x264_8_opencl_precalculate_frame_cost"):
        assert (" T " + s) in syms, s
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the gpu tests")
    lib = _bind(C.CDLL(HOOKED))
    w, h, n = 64, 48, 6
    types, hv, nb, active = encode(lib, w, h, b"medium", b"opencl=1:rc-lookahead=4:bframes=1", clip(w, h, n, 3), n)
    assert active == 0 and nb > 0


CASES = [
    # BASELINE configs[0]: 1280x720 --preset ultrafast
    ("720p ultrafast", 1280, 720, b"ultrafast", b"threads=1", 24),
    # BASELINE configs[1]: 1920x1080 --preset medium --rc-lookahead 40 (weightp 2, aq 1, mb-tree, b-adapt 1)
    ("1080p medium", 1920, 1080, b"medium", b"rc-lookahead=40:threads=1", 56),
    # trellis B decision, VBV lookahead, pyramid
    ("360p slow trellis vbv", 640, 360, b"slow", b"b-adapt=2:bframes=5:rc-lookahead=30:vbv-bufsize=3000:vbv-maxrate=3000:threads=1", 48),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_encoder_on_the_b200_lookahead_is_bit_identical(case):
    name, w, h, preset, opts, n = case
    yuv = clip(w, h, n, 720 + w)
    cpu = _bind(_libs.ref())
    hooked = _bind(C.CDLL(HOOKED))
    hooked.x264_b200_hooks_calls.restype = C.c_long
    hooked.x264_b200_hooks_calls.argtypes = [C.c_int]
    want = encode(cpu, w, h, preset, opts, yuv, n)
    before = [hooked.x264_b200_hooks_calls(i) for i in range(4)]
    got = encode(hooked, w, h, preset, opts + b":opencl=1", yuv, n)
    calls = [hooked.x264_b200_hooks_calls(i) - before[i] for i in range(4)]
    assert got[3] == 1, "the offload was switched off during the run"
    assert calls[0] == n and calls[1] > 0, calls            # every picture uploaded once, cost requests answered on the device
    assert got[0] == want[0], "frame types differ"
    assert (got[1], got[2]) == (want[1], want[2]), "bitstream differs: %d vs %d bytes" % (got[2], want[2])
