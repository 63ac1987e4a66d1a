"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/x264_b200.h
declares, and refuses loudly to work without a device (no CPU fallback)."""
import pytest
import x264_b200 as x


def test_library_exports_every_declared_symbol():
    x.lib()
    declared = set(x.header_symbols())
    exported = set(x.exported_symbols())
    assert declared, "header declares nothing?"
    assert not (declared - exported), "declared but not exported: %s" % sorted(declared - exported)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(x.X264CUError) as e:
        x.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """the shipped package must never import / link the oracle"""
    import os
    root = os.path.dirname(os.path.abspath(x.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "libx264ref" not in txt and "oracle/" not in txt, f


def test_open_fails_loudly_without_a_device():
    """there is no CPU fallback: on a machine without an sm_100 GPU x264cu_open returns -1 and says why"""
    import ctypes as C
    import shutil
    import subprocess
    import x264_b200 as x
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        import pytest
        pytest.skip("a GPU is present")
    L = x.lib()
    h = C.c_void_p()
    assert L.x264cu_open(C.byref(h), 0) == -1 and not h.value
    msg = L.x264cu_strerror(None).decode()
    assert "no CPU fallback" in msg or "CUDA" in msg, msg


def test_entry_points_reject_null_handles_without_a_device():
    """every batched / stateful entry point returns -1 for a NULL context or object before touching CUDA (the reference's hooks
    return -1 and leave the fallback to the caller); nothing here needs a GPU"""
    import ctypes as C
    import x264_b200 as x
    L = x.lib()
    n = None
    assert L.x264cu_me_search_batch(n, n, n, 0, n, n, 0, n, 1, n) == -1
    assert L.x264cu_me_refine_qpel_batch(n, n, 0, n, 0, n, 0, n, 1, n) == -1
    assert L.x264cu_me_refine_bidir_batch(n, n, n, 0, n, n, 0, n, 1, n) == -1
    L.x264cu_lookahead_frame_cost_recalculate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    assert L.x264cu_lookahead_frame_cost_recalculate(n, 0, 0, 0, 0, n, n) == -1
    L.x264cu_slicetype_rc_analyse_slice.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.x264cu_slicetype_rc_analyse_slice(n, 0, n, n, n) == -1
    L.x264cu_slicetype_get_planned.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    assert L.x264cu_slicetype_get_planned(n, 0, n, n, 0) == -1
    L.x264cu_slicetype_get_qp_offset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    assert L.x264cu_slicetype_get_qp_offset(n, 0, n) == -1
