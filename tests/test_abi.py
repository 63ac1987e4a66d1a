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
