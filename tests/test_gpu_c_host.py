"""The boundary from the reference's side: a host written in plain C (examples/lookahead_host.c: only include/x264_b200.h and
libx264_b200.so, built with gcc) reads raw I420 and drives the lookahead like x264_encoder_encode would (adaptive quantisation inside).  Its frame types and MB-tree offsets must
be those the Python binding gets for the same pictures (which the other GPU tests pin to the reference)."""
import os
import subprocess

import numpy as np
import pytest
import x264_b200 as x
from _libs import synth_sequence

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {1: "IDR", 2: "I", 3: "P", 4: "Bref", 5: "B"}


def test_plain_c_host_gets_the_same_decisions(tmp_path):
    exe = str(tmp_path / "lookahead_host")
    libdir = os.path.join(ROOT, "x264_b200", "csrc")
    x.lib()                                                        # builds the library if it is missing
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "lookahead_host.c"),
                           "-o", exe, "-L" + libdir, "-lx264_b200", "-Wl,-rpath," + libdir])
    w, h, n = 320, 192, 60
    frames = synth_sequence(w, h, n, seed=21, cut_at=37)
    rng = np.random.default_rng(2)
    chroma = [(rng.integers(90, 170, (h // 2, w // 2), dtype=np.uint8), rng.integers(90, 170, (h // 2, w // 2), dtype=np.uint8)) for _ in range(n)]
    raw = str(tmp_path / "pictures.i420")
    with open(raw, "wb") as f:
        for y, (cb, cr) in zip(frames, chroma):
            f.write(y.tobytes()); f.write(cb.tobytes()); f.write(cr.tobytes())
    out = subprocess.run([exe, str(w), str(h), str(n), raw], check=True, capture_output=True, text=True, timeout=300).stdout
    got = [(int(l.split()[1]), l.split()[3], float(l.split()[5])) for l in out.strip().splitlines()]
    ctx = x.Context(0)
    try:
        st = x.Slicetype(ctx, w, h, rc_lookahead=20, psy=0, aq_mode=1, aq_strength=1.0, mb_tree=1, bframes=3)
        qp = {}
        want = st.decide(frames, qp, chroma=chroma)
        st.close()
    finally:
        ctx.close()
    assert [(f, t) for f, t, _ in got] == [(f, NAMES[t]) for f, t in want]
    assert any(t in ("IDR", "I") for f, t, _ in got if f == 37)                       # the cut
    for f, t, mean in got:
        if t not in ("B", "Bref"):
            assert abs(mean - float(qp[f].astype(np.float64).mean())) < 1e-3, (f, mean)
    assert any(abs(m) > 0.1 for _, _, m in got)
