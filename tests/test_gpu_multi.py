"""Two GPUs of one box (skipped with fewer): ONE picture stream sharded over the GPUs from a host written in plain C --
examples/lookahead_host.c --ranks 2: one process per GPU (fork), searches and cost requests split by picture, the all-gathers done
by x264cu_exchange_nccl (ncclAllGather on the lookahead's exchange stream).  Every rank must print the decisions and MB-tree
offsets of the single-GPU run."""
import os
import subprocess

import numpy as np
import pytest
from _libs import synth_sequence

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_c_host_sharded_over_two_gpus_with_nccl(tmp_path):
    import x264_b200 as x
    exe = str(tmp_path / "lookahead_host")
    libdir = os.path.join(ROOT, "x264_b200", "csrc")
    x.lib()
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "lookahead_host.c"),
                           "-o", exe, "-L" + libdir, "-lx264_b200", "-Wl,-rpath," + libdir])
    w, h, n = 640, 368, 96
    frames = synth_sequence(w, h, n, seed=33, cut_at=57)
    rng = np.random.default_rng(4)
    raw = str(tmp_path / "pictures.i420")
    with open(raw, "wb") as f:
        for y in frames:
            f.write(y.tobytes())
            f.write(rng.integers(90, 170, (h // 2) * (w // 2), dtype=np.uint8).tobytes())
            f.write(rng.integers(90, 170, (h // 2) * (w // 2), dtype=np.uint8).tobytes())
    one = subprocess.run([exe, str(w), str(h), str(n), raw], check=True, capture_output=True, text=True, timeout=300)
    two = subprocess.run([exe, "--ranks", "2", str(w), str(h), str(n), raw], check=True, capture_output=True, text=True, timeout=300)
    frames_of = lambda text: [l for l in text.splitlines() if l.startswith("frame ")]       # NCCL prints its version on stdout
    assert len(frames_of(one.stdout)) == n
    assert frames_of(two.stdout) == frames_of(one.stdout)                      # rank 0
    rank1 = [l[len("rank 1 "):] for l in two.stderr.splitlines() if l.startswith("rank 1 frame")]
    assert rank1 == frames_of(one.stdout)                                       # rank 1 took the same decisions
    gathers = [l for l in two.stderr.splitlines() if "NCCL all-gathers" in l]
    assert len(gathers) == 2 and all(int(l.split(":")[1].split()[0]) >= 4 for l in gathers), gathers


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_config3_settings_sharded_over_two_gpus_equal_the_reference(tmp_path):
    """BASELINE configs[3]'s lookahead settings (rc-lookahead 250, bframes 16, b-adapt 2, B pyramid, MB-tree, AQ) on ONE stream sharded
    over two GPUs by the C host: every rank's frame types are those of the UNMODIFIED reference's own lookahead stage
    (oracle/_ref: x264_lookahead_put_frame / _get_frames -> x264_slicetype_decide) on the same pictures -- all of them, not a prefix."""
    import ctypes as C
    import _libs
    if not _libs.have_ref():
        pytest.skip("oracle/_ref did not travel")
    import x264_b200 as x
    exe = str(tmp_path / "lookahead_host")
    libdir = os.path.join(ROOT, "x264_b200", "csrc")
    x.lib()
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "lookahead_host.c"),
                           "-o", exe, "-L" + libdir, "-lx264_b200", "-Wl,-rpath," + libdir])
    w, h, n = 640, 368, 80
    frames = synth_sequence(w, h, n, seed=77, cut_at=49)
    raw = str(tmp_path / "pictures.i420")
    grey = np.full((h // 2) * (w // 2), 128, np.uint8).tobytes()
    with open(raw, "wb") as f:
        for y in frames:
            f.write(y.tobytes()); f.write(grey); f.write(grey)
    cfg = ["--bframes", "16", "--b-adapt", "2", "--rc-lookahead", "250"]
    two = subprocess.run([exe, "--ranks", "2"] + cfg + [str(w), str(h), str(n), raw], check=True, capture_output=True, text=True, timeout=600)
    got = [(int(l.split()[1]), l.split()[3]) for l in two.stdout.splitlines() if l.startswith("frame ")]
    rank1 = [(int(l.split()[3]), l.split()[5]) for l in two.stderr.splitlines() if l.startswith("rank 1 frame")]
    r = _libs.ref()
    r.xref_lookahead_types.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hnd = r.xref_open(w, h, b"medium", b"weightp=0:no-psy=1:aq-mode=1:bframes=16:b-adapt=2:rc-lookahead=250:mvrange=512", 0)
    assert hnd
    clip = np.ascontiguousarray(np.stack(frames))
    idx, ty = (C.c_int * n)(), (C.c_int * n)()
    k = r.xref_lookahead_types(hnd, clip.ctypes.data, n, idx, ty)
    r.xref_close(hnd)
    names = {1: "IDR", 2: "I", 3: "P", 4: "Bref", 5: "B"}
    want = [(int(idx[i]), names[int(ty[i])]) for i in range(k)]
    assert k == n and len(got) == n
    assert got == want
    assert rank1 == want
    assert sum(t == "B" for _, t in want) > n // 2 and any(t in ("I", "IDR") for f, t in want if f == 49)     # long mini-GOPs, and the cut
