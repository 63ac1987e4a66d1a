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
