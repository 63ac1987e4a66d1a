"""world_size-2 gloo test of the N>1 plumbing (x264_b200/dist.py): every rank contributes its stream's decision records,
one all-gather, every rank ends up with all streams."""
import json
import os
import subprocess
import sys

from x264_b200 import dist as xd

HERE = os.path.dirname(os.path.abspath(__file__))


def test_all_gather_of_decision_records(tmp_path):
    world = 2
    port = str(29500 + os.getpid() % 1000)
    outs = [str(tmp_path / ("r%d.json" % r)) for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(r), str(world), port, outs[r]])
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    want = {str(r): [[i, 1 + (i + r) % 5] for i in range(10 + r)] for r in range(world)}
    for o in outs:
        assert json.load(open(o)) == want


def test_pack_unpack_roundtrip():
    d = [(3, 1), (1, 5), (2, 5), (0, 3)]
    rec = xd.pack_records(7, d, 8)
    assert rec.shape == (8, xd.RECORD_INTS) and (rec[4:] == -1).all()
    assert xd.unpack_records(rec[None]) == {7: d}
