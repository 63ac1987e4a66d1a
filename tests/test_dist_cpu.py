"""world_size-2 gloo tests of the N>1 plumbing (x264_b200/dist.py): (1) every rank contributes its stream's decision records,
one all-gather, every rank ends up with all streams; (2) ONE stream sharded over two ranks: the host logic in sharded mode
(job partition by picture, one exchange per prefetch group, one group late) takes the decisions of the unsharded run."""
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


def _run_shard(tmp_path, world, tag, *extra):
    port = str(30500 + (os.getpid() + world + 7 * len(extra)) % 1000)
    outs = [str(tmp_path / ("%s%d.json" % (tag, r))) for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_shard_worker.py"), str(r), str(world), port, outs[r]] + list(extra))
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    return [json.load(open(o)) for o in outs]


def test_sharded_stream_matches_single_rank(tmp_path):
    single = _run_shard(tmp_path, 1, "s")[0]
    both = _run_shard(tmp_path, 2, "d")
    assert single["exchanges"] == 0
    for r in both:
        assert r["types"] == single["types"]
        assert r["exchanges"] >= 3 and r["exchanges"] == both[0]["exchanges"] and r["bytes"] == both[0]["bytes"]


def test_sharded_trellis_stream_matches_single_rank(tmp_path):
    """the same with BASELINE configs[3]'s kind of window (b-adapt 2 over a B pyramid, bframes 8): more searches per picture, the
    pruned set of speculated triples, long mini-GOPs"""
    single = _run_shard(tmp_path, 1, "ts", "trellis")[0]
    both = _run_shard(tmp_path, 2, "td", "trellis")
    assert len(single["types"]) == 70 and sum(t in (4, 5) for _, t in single["types"]) > 35
    for r in both:
        assert r["types"] == single["types"]
        assert r["exchanges"] >= 3 and r["exchanges"] == both[0]["exchanges"] and r["bytes"] == both[0]["bytes"]
