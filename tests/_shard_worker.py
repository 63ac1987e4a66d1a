"""worker for tests/test_dist_cpu.py::test_sharded_stream_*: python _shard_worker.py RANK WORLD PORT OUTFILE [trellis]
One picture stream sharded over WORLD ranks on CPU: the product's host logic (x264_b200/csrc/slicetype.c, sharded mode) over the
oracle glue, the exchange callback over gloo.  Writes the decisions and the exchange statistics."""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import torch.distributed as dist
from x264_b200 import dist as xd
from x264_b200.binding_ext import SlicetypeParams, LookaheadParams
from _libs import slicetype_oracle_lib, synth_sequence

rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = port
dist.init_process_group("gloo", rank=rank, world_size=world)
w, h, n = 96, 64, 70
frames = synth_sequence(w, h, n, seed=5, cut_at=33)
trellis = len(sys.argv) > 5 and sys.argv[5] == "trellis"       # BASELINE configs[3]'s kind of window: b-adapt 2 over a B pyramid, long mini-GOPs
la = LookaheadParams(w, h, 7, 1, 16, 512, 8 if trellis else 3, 0, 1, 0, 1, 0, 0, 0)
p = SlicetypeParams(la, 250, 25, 40, 2 if trellis else 1, 2, 40 if trellis else 20, 0, 3, 0)
lib = slicetype_oracle_lib()
lib.x264cu_slicetype_open.argtypes = [C.c_void_p, C.POINTER(SlicetypeParams), C.POINTER(C.c_void_p)]
lib.x264cu_slicetype_step.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.x264cu_slicetype_close.argtypes = [C.c_void_p]
lib.x264cu_slicetype_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
st = C.c_void_p()
assert lib.x264cu_slicetype_open(C.c_void_p(1), C.byref(p), C.byref(st)) == 0
ex = xd.ShardExchange(dist, device=None)
if world > 1:
    assert lib.x264cu_slicetype_set_shard(st, rank, world, C.cast(ex.cb, C.c_void_p), None) == 0
types = []
fr, ty = C.c_int(), C.c_int()
for f in frames:
    assert lib.x264cu_slicetype_step(st, f.ctypes.data, f.shape[1], None, C.byref(fr), C.byref(ty)) == 0
    if fr.value >= 0:
        types.append((fr.value, ty.value))
while True:
    assert lib.x264cu_slicetype_step(st, None, 0, None, C.byref(fr), C.byref(ty)) == 0
    if fr.value < 0:
        break
    types.append((fr.value, ty.value))
lib.x264cu_slicetype_close(st)
json.dump({"types": types, "exchanges": ex.calls, "bytes": ex.bytes}, open(out, "w"))
dist.barrier()
dist.destroy_process_group()
