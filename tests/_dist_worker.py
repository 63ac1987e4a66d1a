"""worker for tests/test_dist_cpu.py: python _dist_worker.py RANK WORLD PORT OUTFILE"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from x264_b200 import dist as xd

rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
os.environ["MASTER_ADDR"] = "127.0.0.1"
os.environ["MASTER_PORT"] = port
dist.init_process_group("gloo", rank=rank, world_size=world)
decisions = [(i, 1 + (i + rank) % 5) for i in range(10 + rank)]
got = xd.unpack_records(xd.all_gather_records(dist, xd.pack_records(rank, decisions, 16)))
json.dump({str(k): v for k, v in got.items()}, open(out, "w"))
dist.destroy_process_group()
