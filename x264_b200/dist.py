"""Multi-GPU plumbing (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

Round 1 sharding: one independent picture stream per GPU (no data-path collective, SURVEY 8e "replicas" row for the
searches), followed by ONE all-gather of the fixed-size per-picture decision records so that rank 0 holds every stream's
slice-type decisions -- the exchange step the north-star describes for the lookahead results."""
import numpy as np

RECORD_INTS = 4          # (stream, display index, slice type, reserved)


def pack_records(stream, decisions, capacity):
    """decisions: [(display_index, type)] -> int32[capacity, RECORD_INTS], unused rows = -1"""
    out = np.full((capacity, RECORD_INTS), -1, np.int32)
    n = min(len(decisions), capacity)
    for i in range(n):
        out[i] = (stream, decisions[i][0], decisions[i][1], 0)
    return out


def all_gather_records(dist, records, device=None):
    """records: int32[capacity, RECORD_INTS] on every rank -> int32[world, capacity, RECORD_INTS] on every rank"""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(records))
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)          # concatenation along dim 0 (the layout both NCCL and gloo accept)
    return out.cpu().numpy().reshape((world,) + tuple(t.shape))


def unpack_records(gathered):
    """-> {stream: [(display_index, type), ...]} in coded order"""
    res = {}
    for r in gathered.reshape(-1, RECORD_INTS):
        if r[0] >= 0:
            res.setdefault(int(r[0]), []).append((int(r[1]), int(r[2])))
    return res
