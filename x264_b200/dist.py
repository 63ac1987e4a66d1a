"""Multi-GPU plumbing (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

Two ways of using N GPUs:
  * one independent picture stream per GPU (no data-path collective; bench.py's weak-scaling line), followed by ONE all-gather
    of the fixed-size per-picture decision records so that rank 0 holds every stream's slice-type decisions;
  * ONE stream sharded over the GPUs: every rank is fed the same pictures and takes the same decisions, the lowres
    searches of each prefetch group are split by picture and their results (8 bytes per macroblock and search) exchanged with one
    all-gather per group -- ShardExchange below is the collective x264cu_slicetype_set_shard calls back into."""
import ctypes as C

import numpy as np

EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p)


class ShardExchange:
    """the all-gather of x264cu_slicetype_set_shard over torch.distributed: NCCL on the lookahead's exchange stream (device
    buffers), or gloo on host buffers in the CPU harness, where nothing real travels"""

    def __init__(self, dist, device=None):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.world = dist.get_world_size()
        self.buf = {0: [None, None], 1: [None, None]}      # channel -> [send, recv]: searches (phases 0/1), cost requests (phases 2/3)
        self.calls = 0
        self.bytes = 0
        self.cb = EXCHANGE_FN(self._call)          # keep the trampoline alive as long as this object

    def _call(self, user, phase, per_rank, d_send, d_recv, stream):
        try:
            torch = self.torch
            ch, phase = phase >> 1, phase & 1
            n = max(int(per_rank), 16)
            send, recv = self.buf[ch]
            if phase == 0:
                if send is None or send.numel() < n:
                    dev = self.device if self.device is not None else "cpu"
                    if send is not None and self.device is not None:
                        # copies out of the old receive buffer may still be in flight on the exchange stream
                        torch.cuda.ExternalStream(int(stream), device=self.device).synchronize()
                    grow = n + n // 4
                    send = torch.zeros(grow, dtype=torch.uint8, device=dev)
                    recv = torch.zeros(grow * self.world, dtype=torch.uint8, device=dev)
                    self.buf[ch] = [send, recv]
                # rank r's block must sit at r * per_rank: gather into a view of exactly world * per_rank bytes
                d_send[0] = send.data_ptr()
                d_recv[0] = recv.data_ptr()
                return 0
            send, recv = send[:int(per_rank)] if per_rank else send[:0], recv[:int(per_rank) * self.world]
            if per_rank:
                if self.device is not None:
                    with torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=self.device)):
                        self.dist.all_gather_into_tensor(recv, send)
                else:
                    self.dist.all_gather_into_tensor(recv, send)
            self.calls += 1
            self.bytes += int(per_rank) * self.world
            return 0
        except Exception as e:                      # never let an exception cross the C boundary
            import sys
            sys.stderr.write("ShardExchange: %r\n" % (e,))
            return -1

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
