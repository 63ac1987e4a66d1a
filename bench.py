#!/usr/bin/env python
"""bench.py -- headline benchmark of the x264 ME / lookahead cost path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload satd|lookahead|me|sweep]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:
  workload satd      : 16x16 SATD of every macroblock of 32 synthetic 4K frame pairs (1 036 800 candidates, each
                       against a reference block displaced by a random full-pel vector within +-16) -- the
                       "16x16 SATD macroblocks/s vs HBM roofline" half of BASELINE.json's metric.
  workload lookahead : lowres lookahead frames/s at 4K (default; the 16x16 SATD line rides on it).
  workload me        : BASELINE configs[2]: the x264_me_t stream of a 4K --preset slower --me umh --merange 64 encode, recorded from
                       the reference encoder on this box and replayed through x264cu_me_search_frame (100 % parity checked).
  workload sweep     : BASELINE configs[4]: SAD / SATD / SSD x 4x4..16x16, HBM-roofline fraction per cell.
`value` is measured with the inputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry point with
pinned host memory, H2D/D2H copies inside the timed region.  Under torchrun (N>1) every rank processes its own
batch (weak scaling, no data-path collective); timing is the max over ranks of device-event time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W4K, H4K = 3840, 2160
PAD = 32
N_PAIRS = 32                      # 32 x 32400 = 1 036 800 candidates ~ "1M candidates" of BASELINE config 5
ALG_BYTES_16x16 = 2 * 16 * 16 + 4  # unique fenc block + unique ref block + one int32 result


def x264_stride(width):
    s = (width + 96 + 63) // 64 * 64          # align_stride(width + PADH2, 64, 1024), common/frame.c:30-36,87
    if s % 1024 == 0:
        s += 64
    return s


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel, units_key, units_now):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01b_kernel_traffic.json), scaled to the
    units this launch processes; None if the capture is missing"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01b_kernel_traffic.json")))[kernel]
        per_unit = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t[units_key]
        return per_unit * units_now, t["source"]
    except Exception:
        return None, None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# synthetic input
# ------------------------------------------------------------------------------------------------------
def make_satd_inputs(seed, n_pairs, alloc):
    """n_pairs padded 4K luma planes for fenc and ref (uniform random u8) + mv field"""
    stride = x264_stride(W4K)
    pitch = stride * (H4K + 2 * PAD)
    rng = np.random.default_rng(seed)
    fenc = alloc(pitch * n_pairs)
    ref = alloc(pitch * n_pairs)
    for buf in (fenc, ref):
        for f in range(n_pairs):        # whole padded plane random: the border is as good as replicated data here
            buf[f * pitch:(f + 1) * pitch] = rng.integers(0, 256, pitch, dtype=np.uint8)
    mv = rng.integers(-16, 17, (n_pairs, H4K // 16, W4K // 16, 2)).astype(np.int16)
    return fenc, ref, mv, stride, pitch


def cand_list_for_frame(mv_f, stride):
    by, bx = mv_f.shape[:2]
    yy, xx = np.meshgrid(np.arange(by), np.arange(bx), indexing="ij")
    org = PAD * stride + PAD
    fo = org + yy * 16 * stride + xx * 16
    ro = fo + mv_f[..., 1].astype(np.int64) * stride + mv_f[..., 0]
    c = np.zeros(by * bx, dtype=[("fenc_off", np.uint32), ("ref_off", np.uint32)])
    c["fenc_off"] = fo.reshape(-1)
    c["ref_off"] = ro.reshape(-1)
    return c


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own pixel table (oracle/_ref) or, if it did not travel, the oracle port
# ------------------------------------------------------------------------------------------------------
def cpu_satd_rate(fenc, ref, mv, stride, pitch, budget_s, n_frames_sample, lib=None):
    """lib: another build of the reference (the auto-vectorised one) instead of the stock C path"""
    import _libs
    if lib is not None:
        L, fn, kind = lib, "xref_pixel_cmp_batch", "reference (auto-vectorised C)"
    elif _libs.have_ref():
        L, fn, kind = _libs.ref(), "xref_pixel_cmp_batch", "reference"
    else:
        L, fn, kind = _libs.oracle(), "orc_pixel_cmp_batch", "port"
    f = getattr(L, fn)
    cores = os.cpu_count() or 1
    frames = list(range(min(n_frames_sample, mv.shape[0])))
    cands = [cand_list_for_frame(mv[i], stride) for i in frames]
    outs = [np.zeros(len(c), np.int32) for c in cands]

    def work(tid, reps):
        for _ in range(reps):
            for i in frames:
                n = len(cands[i])
                lo, hi = n * tid // cores, n * (tid + 1) // cores
                if hi > lo:
                    fv = fenc[i * pitch:(i + 1) * pitch]
                    rv = ref[i * pitch:(i + 1) * pitch]
                    f(2, 0, fv, stride, rv, stride, cands[i][lo:hi], hi - lo, outs[i][lo:hi])

    def run(reps):
        th = [threading.Thread(target=work, args=(t, reps)) for t in range(cores)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return time.perf_counter() - t0

    t1 = run(1)
    reps = max(1, int(budget_s / max(t1, 1e-3)))
    t = run(reps)
    n = reps * sum(len(c) for c in cands)
    return n / t, kind, cores, "%d x %d 4K frames (%d candidates) on %d threads, %.1f s" % (reps, len(frames), n, cores, t)


# ------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    return rank, world, local, dist


def max_over_ranks(dist, v, local):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def run_satd_b200(args, rank, world, local, dist):
    import x264_b200 as x
    ctx = x.Context(local)
    info = ctx.device_info()
    pinned = lambda n: ctx.malloc_host(n)
    fenc, ref, mv, stride, pitch = make_satd_inputs(1 + rank, N_PAIRS, pinned)
    n_cand = mv.shape[0] * mv.shape[1] * mv.shape[2]
    d_fenc, d_ref = ctx.malloc(fenc.nbytes + 256), ctx.malloc(ref.nbytes + 256)
    ctx.h2d(d_fenc, fenc)
    ctx.h2d(d_ref, ref)
    d_mv = ctx.upload(mv)
    d_out = ctx.malloc(n_cand * 4)
    org = PAD * stride + PAD
    pf = (d_fenc + org, stride, pitch, W4K, H4K, N_PAIRS)
    pr = (d_ref + org, stride, pitch, W4K, H4K, N_PAIRS)

    def step():
        ctx.pixel_cmp_mvfield(x.SATD, 0, pf, pr, 1, d_mv, d_out)

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()
    sampler = ClockSampler(local)
    barrier(dist, local)
    sampler.start()
    l0 = ctx.launches
    t_wall0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    ctx.sync()
    barrier(dist, local)
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launches - l0
    ms = max_over_ranks(dist, ms, local)
    clocks = sampler.stop()

    # correctness spot check of the timed output against the checker (not part of the timed region)
    import _libs
    got = ctx.download(d_out, (n_cand,), np.int32)[:32400]
    want = np.zeros(32400, np.int32)
    c0 = cand_list_for_frame(mv[0], stride)
    (_libs.ref().xref_pixel_cmp_batch if _libs.have_ref() else _libs.oracle().orc_pixel_cmp_batch)(
        2, 0, fenc[:pitch], stride, ref[:pitch], stride, c0, len(c0), want)
    parity_ok = bool(np.array_equal(got, want))

    # e2e: host planes -> C-ABI host entry point -> host costs, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5)) if not args.quick else 1
    ctx.pixel_cmp_mvfield_host(x.SATD, 0, fenc, ref, stride, pitch, W4K, H4K, N_PAIRS, 1, mv)   # warm (allocs scratch)
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_h = ctx.pixel_cmp_mvfield_host(x.SATD, 0, fenc, ref, stride, pitch, W4K, H4K, N_PAIRS, 1, mv)
    barrier(dist, local)
    e2e_s = max_over_ranks(dist, time.perf_counter() - t0, local)
    parity_ok &= bool(np.array_equal(out_h[:32400], want))

    kernel_ms = ms / args.steps
    peaks, peak_src = measured_peaks()
    achieved = ALG_BYTES_16x16 * n_cand / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic("mvfield_kernel<SATD,16,16>", "candidates_per_launch", n_cand)
    res = {
        "metric": "satd_16x16_macroblocks_per_sec_4k", "value": n_cand * world / (kernel_ms * 1e-3), "unit": "macroblocks/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] at 4K: 16x16 SATD, %d candidates per GPU = every macroblock of %d "
                               "random 3840x2160 luma frame pairs, ref displaced by a random full-pel mv in +-16" % (n_cand, N_PAIRS),
                   "l2": "inputs %.0f MB per step > 126 MB L2, no flush needed" % ((fenc.nbytes + ref.nbytes) / 1e6),
                   "kernel": "mvfield_kernel<SATD,16,16> TMA tile 128x64 +-16 halo, 2-stage", "parity_spot_check": parity_ok},
        "clocks": clocks,
        "e2e": {"value": n_cand * world * e2e_steps / e2e_s, "unit": "macroblocks/s",
                "h2d_bytes_per_step": int(fenc.nbytes + ref.nbytes + mv.nbytes), "d2h_bytes_per_step": int(n_cand * 4),
                "api": "x264cu_pixel_cmp_mvfield_host (pinned host planes)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_16x16 * n_cand, "kernel": "mvfield_kernel<SATD,16,16>"},
        "wall_s": t_wall, "sm_count": info["sm_count"],
    }
    if rank == 0 and world == 1 and not args.quick:
        rate, kind, cores, sample = cpu_satd_rate(fenc, ref, mv, stride, pitch, args.cpu_budget, 8)
        res["cpu_baseline"] = {"value": rate, "unit": "macroblocks/s", "cores": cores, "kind": kind, "sample": sample}
        import _libs
        if _libs.ref_vec() is not None:
            # a second, labelled row: the same C built -O3 -ftree-vectorize -march=x86-64-v3 (the reference's configure builds its C
            # with -fno-tree-vectorize; its x86 asm cannot be assembled here: no nasm / yasm in the image, none in the wheelhouse)
            rate_v, _, _, sample_v = cpu_satd_rate(fenc, ref, mv, stride, pitch, args.cpu_budget / 2, 8, lib=_libs.ref_vec())
            res["cpu_baseline"]["autovectorized_c"] = {"value": rate_v, "unit": "macroblocks/s", "cores": cores, "sample": sample_v,
                                                       "build": "oracle/Makefile.ref vec: -O3 -ffast-math -ftree-vectorize -march=x86-64-v3"}
    ctx.close()
    return res


def run_satd_reference(args, rank, world):
    alloc = lambda n: np.empty(n, np.uint8)
    fenc, ref, mv, stride, pitch = make_satd_inputs(1, 8, alloc)
    rates = []
    kind = cores = sample = None
    for i in range(args.warmup + args.steps):
        r, kind, cores, sample = cpu_satd_rate(fenc, ref, mv, stride, pitch, args.cpu_budget / max(1, args.steps), 8)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    n_step = 8 * 32400
    return {
        "impl": "reference", "metric": "satd_16x16_macroblocks_per_sec_4k", "value": v, "unit": "macroblocks/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_step / v * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] at 4K: 16x16 SATD, bounded sample of 8 of the 32 frame pairs per step"},
        "cpu_baseline": {"value": v, "unit": "macroblocks/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "macroblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference C path (pixf.satd[PIXEL_16x16], common/pixel.c:312); x86 asm unavailable: no nasm in the image",
    }



# ------------------------------------------------------------------------------------------------------
# workload "lookahead": lowres lookahead frames/s at 4K (BASELINE.json metric, first half)
# ------------------------------------------------------------------------------------------------------
LA_W, LA_H = 3840, 2160
LA_FRAMES = 32                    # pictures per step
LA_OPTS = dict(subpel_refine=7, me_method=1, me_range=16, mv_range=512, bframes=3, bframe_bias=0, weighted_bipred=1,
               aq_mode=0, mb_tree=1, vbv=0)
LA_ST = dict(keyint_max=250, keyint_min=25, scenecut_threshold=40, b_adapt=1, b_pyramid=2, rc_lookahead=40, psy=0,
             frame_reference=3, rc_cqp=0)
LA_REF_OPTS = b"weightp=0:no-psy=1:aq-mode=0:bframes=3:rc-lookahead=40"          # the same configuration, reference spelling
LA_REF_OPTS_W = b"weightp=2:aq-mode=0:bframes=3:rc-lookahead=40"                  # --weightp 1: preset medium's own weightp / psy
# algorithmic bytes of one lowres motion search at WxH: fenc lowres + 4 reference lowres planes + 8 B/MB out
LA_WARP_INSTR_PER_MB_SEARCH = 1334          # ncu --set full of search_kernel<8> (84-search launch): 3 631 M warp instructions / 2.72 M macroblock searches
LA_SEARCH_BYTES = LA_W * LA_H // 4 + LA_W * LA_H + 8 * ((LA_W + 15) // 16) * ((LA_H + 15) // 16)


def make_la_frames(seed, n, alloc, LA_W=LA_W, LA_H=LA_H):
    """synthetic sequence (4K unless told otherwise): low-pass texture translated by a per-frame global motion + noise, one hard cut"""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (LA_H // 8 + 64, LA_W // 8 + 64)).astype(np.float32)
    k = np.ones(5, np.float32) / 5
    small = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 0, small)
    small = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 1, small)
    master = np.kron(small, np.ones((8, 8), np.float32))
    master2 = master[::-1, ::-1].copy()
    out = alloc(n * LA_W * LA_H).reshape(n, LA_H, LA_W)
    x = y = 128
    for i in range(n):
        m = master2 if i >= (2 * n) // 3 else master
        x = int(np.clip(x + rng.integers(-5, 6), 0, 500))
        y = int(np.clip(y + rng.integers(-3, 4), 0, 500))
        out[i] = np.clip(m[y:y + LA_H, x:x + LA_W] + rng.integers(-2, 3, (LA_H, LA_W)), 0, 255).astype(np.uint8)
    return out


def ref_lookahead_types(frames, n, weightp, threads, size=None, opts=None, lib=None):
    """The reference's OWN lookahead stage with its stock control flow (oracle/ref_shim.c: xref_lookahead_types = steps 1-4 of
    x264_encoder_encode, encoder.c:3360-3445: x264_frame_copy_picture, x264_adaptive_quant_frame, x264_frame_init_lowres,
    x264_lookahead_put_frame / _get_frames -> x264_slicetype_decide / _analyse / macroblock_tree) over the first n pictures of
    the cyclic clip `frames`, flushed at the end.  threads > 1: the reference's sliced lookahead (lookahead-threads, at most
    X264_LOOKAHEAD_THREAD_MAX = 16), whose results differ slightly from the one-thread ones (slicetype.c:668).
    -> ([(display index, type)] in coded order, seconds)"""
    import _libs
    r = lib or _libs.ref()
    r.xref_lookahead_types.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    if opts is None:
        opts = LA_REF_OPTS_W if weightp else LA_REF_OPTS
    if threads > 1:
        opts += b":threads=%d:lookahead-threads=%d:sync-lookahead=0" % (threads, threads)
    w_, h_ = size or (LA_W, LA_H)
    hnd = r.xref_open(w_, h_, b"medium", opts, 0)
    assert hnd
    clip = np.ascontiguousarray(np.stack([frames[i % len(frames)] for i in range(n)]))
    idx, ty = (C.c_int * n)(), (C.c_int * n)()
    t0 = time.perf_counter()
    k = r.xref_lookahead_types(hnd, clip.ctypes.data, n, idx, ty)
    dt = time.perf_counter() - t0
    r.xref_close(hnd)
    assert k == n, (k, n)
    return [(int(idx[i]), int(ty[i])) for i in range(k)], dt


def cpu_lookahead_rate(frames, budget_s, weightp=0, threads=1):
    """the same decision workload on the host cores: the unmodified reference's lookahead stage (ref_lookahead_types), or --
    only if oracle/_ref did not travel -- the product's host logic over the oracle port"""
    import _libs
    threads = max(1, min(int(threads), 16))
    if _libs.have_ref():
        # pictures to feed: enough for a steady state (several times the 40-picture lookahead), bounded by the budget
        est = 25.0 if threads > 1 else 6.0                   # pictures/s seen on this pool's hosts
        n = int(max(56, min(4 * len(frames), budget_s * est)))
        types, dt = ref_lookahead_types(frames, n, weightp, threads)
        return n / dt, "reference", threads, "%d 4K pictures through the reference's own lookahead stage (stock x264_slicetype_decide control flow, " \
            "cyclic %d-picture clip, flushed) in %.1f s, %d lookahead thread%s" % (n, len(frames), dt, threads, "s" if threads > 1 else ""), types
    from x264_b200.binding_ext import SlicetypeParams, LookaheadParams
    lib, kind, threads = _libs.slicetype_oracle_lib(), "port", 1
    la = LookaheadParams(LA_W, LA_H, *[LA_OPTS[k] for k in ("subpel_refine", "me_method", "me_range", "mv_range", "bframes",
                                                            "bframe_bias", "weighted_bipred", "aq_mode", "mb_tree", "vbv")], 0, int(weightp))
    p = SlicetypeParams(la, *[dict(LA_ST, psy=1 if weightp else 0)[k] for k in ("keyint_max", "keyint_min", "scenecut_threshold", "b_adapt", "b_pyramid",
                                                  "rc_lookahead", "psy", "frame_reference", "rc_cqp")])
    lib.x264cu_slicetype_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.x264cu_slicetype_step.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.x264cu_slicetype_close.argtypes = [C.c_void_p]
    st = C.c_void_p()
    assert lib.x264cu_slicetype_open(C.c_void_p(1), C.byref(p), C.byref(st)) == 0
    fr, ty = C.c_int(), C.c_int()
    t0 = time.perf_counter()
    fed = decided = 0
    types = []
    while fed < 4 * len(frames) and (time.perf_counter() - t0 < budget_s or decided < 2):
        f = frames[fed % len(frames)]
        assert lib.x264cu_slicetype_step(st, f.ctypes.data, f.shape[1], None, C.byref(fr), C.byref(ty)) == 0
        fed += 1
        if fr.value >= 0:
            decided += 1
            types.append((fr.value, ty.value))
    while True:
        assert lib.x264cu_slicetype_step(st, None, 0, None, C.byref(fr), C.byref(ty)) == 0
        if fr.value < 0:
            break
        decided += 1
        types.append((fr.value, ty.value))
    t = time.perf_counter() - t0
    lib.x264cu_slicetype_close(st)
    return decided / t, kind, threads, "%d 4K pictures decided in %.1f s by the oracle port, 1 thread" % (decided, t), types


C3_W, C3_H, C3_CLIP, C3_STEP = 7680, 4320, 48, 32          # a "step" of the configs[3] stream: 32 pictures, as at 4K
C3_REF_OPTS = b"weightp=0:no-psy=1:aq-mode=0:bframes=16:b-adapt=2:rc-lookahead=250"     # the same configuration, reference spelling
C3_OPTS = dict(subpel_refine=7, me_method=1, me_range=16, mv_range=512, bframes=16, bframe_bias=0, weighted_bipred=1, aq_mode=0, mb_tree=1, vbv=0)
C3_ST = dict(keyint_max=250, keyint_min=25, scenecut_threshold=40, b_adapt=2, b_pyramid=2, rc_lookahead=250, psy=0, frame_reference=3, rc_cqp=0)


def run_config3(ctx, x, rank, world, local, dist, args):
    """BASELINE configs[3]: ONE 7680x4320 stream, --rc-lookahead 250 --bframes 16 --b-adapt 2, sharded over the GPUs -- next to the same
    stream on ONE GPU, measured by rank 0 alone in the same run (the other ranks wait).  The WHOLE stream is timed, first picture in
    to last decision out: the host runs far ahead of the device here (a picture's 33 searches take ~24 ms of one GPU at 8K and the
    decisions arrive in bursts), so no window shorter than the stream is a steady state."""
    if dist is not None:
        import torch
        from x264_b200 import dist as xd
    steps = max(args.steps, 10)                       # at least 320 pictures, so that the 250-picture lookahead fills
    total = C3_STEP * steps
    frames = make_la_frames(4320, C3_CLIP, lambda b: ctx.malloc_host(b), C3_W, C3_H)
    d_frames = ctx.malloc(frames.nbytes + 256)
    ctx.h2d(d_frames, frames)

    def run(shard, host_input, n_pic):
        st = x.Slicetype(ctx, C3_W, C3_H, **C3_ST, **C3_OPTS, weighted_pred=0)
        ex = None
        if shard:
            ex = xd.ShardExchange(dist, device=torch.device("cuda", local))
            st.set_shard(rank, world, ex)
            barrier(dist, local)
        if host_input:
            st.set_async_upload(4)
        types = []
        ctx.sync()
        t0 = time.perf_counter()
        ctx.timer_start()
        for i in range(n_pic):
            fr, ty = st.step(frames[i % C3_CLIP]) if host_input else st.step_device(d_frames + (i % C3_CLIP) * C3_W * C3_H, C3_W)
            if fr >= 0:
                types.append((fr, ty))
        while len(types) < n_pic:
            fr, ty = st.step(None)
            if fr < 0:
                break
            types.append((fr, ty))
        ms = ctx.timer_stop()
        ctx.sync()
        wall = time.perf_counter() - t0
        if shard:
            ms = max_over_ranks(dist, ms, local)
            wall = max_over_ranks(dist, wall, local)
        spec = st.speculation_stats()
        busy = st.search_stats()
        st.close()
        return n_pic / (ms * 1e-3), n_pic / wall, types, spec, ex, busy, ms

    if dist is None:
        # one GPU (the N = 1 line carries this as an extra key, so that the 1 -> N series of this workload is complete)
        run(False, False, C3_STEP * 3)
        one = run(False, False, total)
        e2e = run(False, True, total)
        ctx.free(d_frames)
        return {"workload": "BASELINE configs[3]: ONE 7680x4320 stream, rc-lookahead 250, bframes 16, b-adapt 2, b-pyramid, mb-tree, scenecut 40 "
                            "(33 lowres searches per picture) on ONE GPU; the whole stream of %d pictures, first picture in to last decision "
                            "out; what `value` is at N > 1" % total,
                "pictures": total, "frames_per_s_1_gpu": one[0], "wall_frames_per_s_1_gpu": one[1], "e2e_frames_per_s_1_gpu": e2e[1],
                "e2e_same_decisions": e2e[2] == one[2],
                "searches": {"searches": one[5][2], "launches": one[5][1], "device_ms_of_search_launches": one[5][0]},
                "cost_requests": {"computed_ahead": one[3][0], "served_from_them": one[3][1], "computed_on_demand": one[3][2]}}
    for _ in range(1):                                 # warm-up: a short sharded stream (allocator, NCCL channels, clocks)
        run(True, False, C3_STEP * min(max(args.warmup, 3), 4))
    one = None
    if rank == 0:
        one = run(False, False, total)
    barrier(dist, local)
    many = run(True, False, total)
    e2e = run(True, True, total)
    sig = torch.tensor([hash(tuple(many[2])) & 0x7fffffffffff], dtype=torch.int64, device=torch.device("cuda", local))
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    ctx.free(d_frames)
    out = {"workload": "BASELINE configs[3]: ONE 7680x4320 stream, rc-lookahead 250, bframes 16, b-adapt 2, b-pyramid, mb-tree, scenecut 40 "
                       "(33 lowres searches per picture), sharded over the GPUs; the whole stream of %d pictures (%d steps of %d), first "
                       "picture in to last decision out (pipeline fill and flush included); cyclic %d-picture clip resident in "
                       "HBM" % (total, steps, C3_STEP, C3_CLIP),
           "steps": steps, "pictures": total, "pictures_decided": len(many[2]), "ms": many[6],
           "searches_on_this_rank": {"searches": many[5][2], "launches": many[5][1], "device_ms_of_search_launches": many[5][0]},
           "frames_per_s_%d_gpus" % world: many[0], "wall_frames_per_s_%d_gpus" % world: many[1],
           "e2e_frames_per_s_%d_gpus" % world: e2e[1], "e2e_same_decisions": e2e[2] == many[2],
           "h2d_bytes_per_step": int(C3_STEP * C3_W * C3_H), "d2h_bytes_per_step": int(32 * many[3][0] / steps),
           "same_decisions_on_every_rank": bool(all(int(t.item()) == int(sig.item()) for t in sigs)),
           "exchanges": many[4].calls, "MB_per_exchange": many[4].bytes / max(many[4].calls, 1) / 1e6,
           "cost_requests": {"computed_ahead": many[3][0], "served_from_them": many[3][1], "computed_on_demand": many[3][2]}}
    if one is not None:
        out["frames_per_s_1_gpu"], out["wall_frames_per_s_1_gpu"] = one[0], one[1]
        out["one_gpu_searches"] = {"searches": one[5][2], "launches": one[5][1], "device_ms_of_search_launches": one[5][0]}
        out["speedup"] = many[0] / one[0]
        out["same_decisions_as_1_gpu"] = one[2] == many[2]
    return out


def run_lookahead_b200(args, rank, world, local, dist):
    import x264_b200 as x
    ctx = x.Context(local)
    info = ctx.device_info()
    n = LA_FRAMES
    frames = make_la_frames(2160 + rank, n, lambda b: ctx.malloc_host(b))
    stride = LA_W
    d_frames = ctx.malloc(frames.nbytes + 256)
    ctx.h2d(d_frames, frames)

    la_st = dict(LA_ST, psy=1) if args.weightp else LA_ST

    def make_st():
        st_ = x.Slicetype(ctx, LA_W, LA_H, **la_st, **LA_OPTS, weighted_pred=args.weightp)
        if args.group:
            st_.set_prefetch_group(args.group)
        if args.run_ahead is not None:
            st_.set_run_ahead(args.run_ahead)
        return st_

    # ---- device-resident pictures -----------------------------------------------------------------
    st = make_st()
    decided = []

    def step_dev():
        for i in range(n):
            fr, ty = st.step_device(d_frames + i * LA_W * LA_H, stride)
            if fr >= 0:
                decided.append((fr, ty))

    for _ in range(max(args.warmup, 3)):
        step_dev()
    ctx.sync()
    ss0 = st.search_stats()
    sampler = ClockSampler(local)
    barrier(dist, local)
    sampler.start()
    l0, r0 = ctx.launches, st.cost_requests
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms = ctx.timer_stop()
    ctx.sync()
    barrier(dist, local)
    wall = time.perf_counter() - t0
    launches, requests = ctx.launches - l0, st.cost_requests - r0
    ms = max_over_ranks(dist, ms, local)
    clocks = sampler.stop()
    ss1 = st.search_stats()            # CUDA events around every search launch, on the (low-priority) stream it ran on
    live_launches, live_searches, live_ms = ss1[1] - ss0[1], ss1[2] - ss0[2], ss1[0] - ss0[0]
    st.close()
    # the one exchange step: every rank's per-picture decision records are all-gathered (NCCL) so that rank 0 holds the
    # decisions of all streams; outside the timed region only because it happens once per run, not per step
    gathered_streams = 1
    if dist is not None:
        import torch
        from x264_b200 import dist as xd
        rec = xd.pack_records(rank, decided[-(n * args.steps):], n * args.steps)
        allrec = xd.unpack_records(xd.all_gather_records(dist, rec, device=torch.device("cuda", local)))
        gathered_streams = len(allrec)

    # ---- ONE stream sharded over the GPUs: every rank is fed rank 0's pictures; the searches AND the cost requests of
    # each prefetch group are split by picture, two all-gathers per group (search results; cost records + lowres_costs).  At N > 1
    # this is the line's `value` (strong scaling of the N = 1 workload); the independent-streams figure above becomes `replicas`.
    sharded = None
    if dist is not None:
        import torch
        from x264_b200 import dist as xd
        same_frames = make_la_frames(2160, n, lambda b: ctx.malloc_host(b))
        ctx.h2d(d_frames, same_frames)

        def sharded_run(host_input, steps):
            st = make_st()
            ex = xd.ShardExchange(dist, device=torch.device("cuda", local))
            st.set_shard(rank, world, ex)
            if host_input:
                st.set_async_upload(4)
            types = []

            def one_step():
                for i in range(n):
                    fr, ty = st.step(same_frames[i]) if host_input else st.step_device(d_frames + i * LA_W * LA_H, stride)
                    if fr >= 0:
                        types.append((fr, ty))

            for _ in range(3):
                one_step()
            ctx.sync()
            barrier(dist, local)
            t0_ = time.perf_counter()
            ctx.timer_start()
            for _ in range(steps):
                one_step()
            ms_ = max_over_ranks(dist, ctx.timer_stop(), local)
            ctx.sync()
            barrier(dist, local)
            wall_ = max_over_ranks(dist, time.perf_counter() - t0_, local)
            spec = st.speculation_stats()
            st.close()
            return ms_, wall_, types, ex, spec

        sh_steps = max(1, args.steps)
        sh_ms, sh_wall, sh_types, ex, sh_spec = sharded_run(False, sh_steps)
        e2e_sh_steps = 3 if args.quick else max(1, min(args.steps, 10))
        _, e2e_sh_wall, e2e_sh_types, _, _ = sharded_run(True, e2e_sh_steps)
        # every rank must have taken the same decisions
        sig = torch.tensor([hash(tuple(sh_types)) & 0x7fffffffffff], dtype=torch.int64, device=torch.device("cuda", local))
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        sharded = {"value": n * sh_steps / (sh_ms * 1e-3), "unit": "frames/s", "steps": sh_steps, "ms_per_step": sh_ms / sh_steps,
                   "wall_frames_per_s": n * sh_steps / sh_wall,
                   "e2e": n * e2e_sh_steps / e2e_sh_wall,
                   "cost_requests": {"computed_ahead": sh_spec[0], "served_from_them": sh_spec[1], "computed_on_demand": sh_spec[2]},
                   "note": "ONE 4K stream over %d GPUs: searches and cost requests split by picture, decisions replicated, two NCCL all-gathers "
                           "per 12-picture group (%.1f MB gathered per exchange on average, %d exchanges)" % (world, ex.bytes / max(ex.calls, 1) / 1e6, ex.calls),
                   "same_decisions_on_every_rank": bool(all(int(t.item()) == int(sig.item()) for t in sigs))}
        ctx.h2d(d_frames, frames)

    # ---- the search kernel alone: 8 searches (4 distances x 2 lists) of one picture per launch --------
    la = x.Lookahead(ctx, LA_W, LA_H, n_slots=8, **LA_OPTS)
    jobs = [(4, 4 - d, 0, d) for d in range(1, 5)] + [(4 - d, 4, 1, d) for d in range(1, 4)]
    # the prefetcher's real launches carry the searches of 4 pictures: time that shape too
    jobs4 = [(b, b - d, 0, d) for b in range(4, 8) for d in range(1, 5)] + [(b, b + d, 1, d) for b in range(1, 5) for d in range(1, 4)]
    reps = 5

    def time_jobs(jl):
        out = []
        for r in range(reps + 1):
            for i in range(8):
                la.frame_put_device(i, d_frames + i * LA_W * LA_H, stride)     # resets the memo -> searches run again
            ctx.sync()
            ctx.timer_start()                                                    # event on the context's stream ...
            la.search_batch(jl)                                                  # ... which the search stream is ordered after
            la.join()                                                            # and the context's stream after the searches
            out.append(ctx.timer_stop())
        return float(np.median(out[1:]))

    k_ms = [time_jobs(jobs)]
    search4_ms = time_jobs(jobs4)
    la.close()
    search_ms = float(np.median(k_ms))

    # ---- e2e: host pictures through the public entry point, H2D inside --------------------------------
    st = make_st()
    st.set_async_upload(4)        # the pictures live in page-locked memory and are not touched while queued (as x264 holds its frames)
    all_types = []                # every decision from picture 0 on: compared with the reference's own below
    for i in range(n):
        fr, ty = st.step(frames[i])
        if fr >= 0:
            all_types.append((fr, ty))
    barrier(dist, local)
    t1 = time.perf_counter()
    e2e_steps = 3 if args.quick else max(1, min(args.steps, 10))
    e2e_types = []
    for _ in range(e2e_steps):
        for i in range(n):
            fr, ty = st.step(frames[i])
            if fr >= 0:
                e2e_types.append((fr, ty))
                all_types.append((fr, ty))
    ctx.sync()
    barrier(dist, local)
    e2e_s = max_over_ranks(dist, time.perf_counter() - t1, local)
    st.close()

    ms_step = ms / args.steps
    peaks, peak_src = measured_peaks()
    n_jobs = len(jobs)
    # the dominant kernel as it ran inside the timed region: average searches per launch and average launch duration
    per_launch = live_searches / max(1, live_launches)
    launch_ms = live_ms / max(1, live_launches)
    achieved = LA_SEARCH_BYTES * per_launch / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic("search_kernel<8>", "searches_per_launch", per_launch)
    res = {
        "metric": "lowres_lookahead_frames_per_sec_4k", "value": n * world / (ms_step * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "3840x2160 lowres lookahead + slice-type decision, %d pictures per step per GPU, preset-medium "
                               "lookahead settings (hex, subme 7 -> lookahead subpel 4, bframes 3, b-adapt 1, rc-lookahead 40, "
                               "mb-tree requests, scenecut 40; lookahead weightp analysis %s, aq off)" % (n, "on" if args.weightp else "off"),
                   "l2": "each picture's 4 lowres planes (9.4 MB) stay L2-resident by design; pictures cycle through %d MB" % (frames.nbytes // 2**20),
                   "cost_requests_per_step": requests / args.steps, "decided_per_step": len(decided) / (args.steps + max(args.warmup, 3)),
                   "scheduling": "searches prefetched on two low-priority streams in groups of 12 pictures (72 searches per launch: list 0 at distances 1-4, list 1 at distances 1-2 -- with a B pyramid nobody reads further), decisions run 24 pictures behind the newest one (sync-lookahead twin); uploads on their own stream",
                   "multi_gpu": "value / e2e: one independent stream per GPU (weak scaling), one NCCL all-gather of decision records (%d streams gathered); "
                                "sharded_stream: ONE stream over all GPUs" % gathered_streams},
        "clocks": clocks,
        "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes),
                "d2h_bytes_per_step": int(32 * requests / args.steps), "api": "x264cu_slicetype_step (page-locked host luma read in place by the copy engine on the upload stream, async_upload=4)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "search_kernel<8> (%.1f searches per launch, %d launches in the timed region, two launches in flight)" % (per_launch, live_launches),
                     "algorithmic_bytes_per_launch": LA_SEARCH_BYTES * per_launch, "ms_per_launch": launch_ms,
                     "share_of_step": live_ms / max(ms, 1e-9),
                     "ms_per_launch_alone_%d_searches" % n_jobs: search_ms, "ms_per_launch_alone_%d_searches" % len(jobs4): search4_ms,
                     "note": "dependency-bound wavefront (508 pipeline steps at 4K), not a streaming kernel: the planes of the ~13 "
                             "pictures a launch touches stay in L2 (traffic << algorithmic bytes); see DESIGN.md and issue_model",
                     # what actually bounds this kernel: warp-instruction issue.  1 334 warp instructions per macroblock search (ncu,
                     # profiles/r01b_search_kernel_ncu_full.txt) against the SMs' issue slots (4 schedulers per SM, one per clock)
                     "issue_model": (lambda mbs, clk: {
                         "warp_instructions_per_mb_search": LA_WARP_INSTR_PER_MB_SEARCH, "mb_searches_in_timed_region": live_searches * mbs,
                         "issue_slots_per_s": info["sm_count"] * 4 * clk * 1e6,
                         "min_ms_at_full_issue": live_searches * mbs * LA_WARP_INSTR_PER_MB_SEARCH / (info["sm_count"] * 4 * clk * 1e6) * 1e3,
                         "timed_region_ms": ms,
                         "frac_of_issue_peak": live_searches * mbs * LA_WARP_INSTR_PER_MB_SEARCH / (info["sm_count"] * 4 * clk * 1e6) / (ms * 1e-3),
                         "note": "the whole step's wall time against the search kernels' instruction count alone (the other kernels' "
                                 "instructions are not counted): ncu's own figure for the kernel in isolation is 43 % issue slots busy"})(
                         ((LA_W // 2 + 7) // 8) * ((LA_H // 2 + 7) // 8), float((clocks or {}).get("sm_mhz") or 1965.0))},
        "wall_s": wall, "sm_count": info["sm_count"],
    }
    if sharded is not None:
        res["replicas"] = {"value": res["value"], "unit": "frames/s", "ms_per_step": res["ms_per_step"], "e2e": res["e2e"]["value"], "scaling": "weak",
                           "note": "N independent 4K streams, one per GPU, no data-path collective (one all-gather of decision records)"}
        res["value"], res["ms_per_step"], res["scaling"] = sharded["value"], sharded["ms_per_step"], "strong"
        res["e2e"] = dict(res["e2e"], value=sharded["e2e"], api="x264cu_slicetype_step on every rank (sharded stream: each rank is fed the same page-locked host pictures)")
        res["sharded_stream"] = sharded
        if not args.quick:
            # N > 1, full run: the line's value is BASELINE configs[3] -- the 8K stream whose lookahead is what the sharding is for
            # (33 searches per picture instead of 7); the 4K figures above stay as `sharded_stream` and `replicas`
            c3 = run_config3(ctx, x, rank, world, local, dist, args)
            res["config3_8k"] = c3
            res["sharded_stream_4k"] = dict(sharded, e2e_api=res["e2e"]["api"])
            res["metric"] = "lowres_lookahead_frames_per_sec_8k_sharded_stream"
            res["value"], res["steps"], res["ms_per_step"], res["scaling"] = c3["frames_per_s_%d_gpus" % world], c3["steps"], c3["ms"] / c3["steps"], "strong"
            res["value_1_gpu_same_workload"] = c3.get("frames_per_s_1_gpu")
            res["speedup_vs_1_gpu_same_workload"] = c3.get("speedup")
            res["config"]["workload"] = c3["workload"]
            res["config"]["n1_line"] = "the N = 1 line of this bench is the 4K stream (BASELINE configs[1]-style); value_1_gpu_same_workload is the "\
                                       "one-GPU figure for THIS workload, measured by rank 0 in this run -- the 1 -> N ratio to read"
            res["config"]["l2"] = "17 pictures' lowres planes (35 MB each) are live per search group: far beyond the 126 MB L2"
            res["roofline"]["measured_on"] = "the 4K leg of this run (replicas): the dominant kernel is the same search_kernel<8>; gpu_launches and clocks are that leg's too"
            res["e2e"] = {"value": c3["e2e_frames_per_s_%d_gpus" % world], "unit": "frames/s", "h2d_bytes_per_step": c3["h2d_bytes_per_step"],
                          "d2h_bytes_per_step": c3["d2h_bytes_per_step"],
                          "api": "x264cu_slicetype_step on every rank (each rank is fed the same page-locked host pictures; whole stream, host clock)"}
    if world == 1 and not args.quick:
        res["config3_8k_one_gpu"] = run_config3(ctx, x, rank, world, local, None, args)
    if rank == 0 and world == 1 and not args.quick:
        rate, kind, cores, sample, _ = cpu_lookahead_rate(frames, args.cpu_budget, args.weightp, os.cpu_count() or 1)
        rate1, _, _, sample1, types1 = cpu_lookahead_rate(frames, args.cpu_budget, args.weightp, 1)
        # in-run parity: the reference's one-thread decisions (its documented deterministic mode) against the B200 arm's on the
        # same cyclic clip; the reference was flushed after its last picture, so only decisions taken with a full lookahead count
        n_ref = len(types1)
        keep = max(0, n_ref - LA_ST["rc_lookahead"] - LA_OPTS["bframes"] - 2)
        got = [t_ for t_ in all_types][:keep]
        res["config"]["parity_spot_check"] = {"decisions_compared": keep, "identical": bool(keep > 0 and got == types1[:keep]),
                                              "against": "reference x264_slicetype_decide, 1 lookahead thread, %d pictures" % n_ref}
        vec = None
        import _libs
        if kind == "reference" and _libs.ref_vec() is not None:
            n_v = int(max(56, min(4 * len(frames), args.cpu_budget * 25.0)))
            _, dt_v = ref_lookahead_types(frames, n_v, args.weightp, cores, lib=_libs.ref_vec())
            vec = {"value": n_v / dt_v, "unit": "frames/s", "cores": cores, "sample": "%d 4K pictures in %.1f s" % (n_v, dt_v),
                   "build": "the same C auto-vectorised (oracle/Makefile.ref vec: -O3 -ffast-math -ftree-vectorize -march=x86-64-v3); "
                            "the reference's x86 asm cannot be assembled here (no nasm / yasm in the image or the wheelhouse)"}
        res["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                               "value_1_thread": rate1, "sample_1_thread": sample1, "autovectorized_c": vec,
                               "note": "the unmodified reference's lookahead stage (stock control flow) with its sliced lookahead threads on all "
                                       "host cores it can use (max 16); with one thread it returns exactly the decisions the B200 arm is "
                                       "checked against (parity_spot_check)"}
    ctx.close()
    return res


def run_lookahead_reference_config3(args, world):
    """N > 1: the b200 arm's line is BASELINE configs[3]; the same configuration through the unmodified reference's lookahead stage
    on the host cores.  A bounded sample: 36 pictures (two 17-picture mini-GOPs), at most two samples -- one takes about a minute."""
    import _libs
    assert _libs.have_ref(), "oracle/_ref did not travel"
    threads = max(1, min(os.cpu_count() or 1, 16))
    frames = make_la_frames(4320, 16, lambda b: np.empty(b, np.uint8), C3_W, C3_H)
    n, rates, dts = 36, [], []
    for _ in range(max(1, min(args.steps, 2))):
        types, dt = ref_lookahead_types(frames, n, 0, threads, size=(C3_W, C3_H), opts=C3_REF_OPTS)
        rates.append(n / dt)
        dts.append(dt)
    v = float(np.mean(rates))
    sample = "%d 8K pictures (cyclic 16-picture clip, flushed: the 250-picture lookahead never fills, every decision sees the whole sample) " \
             "through the reference's own lookahead stage, %d lookahead threads, %d sample(s) of %.0f s" % (n, threads, len(rates), float(np.mean(dts)))
    return {
        "impl": "reference", "metric": "lowres_lookahead_frames_per_sec_8k_sharded_stream", "value": v, "unit": "frames/s",
        "n_gpus": world, "steps": len(rates), "warmup": 0, "ms_per_step": C3_STEP / v * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[3]: 7680x4320, rc-lookahead 250, bframes 16, b-adapt 2 (preset medium otherwise, weightp / aq / psy off "
                               "as in the b200 arm), bounded sample"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the unmodified reference's own lookahead stage (C path: no nasm in the image) on the host cores; it does not use the GPUs, "
                "so the figure is the same at every N",
    }


def run_lookahead_reference(args, rank, world):
    if world > 1 and not args.quick:
        return run_lookahead_reference_config3(args, world)
    frames = make_la_frames(2160, LA_FRAMES, lambda b: np.empty(b, np.uint8))
    rates = []
    for i in range(args.warmup + args.steps):
        r, kind, cores, sample, _ = cpu_lookahead_rate(frames, args.cpu_budget / max(1, args.steps), args.weightp, os.cpu_count() or 1)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    return {
        "impl": "reference", "metric": "lowres_lookahead_frames_per_sec_4k", "value": v, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": LA_FRAMES / v * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "3840x2160 lowres lookahead + slice-type decision, bounded sample per step, same settings as the b200 arm"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the unmodified reference's own lookahead stage (x264_lookahead_put_frame / _get_frames -> x264_slicetype_decide, "
                "encoder/lookahead.c, encoder/slicetype.c; C path: no nasm in the image) with its sliced lookahead threads on all host cores (max 16)",
    }


# ------------------------------------------------------------------------------------------------------
# workload "me": full-resolution motion estimation (BASELINE configs[2]): the x264_me_t stream of a 3840x2160 --preset slower
# --me umh --merange 64 encode, recorded from the reference encoder on this box, replayed through x264cu_me_search_frame
# ------------------------------------------------------------------------------------------------------
ME_OPTS = b"me=umh:merange=64:threads=1"
ME_FRAMES_IN, ME_TRACED = 6, 3


def me_record(seed=2160):
    import _me_trace as T
    t0 = time.perf_counter()
    frames = T.record(W4K, H4K, ME_FRAMES_IN, ME_OPTS, max_frames=ME_TRACED, skip=1, seed=seed)
    return T, frames, time.perf_counter() - t0


def me_check(T, t, got):
    want, ee = T.expected(t), T.early_exit(t)
    bad = (got[:, [0, 1, 2, 4]] != want[:, [0, 1, 2, 4]]).any(1) | ((got[:, 3] != want[:, 3]) & ~ee)
    return int(bad.sum())


def me_algorithmic_bytes(t):
    """every plane a picture's searches can read once (source luma + chroma, per reference 4 luma planes + chroma + the weighted
    plane), the job records in and the results out"""
    luma, chroma = t.stride * t.lines, t.stride_uv * t.lines_uv
    b = luma + chroma
    for rf in t.refs:
        b += 4 * luma + chroma + (luma if rf["weighted"] else 0)
    return b + t.n_recs * (72 + 16)


def run_me_b200(args, rank, world, local, dist):
    import x264_b200 as x
    T, frames, rec_s = me_record()
    ctx = x.Context(local)
    info = ctx.device_info()
    devs = [T.DeviceTrace(ctx, t, x) for t in frames]
    n_search = sum(d.n for d in devs)

    def step():
        for d in devs:
            d.launch()

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()
    bad = sum(me_check(T, t, d.results()) for t, d in zip(frames, devs))
    sampler = ClockSampler(local)
    barrier(dist, local)
    sampler.start()
    l0 = ctx.launches
    t_wall0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    ctx.sync()
    barrier(dist, local)
    wall = time.perf_counter() - t_wall0
    launches = ctx.launches - l0
    ms = max_over_ranks(dist, ms, local)
    clocks = sampler.stop()
    per_pic = []
    for d in devs:                                  # each picture alone (its own launch duration)
        ctx.sync()
        ctx.timer_start()
        for _ in range(5):
            d.launch()
        per_pic.append(ctx.timer_stop() / 5)

    # e2e: planes, job records in page-locked host memory -> HBM -> searches -> results back on the host, every step
    host = []
    for t in frames:
        bufs = [t.fenc, t.fenc_uv, t.recs]
        for rf in t.refs:
            bufs += rf["planes"] + [rf["uv"]] + ([rf["wplane"]] if rf["weighted"] else [])
        pinned = []
        for b_ in bufs:
            hb = ctx.malloc_host(b_.nbytes)
            hb[:] = b_.view(np.uint8).reshape(-1)
            pinned.append(hb)
        host.append(pinned)
    h2d = sum(b_.nbytes for hb in host for b_ in hb)
    e2e_steps = 1 if args.quick else max(1, min(args.steps, 5))

    e2e_devs = []
    for t, hb in zip(frames, host):
        t2 = T.TraceFrame()
        t2.__dict__.update(t.__dict__)
        it = iter(hb)
        t2.fenc, t2.fenc_uv = next(it), next(it)
        t2.recs = np.frombuffer(next(it), T.REC)
        t2.refs = []
        for rf in t.refs:
            r2 = dict(rf)
            r2["planes"] = [next(it) for _ in range(4)]
            r2["uv"] = next(it)
            r2["wplane"] = next(it) if rf["weighted"] else None
            t2.refs.append(r2)
        e2e_devs.append(T.DeviceTrace(ctx, t2, x))         # device buffers allocated once; every step copies into them again

    def step_e2e(check=False):
        nbad = 0
        for t, d in zip(frames, e2e_devs):
            d.reupload()
            d.launch()
            got = d.results()
            if check:
                nbad += me_check(T, t, got)
        return nbad

    bad += step_e2e(check=True)
    barrier(dist, local)
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    ctx.sync()
    barrier(dist, local)
    e2e_s = max_over_ranks(dist, time.perf_counter() - t1, local)

    ms_step = ms / args.steps
    peaks, peak_src = measured_peaks()
    alg = sum(me_algorithmic_bytes(t) for t in frames)
    achieved = alg / (ms_step * 1e-3) / 1e9
    res = {
        "metric": "fullres_me_searches_per_sec_4k", "value": n_search * world / (ms_step * 1e-3), "unit": "searches/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: 3840x2160 --preset slower --me umh --merange 64 (subme 9, chroma ME, ref 8, weightp 2, "
                               "aq 1): EVERY x264_me_t the reference encoder passed to x264_me_search_ref while coding pictures %s of a "
                               "%d-picture synthetic clip (recorded on this box in %.1f s of CPU), %d searches per step, all partition sizes, "
                               "multi-ref, per-macroblock lambdas" % ([("PB"[t.slice_type], t.display) for t in frames], ME_FRAMES_IN, rec_s, n_search),
                   "pictures": [{"coded": t.coded, "type": "PB"[t.slice_type], "searches": d.n, "refs": t.n_refs, "ms": m_}
                                for t, d, m_ in zip(frames, devs, per_pic)],
                   "l2": "one step reads %.0f MB of planes and job records > 126 MB L2" % (alg / 1e6),
                   "parity": "%d of %d searches differ from the reference's recorded (mv, cost, cost_mv, threshold)" % (bad, 2 * n_search),
                   "multi_gpu": "replicas only (full-res ME depends on reconstructed references): every rank replays the stream"},
        "clocks": clocks,
        "e2e": {"value": n_search * world * e2e_steps / e2e_s, "unit": "searches/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(16 * n_search), "api": "x264cu_me_search_frame (page-locked host planes and jobs uploaded every step)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peak_src, "kernel": "me_search_kernel<false>", "algorithmic_bytes_per_launch": alg / len(frames),
                     "note": "one warp per search walking the reference's own candidate order: bounded by L1/issue latency, not by HBM; "
                             "see profiles/ for the ncu summary"},
        "wall_s": wall, "sm_count": info["sm_count"], "parity_ok": bad == 0,
    }
    if rank == 0 and world == 1 and not args.quick:
        cores = os.cpu_count() or 1
        t = frames[0]
        n_s = min(t.n_recs, 40000)
        _, dt1 = T.replay_reference(t, ME_OPTS, 1, sel=slice(0, n_s // 8))
        big = max(frames, key=lambda f: f.n_recs)
        n_b = min(big.n_recs, int(args.cpu_budget * 2500 * cores))
        got, dt = T.replay_reference(big, ME_OPTS, cores, sel=slice(0, n_b))
        res["cpu_baseline"] = {"value": n_b / dt, "unit": "searches/s", "cores": cores, "kind": "reference",
                               "sample": "x264_me_search_ref (C path: no nasm in the image) on the first %d recorded searches of coded picture %d, "
                                         "%d threads with one encoder handle each" % (n_b, big.coded, cores),
                               "value_1_thread": (n_s // 8) / dt1}
    for d in devs + e2e_devs:
        d.close()
    ctx.close()
    return res


def run_me_reference(args, rank, world):
    T, frames, rec_s = me_record()
    cores = os.cpu_count() or 1
    big = max(frames, key=lambda f: f.n_recs)
    rates = []
    n_b = 0
    for i in range(args.warmup + args.steps):
        n_b = min(big.n_recs, int(args.cpu_budget / max(1, args.steps) * 2500 * cores))
        _, dt = T.replay_reference(big, ME_OPTS, cores, sel=slice(0, n_b))
        if i >= args.warmup:
            rates.append(n_b / dt)
    v = float(np.mean(rates))
    n_search = sum(t.n_recs for t in frames)
    return {
        "impl": "reference", "metric": "fullres_me_searches_per_sec_4k", "value": v, "unit": "searches/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_search / v * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: the same recorded x264_me_t stream; bounded sample of %d searches per step" % n_b},
        "cpu_baseline": {"value": v, "unit": "searches/s", "cores": cores, "kind": "reference",
                         "sample": "x264_me_search_ref on the first %d recorded searches of coded picture %d" % (n_b, big.coded)},
        "e2e": {"value": v, "unit": "searches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference x264_me_search_ref (encoder/me.c:182, C path: no nasm in the image), one encoder handle per host thread",
    }


# ------------------------------------------------------------------------------------------------------
# workload "sweep": BASELINE configs[4], the SAD / SATD / SSD microbench over the block sizes of x264_pixel_function_t
# ------------------------------------------------------------------------------------------------------
SWEEP_SIZES = [(6, "4x4"), (5, "4x8"), (4, "8x4"), (3, "8x8"), (2, "8x16"), (1, "16x8"), (0, "16x16")]
SWEEP_METRICS = [(0, "sad"), (2, "satd"), (1, "ssd")]


def run_sweep_b200(args, rank, world, local, dist):
    import x264_b200 as x
    import _libs
    ctx = x.Context(local)
    info = ctx.device_info()
    fenc, ref, mv16, stride, pitch = make_satd_inputs(1 + rank, N_PAIRS, lambda n: ctx.malloc_host(n))
    d_fenc, d_ref = ctx.malloc(fenc.nbytes + 256), ctx.malloc(ref.nbytes + 256)
    ctx.h2d(d_fenc, fenc)
    ctx.h2d(d_ref, ref)
    org = PAD * stride + PAD
    pf = (d_fenc + org, stride, pitch, W4K, H4K, N_PAIRS)
    pr = (d_ref + org, stride, pitch, W4K, H4K, N_PAIRS)
    peaks, peak_src = measured_peaks()
    rng = np.random.default_rng(100 + rank)
    cells = {}
    sampler = ClockSampler(local)
    sampler.start()
    total_launches, t_all = 0, time.perf_counter()
    parity_ok = True
    cpu = rank == 0 and world == 1 and not args.quick
    cfn = (_libs.ref().xref_pixel_cmp_batch if _libs.have_ref() else _libs.oracle().orc_pixel_cmp_batch)
    vec_lib = _libs.ref_vec() if cpu else None
    for ip, sname in SWEEP_SIZES:
        bw, bh = x.PIXEL_W[ip], x.PIXEL_H[ip]
        nbx, nby = W4K // bw, H4K // bh
        n_cand = N_PAIRS * nbx * nby
        mv = rng.integers(-16, 17, (1, N_PAIRS, nby, nbx, 2)).astype(np.int16)
        d_mv = ctx.upload(mv)
        d_out = ctx.malloc(n_cand * 4)
        # checker's candidate list for plane 0 (bounded)
        by, bx = np.mgrid[0:nby, 0:nbx]
        c0 = np.zeros(nby * nbx, x.cand_dtype)
        c0["fenc_off"] = (org + by * bh * stride + bx * bw).reshape(-1)
        c0["ref_off"] = (org + (by * bh + mv[0, 0, :, :, 1]) * stride + bx * bw + mv[0, 0, :, :, 0]).reshape(-1)
        n_chk = min(len(c0), 20000)
        for mi, mname in SWEEP_METRICS:
            call = lambda: ctx.pixel_cmp_mvfield(mi, ip, pf, pr, 1, d_mv, d_out)
            for _ in range(3):
                call()
            ctx.sync()
            l0 = ctx.launches
            ctx.timer_start()
            for _ in range(args.steps):
                call()
            ms = max_over_ranks(dist, ctx.timer_stop(), local) / args.steps
            total_launches += ctx.launches - l0
            got = ctx.download(d_out, (n_cand,), np.int32)[:n_chk]
            want = np.zeros(n_chk, np.int32)
            cfn(mi, ip, fenc[:pitch], stride, ref[:pitch], stride, c0[:n_chk], n_chk, want)
            ok = bool(np.array_equal(got, want))
            parity_ok &= ok
            alg = (2 * bw * bh + 4) * n_cand
            cell = {"candidates": n_cand, "ms": ms, "blocks_per_s": n_cand * world / (ms * 1e-3), "GB_s": alg / (ms * 1e-3) / 1e9,
                    "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "parity": ok}
            if cpu:
                n_cpu = min(len(c0), 200000)
                outc = np.zeros(n_cpu, np.int32)
                t0 = time.perf_counter()
                cfn(mi, ip, fenc[:pitch], stride, ref[:pitch], stride, c0[:n_cpu], n_cpu, outc)
                cell["cpu_blocks_per_s_1_thread"] = n_cpu / (time.perf_counter() - t0)
                if vec_lib is not None:            # the same C, auto-vectorised for AVX2 (oracle/Makefile.ref vec): a labelled second column
                    t0 = time.perf_counter()
                    vec_lib.xref_pixel_cmp_batch(mi, ip, fenc[:pitch], stride, ref[:pitch], stride, c0[:n_cpu], n_cpu, outc)
                    cell["cpu_autovectorized_blocks_per_s_1_thread"] = n_cpu / (time.perf_counter() - t0)
            cells["%s_%s" % (mname, sname)] = cell
        ctx.free(d_mv)
        ctx.free(d_out)
    clocks = sampler.stop()
    wall = time.perf_counter() - t_all
    head = cells["satd_16x16"]
    res = {
        "metric": "satd_16x16_macroblocks_per_sec_4k", "value": head["blocks_per_s"], "unit": "macroblocks/s",
        "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": head["ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4]: SAD / SATD / SSD x 4x4 .. 16x16 over every block of %d random 3840x2160 luma frame "
                               "pairs per GPU (>= 1 036 800 candidates per cell), ref displaced by a random full-pel mv in +-16" % N_PAIRS,
                   "l2": "inputs %.0f MB per launch > 126 MB L2, no flush needed" % ((fenc.nbytes + ref.nbytes) / 1e6),
                   "parity_spot_check": parity_ok,
                   "cpu": "reference C path (pixf.sad/satd/ssd; the x86 asm cannot be assembled: no nasm / yasm in the image or the wheelhouse), one "
                          "thread, bounded sample; cpu_autovectorized_*: the same C built -O3 -ftree-vectorize -march=x86-64-v3" if cpu else None},
        "cells": cells, "clocks": clocks, "gpu_launches": int(total_launches),
        "e2e": None,
        "roofline": {"bound": "hbm", "achieved": head["GB_s"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": head["frac"],
                     "traffic": None, "peak_source": peak_src, "kernel": "mvfield_kernel<SATD,16,16>",
                     "algorithmic_bytes_per_launch": ALG_BYTES_16x16 * head["candidates"],
                     "note": "algorithmic bytes per candidate = 2*W*H + 4 (SURVEY 8d); every cell's fraction is in `cells`"},
        "wall_s": wall, "sm_count": info["sm_count"],
    }
    ctx.close()
    return res


WORKLOADS = {"satd": (run_satd_b200, run_satd_reference), "lookahead": (run_lookahead_b200, run_lookahead_reference),
             "me": (run_me_b200, run_me_reference), "sweep": (run_sweep_b200, run_satd_reference)}
DEFAULT_WORKLOAD = "lookahead"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--weightp", type=int, default=1, choices=[0, 1],
                    help="lookahead workload: the lookahead weight analysis (slicetype.c:284-501) and psy, as in preset medium (default); 0: off")
    ap.add_argument("--quick", action="store_true", help="tuning runs: skip the e2e and cpu_baseline legs")
    ap.add_argument("--group", type=int, default=0, help="lookahead workload: pictures per prefetch launch (0 = the library's default)")
    ap.add_argument("--run-ahead", type=int, default=None, help="lookahead workload: pictures queued beyond the lookahead before deciding")
    args = ap.parse_args()

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if rank != 0:
            return 0
        print(json.dumps(WORKLOADS[args.workload or DEFAULT_WORKLOAD][1](args, rank, world)))
        return 0

    rank, world, local, dist = dist_setup(args.gpus)
    res = WORKLOADS[args.workload or DEFAULT_WORKLOAD][0](args, rank, world, local, dist)
    if args.workload is None and not args.quick:
        # the second half of BASELINE.json's metric rides on the same line: 16x16 SATD macroblocks/s vs the HBM roofline
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup, a2.quick = 20, 3, True
        sat = run_satd_b200(a2, rank, world, local, dist)
        res["satd_16x16"] = {k: sat[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "roofline", "e2e", "gpu_launches")}
        res["satd_16x16"]["workload"] = sat["config"]["workload"]
        res["satd_16x16"]["parity_spot_check"] = sat["config"]["parity_spot_check"]
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
