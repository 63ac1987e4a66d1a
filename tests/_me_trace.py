"""BASELINE config 3 harness: record every x264_me_t the reference ENCODER hands to x264_me_search_ref while it
codes a few pictures (oracle/ref_shim.c: xref_me_trace_*; the reference itself is unmodified -- the recorder sits behind the
macro of encoder/me.h), then replay the whole stream through the oracle (CPU) or through x264cu_me_search_frame (GPU) and
compare (mv, cost, cost_mv, half-pel threshold) search by search.  Test infrastructure: used by tests/ and by bench.py's
cpu_baseline / parity legs only."""
import ctypes as C

import numpy as np

import _libs
from _libs import oracle, ref, OrcMeCtx, OrcMe, OrcWeight

PADH = PADV = 32
REC = np.dtype([("i_pixel", "i4"), ("bx", "i4"), ("by", "i4"), ("ref_idx", "i4"), ("qp", "i4"), ("lambda", "i4"), ("i_mvc", "i4"),
                ("thresh_in", "i4"), ("mvp", "i2", (2,)), ("mvc", "i2", (9, 2)), ("lim", "i2", (4,)), ("mv", "i2", (2,)),
                ("cost", "i4"), ("cost_mv", "i4"), ("thresh_out", "i4")])
assert REC.itemsize == 96
INFO = ("coded", "display", "slice_type", "chroma_me", "me_method", "subpel", "me_range", "mbcmp_satd", "mv_range", "fpel_border",
        "width", "height", "stride", "lines", "stride_uv", "lines_uv", "n_refs", "n_recs")


def synth_i420(width, height, n, seed):
    """n I420 pictures (packed Y, Cb, Cr): a low-pass texture under global motion with a few independently moving blocks, film
    grain, and chroma that follows the luma motion -- so that multi-reference, sub-partition and chroma ME all matter"""
    rng = np.random.default_rng(seed)
    big_y = _libs.synth_luma(width + 256, height + 256, seed).astype(np.float32)
    big_u = _libs.synth_luma(width // 2 + 128, height // 2 + 128, seed + 1).astype(np.float32)
    big_v = _libs.synth_luma(width // 2 + 128, height // 2 + 128, seed + 2).astype(np.float32)
    sprites = [(int(rng.integers(0, max(1, width - 96))) & ~1, int(rng.integers(0, max(1, height - 96))) & ~1,
                int(rng.integers(-6, 7)) * 2, int(rng.integers(-4, 5)) * 2, int(rng.integers(1 << 30))) for _ in range(6)]
    out = np.empty((n, width * height * 3 // 2), np.uint8)
    x = y = 64
    for i in range(n):
        x = int(np.clip(x + 2 * rng.integers(-3, 4), 0, 250)) & ~1
        y = int(np.clip(y + 2 * rng.integers(-2, 3), 0, 250)) & ~1
        Y = big_y[y:y + height, x:x + width].copy()
        U = big_u[y // 2:y // 2 + height // 2, x // 2:x // 2 + width // 2].copy()
        V = big_v[y // 2:y // 2 + height // 2, x // 2:x // 2 + width // 2].copy()
        for k, (sx, sy, dx, dy, sd) in enumerate(sprites):
            px, py = (sx + i * dx) % max(2, width - 96) & ~1, (sy + i * dy) % max(2, height - 96) & ~1
            tex = _libs.synth_luma(96, 96, sd).astype(np.float32)
            Y[py:py + 96, px:px + 96] = tex[:Y[py:py + 96, px:px + 96].shape[0], :Y[py:py + 96, px:px + 96].shape[1]]
            U[py // 2:py // 2 + 48, px // 2:px // 2 + 48] = 90 + 10 * k
            V[py // 2:py // 2 + 48, px // 2:px // 2 + 48] = 160 - 10 * k
        Y = Y * (1.0 - 0.02 * (i % 3)) + rng.normal(0, 1.5, Y.shape)          # a little brightness drift: weighted prediction
        f = out[i]
        f[:width * height] = np.clip(np.rint(Y), 0, 255).astype(np.uint8).reshape(-1)
        f[width * height:width * height * 5 // 4] = np.clip(np.rint(U), 0, 255).astype(np.uint8).reshape(-1)
        f[width * height * 5 // 4:] = np.clip(np.rint(V), 0, 255).astype(np.uint8).reshape(-1)
    return out


class TraceFrame:
    pass


def record(width, height, n_frames, opts, max_frames, skip=1, preset=b"slower", seed=2160, yuv=None):
    """encode n_frames synthetic pictures with the compiled reference and return the recorded searches of up to max_frames
    coded pictures (starting with coded picture `skip`: 0 is the IDR picture, which has none)"""
    r = ref()
    vp, ci = C.c_void_p, C.c_int
    r.xref_me_trace_start.argtypes = [ci, ci]
    r.xref_me_trace_frame_info.argtypes = [ci, C.POINTER(ci)]
    r.xref_me_trace_recs.restype = vp
    r.xref_me_trace_recs.argtypes = [ci]
    r.xref_me_trace_plane.restype = vp
    r.xref_me_trace_plane.argtypes = [ci, ci, ci]
    r.xref_me_trace_ref_info.argtypes = [ci, ci, C.POINTER(ci)]
    r.xref_encode_i420.argtypes = [vp, vp, ci]
    if yuv is None:
        yuv = synth_i420(width, height, n_frames, seed)
    hnd = r.xref_open(width, height, preset, opts, 0)
    assert hnd, "xref_open failed for %r" % (opts,)
    frames = []
    try:
        assert r.xref_me_trace_start(max_frames, skip) == 0
        n_out = r.xref_encode_i420(hnd, yuv.ctypes.data, n_frames)
        assert n_out == n_frames, n_out
        r.xref_me_trace_stop()
        for i in range(r.xref_me_trace_frames()):
            info = (ci * 18)()
            assert r.xref_me_trace_frame_info(i, info) == 0
            t = TraceFrame()
            for k, name in enumerate(INFO):
                setattr(t, name, int(info[k]))

            def grab(ptr_, nbytes):
                return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr_)).copy()
            t.fenc = grab(r.xref_me_trace_plane(i, 0, -1), t.stride * t.lines)
            t.fenc_uv = grab(r.xref_me_trace_plane(i, 0, -2), t.stride_uv * t.lines_uv)
            t.refs = []
            for k in range(t.n_refs):
                ri = (ci * 16)()
                assert r.xref_me_trace_ref_info(i, k, ri) == 0
                d = dict(list=int(ri[0]), i_ref=int(ri[1]), display=int(ri[2]), weighted=int(ri[3]),
                         weight=[[int(ri[4 + 4 * a + b]) for b in range(4)] for a in range(3)])
                nb = t.stride * (t.lines + 2 * PADV)
                d["planes"] = [grab(r.xref_me_trace_plane(i, k, w), nb) for w in range(4)]
                d["wplane"] = grab(r.xref_me_trace_plane(i, k, 4), nb) if d["weighted"] else None
                d["uv"] = grab(r.xref_me_trace_plane(i, k, 5), t.stride_uv * (t.lines_uv + PADV))
                t.refs.append(d)
            p = r.xref_me_trace_recs(i)
            t.recs = np.frombuffer((C.c_uint8 * (REC.itemsize * t.n_recs)).from_address(p), REC).copy() if t.n_recs else np.zeros(0, REC)
            frames.append(t)
    finally:
        r.xref_me_trace_free()
        r.xref_close(hnd)
    return frames


def luma_origin(t):
    return PADV * t.stride + PADH


def chroma_origin(t):
    return (PADV // 2) * t.stride_uv + PADH


def replay_oracle(t, sel=None):
    """-> int32 [n, 5] = (mvx, mvy, cost, cost_mv or -1 after an early exit, threshold out) from the oracle's search"""
    _libs._bind_me()
    o = oracle()
    recs = t.recs if sel is None else t.recs[sel]
    out = np.zeros((len(recs), 5), np.int32)
    n = 2 * 4 * t.mv_range
    tabs = {}
    for k, rc in enumerate(recs):
        lam = int(rc["lambda"])
        if lam not in tabs:
            tab = np.zeros(2 * n + 1, np.uint16)
            o.orc_cost_mv_table(tab, n, lam)
            tabs[lam] = tab
        rf = t.refs[int(rc["ref_idx"])]
        c = OrcMeCtx()
        c.me_method, c.subpel_refine, c.me_range, c.mbcmp_is_satd, c.chroma_me = t.me_method, t.subpel, t.me_range, t.mbcmp_satd, t.chroma_me
        for i in range(2):
            c.mv_min_spel[i], c.mv_max_spel[i] = int(rc["lim"][i]), int(rc["lim"][2 + i])
            c.mv_limit_fpel[0][i] = (int(rc["lim"][i]) >> 2) + t.fpel_border
            c.mv_limit_fpel[1][i] = (int(rc["lim"][2 + i]) >> 2) - t.fpel_border
        bx, by = int(rc["bx"]), int(rc["by"])
        off = luma_origin(t) + by * t.stride + bx
        m = OrcMe()
        m.i_pixel = int(rc["i_pixel"])
        m.p_cost_mv = tabs[lam].ctypes.data + 2 * n
        for i in range(4):
            m.p_fref[i] = rf["planes"][i].ctypes.data + off
        m.p_fref_w = (rf["wplane"] if rf["weighted"] else rf["planes"][0]).ctypes.data + off
        m.p_fenc = t.fenc.ctypes.data + by * t.stride + bx
        m.fenc_stride, m.stride = t.stride, t.stride
        m.weight = OrcWeight(*rf["weight"][0])
        m.mvp[0], m.mvp[1] = int(rc["mvp"][0]), int(rc["mvp"][1])
        m.p_fref_uv = rf["uv"].ctypes.data + chroma_origin(t) + (by >> 1) * t.stride_uv + (bx & ~1)
        m.stride_uv = t.stride_uv
        m.p_fenc_uv = t.fenc_uv.ctypes.data + (by >> 1) * t.stride_uv + (bx & ~1)
        m.fenc_uv_stride = t.stride_uv
        m.weight_uv[0], m.weight_uv[1] = OrcWeight(*rf["weight"][1]), OrcWeight(*rf["weight"][2])
        mvc = np.ascontiguousarray(rc["mvc"])
        use = int(rc["thresh_in"]) >= 0
        th = C.c_int(int(rc["thresh_in"]))
        m.cost_mv = -1
        o.orc_me_search_ref(C.byref(c), C.byref(m), mvc.ctypes.data, int(rc["i_mvc"]), C.byref(th) if use else None)
        out[k] = (m.mv[0], m.mv[1], m.cost, m.cost_mv, th.value if use else -1)
    return out


def build_jobs(t, x):
    """the recorded searches as x264cu_me_frame_job_t (offsets from pixel (0,0)) and the list of distinct lambdas"""
    recs = t.recs
    lambdas = sorted(set(int(v) for v in np.unique(recs["lambda"])))
    lut = {v: i for i, v in enumerate(lambdas)}
    jobs = np.zeros(len(recs), x.me_frame_job_dtype)
    j = jobs["job"]
    j["i_pixel"] = recs["i_pixel"]
    j["fenc_off"] = recs["by"] * t.stride + recs["bx"]
    j["ref_off"] = recs["by"] * t.stride + recs["bx"]
    j["mvp"] = recs["mvp"]
    j["mvc"] = recs["mvc"]
    j["i_mvc"] = recs["i_mvc"]
    j["mv_min_spel"] = recs["lim"][:, :2]
    j["mv_max_spel"] = recs["lim"][:, 2:]
    j["halfpel_thresh"] = recs["thresh_in"]
    jobs["i_ref"] = recs["ref_idx"]
    jobs["i_lambda"] = np.vectorize(lut.get)(recs["lambda"]) if len(recs) else 0
    return jobs, lambdas


class DeviceTrace:
    """the planes of one traced picture in HBM + its job list, ready for x264cu_me_search_frame"""

    def __init__(self, ctx, t, x):
        self.ctx, self.t, self.x = ctx, t, x
        self.live = []

        self.pairs = []                       # (device buffer, host array): what reupload() copies again

        def up(a, origin=0):
            d = ctx.upload(a)
            self.live.append(d)
            self.pairs.append((d, a))
            return d + origin
        lo, co = luma_origin(t), chroma_origin(t)
        d_fenc, d_fenc_uv = up(t.fenc), up(t.fenc_uv)
        refs = []
        for rf in t.refs:
            refs.append(([up(p, lo) for p in rf["planes"]], up(rf["wplane"], lo) if rf["weighted"] else None, up(rf["uv"], co), rf["weight"]))
        self.jobs, self.lambdas = build_jobs(t, x)
        self.frame, self.keep = x.make_me_frame(d_fenc, t.stride, d_fenc_uv, t.stride_uv, t.stride, t.stride_uv, refs, self.lambdas, t.chroma_me)
        self.params = x.MeParams(t.me_method, t.subpel, t.me_range, t.mbcmp_satd, 1, t.mv_range, 0, 0, 0, 0, t.fpel_border)
        self.n = len(self.jobs)
        self.d_jobs = ctx.upload(self.jobs)
        self.pairs.append((self.d_jobs, self.jobs))
        self.d_res = ctx.malloc(max(self.n, 1) * x.me_result_dtype.itemsize)
        self.live += [self.d_jobs, self.d_res]

    def reupload(self):
        """host -> HBM copy of every plane and of the job records again (bench.py's end-to-end leg)"""
        for d, a in self.pairs:
            self.ctx.h2d(d, a)

    def launch(self):
        self.ctx.check(self.ctx.L.x264cu_me_search_frame(self.ctx.h, C.byref(self.params), C.byref(self.frame), self.d_jobs, self.n, self.d_res))

    def results(self):
        r = self.ctx.download(self.d_res, (self.n,), self.x.me_result_dtype)
        out = np.zeros((self.n, 5), np.int32)
        out[:, 0], out[:, 1], out[:, 2], out[:, 3], out[:, 4] = r["mv"][:, 0], r["mv"][:, 1], r["cost"], r["cost_mv"], r["halfpel_thresh"]
        return out

    def close(self):
        for d in self.live:
            self.ctx.free(d)
        self.live = []


def expected(t):
    """what the reference got: (mvx, mvy, cost, cost_mv, threshold out); cost_mv is undefined after a half-pel early exit"""
    r = t.recs
    out = np.zeros((len(r), 5), np.int32)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3], out[:, 4] = r["mv"][:, 0], r["mv"][:, 1], r["cost"], r["cost_mv"], r["thresh_out"]
    return out


def early_exit(t):
    """searches that left refine_subpel at the half-pel threshold (me.c:934-943): cost_mv is not written"""
    r = t.recs
    return (r["thresh_in"] >= 0) & ((r["cost"].astype(np.int64) * 7 >> 3) > r["thresh_in"])


class ReplayRef(C.Structure):
    _fields_ = [("planes", C.c_void_p * 4), ("wplane", C.c_void_p), ("uv", C.c_void_p), ("weight", (C.c_int * 4) * 3)]


def replay_reference(t, opts, n_threads=1, preset=b"slower", sel=None):
    """the reference's own x264_me_search_ref over the recorded searches of picture t, spread over n_threads host threads
    (one encoder handle each: x264_t is per-thread state) -> (results int32 [n, 5], seconds of wall time)"""
    import threading
    import time
    r = ref()
    vp = C.c_void_p
    r.xref_me_replay.argtypes = [vp, C.POINTER(C.c_int), vp, vp, C.POINTER(ReplayRef), vp, C.c_int, vp]
    recs = np.ascontiguousarray(t.recs if sel is None else t.recs[sel])
    n = len(recs)
    info = (C.c_int * 18)(*[int(getattr(t, k)) for k in INFO])
    refs = (ReplayRef * len(t.refs))()
    for i, rf in enumerate(t.refs):
        for k in range(4):
            refs[i].planes[k] = rf["planes"][k].ctypes.data
        refs[i].wplane = rf["wplane"].ctypes.data if rf["weighted"] else None
        refs[i].uv = rf["uv"].ctypes.data
        for a in range(3):
            for b in range(4):
                refs[i].weight[a][b] = rf["weight"][a][b]
    out = np.zeros((n, 5), np.int32)
    n_threads = max(1, min(n_threads, n))
    handles = [r.xref_open(t.width, t.height, preset, opts, 0) for _ in range(n_threads)]
    assert all(handles)
    bounds = [n * i // n_threads for i in range(n_threads + 1)]

    def work(k):
        lo, hi = bounds[k], bounds[k + 1]
        r.xref_me_replay(handles[k], info, t.fenc.ctypes.data, t.fenc_uv.ctypes.data, refs,
                         recs.ctypes.data + lo * REC.itemsize, hi - lo, out.ctypes.data + lo * 20)
    try:
        th = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
        t0 = time.perf_counter()
        for x_ in th:
            x_.start()
        for x_ in th:
            x_.join()
        dt = time.perf_counter() - t0
    finally:
        for hnd in handles:
            r.xref_close(hnd)
    return out, dt
