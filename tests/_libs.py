"""ctypes loaders for the test-side checkers: the plain-C oracle (oracle/liboracle.so) and the compiled,
unmodified reference (oracle/_ref/libx264ref.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libx264ref.so")

PIXEL_W = [16, 16, 8, 8, 8, 4, 4, 4]
PIXEL_H = [16, 8, 16, 8, 4, 8, 4, 16]
SAD, SSD, SATD, SA8D = 0, 1, 2, 3
PAD = 32

u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
cand_dtype = np.dtype([("fenc_off", np.uint32), ("ref_off", np.uint32)])
candp = np.ctypeslib.ndpointer(dtype=cand_dtype, flags="C_CONTIGUOUS")

_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle"))
                if f.startswith("oracle") and f.endswith((".c", ".h"))]
        if (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(ORACLE_SO)
        L.orc_pixel_cmp.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
        L.orc_pixel_cmp_batch.argtypes = [C.c_int, C.c_int, u8p, C.c_ssize_t, u8p, C.c_ssize_t, candp, C.c_int, i32p]
        L.orc_pixel_cmp_mvfield.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t,
                                            C.c_int, C.c_int, C.c_int, i16p, i32p]
        L.orc_cost_mv_table.argtypes = [u16p, C.c_int, C.c_int]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


REF_VEC_SO = os.path.join(ROOT, "oracle", "_ref", "libx264ref_vec.so")
_ref_vec = None


def ref_vec():
    """the reference's C path auto-vectorised for AVX2 (oracle/Makefile.ref, target vec): bench.py's second, labelled CPU row.
    None if it did not travel or this host cannot run AVX2 code."""
    global _ref_vec
    if _ref_vec is None and os.path.exists(REF_VEC_SO):
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            flags = ""
        if " avx2" in flags and " bmi2" in flags and " fma" in flags:
            L = C.CDLL(REF_VEC_SO)
            L.xref_open.restype = C.c_void_p
            L.xref_open.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
            L.xref_close.argtypes = [C.c_void_p]
            L.xref_pixel_cmp_batch.argtypes = [C.c_int, C.c_int, u8p, C.c_ssize_t, u8p, C.c_ssize_t, candp, C.c_int, i32p]
            L.xref_lookahead_types.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            _ref_vec = L
    return _ref_vec


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.xref_open.restype = C.c_void_p
        L.xref_open.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.xref_close.argtypes = [C.c_void_p]
        L.xref_param.argtypes = [C.c_void_p, C.c_char_p]
        L.xref_pixel_cmp.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
        L.xref_pixel_cmp_batch.argtypes = [C.c_int, C.c_int, u8p, C.c_ssize_t, u8p, C.c_ssize_t, candp, C.c_int, i32p]
        L.xref_cost_mv_table.argtypes = [C.c_void_p, u16p, C.c_int]
        _ref = L
    return _ref


def ptr(a, off=0):
    """address of numpy array element at flat byte offset `off`"""
    return C.c_void_p(a.ctypes.data + off)


class PaddedPlane:
    """u8 plane with a PAD-pixel border, x264 layout: data[(y+PAD)*stride + x+PAD]"""

    def __init__(self, width, height, stride=None, pad=PAD):
        self.w, self.h, self.pad = width, height, pad
        self.stride = stride or ((width + 2 * pad + 63) // 64 * 64)
        self.buf = np.zeros((height + 2 * pad) * self.stride, dtype=np.uint8)
        self.origin = pad * self.stride + pad

    def view(self):
        return self.buf.reshape(self.h + 2 * self.pad, self.stride)

    def inner(self):
        return self.view()[self.pad:self.pad + self.h, self.pad:self.pad + self.w]

    def fill_border(self):
        v = self.view()
        p, w, h = self.pad, self.w, self.h
        v[p:p + h, :p] = v[p:p + h, p:p + 1]
        v[p:p + h, p + w:p + w + p] = v[p:p + h, p + w - 1:p + w]
        v[:p, :w + 2 * p] = v[p:p + 1, :w + 2 * p]
        v[p + h:, :w + 2 * p] = v[p + h - 1:p + h, :w + 2 * p]
        return self

    def off(self, x, y):
        return self.origin + y * self.stride + x


def worst_case_pair(n, rng):
    """checkasm-style overflow patterns (tools/checkasm.c:381-394): maxed alternating differences"""
    a = np.zeros(n, np.uint8)
    b = np.zeros(n, np.uint8)
    pat = rng.integers(0, 4)
    idx = np.arange(n)
    if pat == 0:
        a[:] = 255
    elif pat == 1:
        b[:] = 255
    elif pat == 2:
        a[idx % 2 == 0] = 255
        b[idx % 2 == 1] = 255
    else:
        a[(idx // 4) % 2 == 0] = 255
        b[(idx // 4) % 2 == 1] = 255
    return a, b


def _bind_mc():
    o, r = oracle(), (ref() if have_ref() else None)
    vp, ss, ci = C.c_void_p, C.c_ssize_t, C.c_int
    o.orc_frame_init_lowres.argtypes = [vp, ss, ci, ci, C.POINTER(vp), ss, ci, ci]
    o.orc_hpel_filter_plane.argtypes = [vp, ss, ci, ci, vp, vp, vp, ss, ci]
    o.orc_mc_luma.argtypes = [vp, ss, C.POINTER(vp), ss, ci, ci, ci, ci, vp]
    o.orc_pixel_avg.argtypes = [vp, ss, vp, ss, vp, ss, ci, ci, ci]
    if r is not None:
        r.xref_mc_luma.argtypes = [vp, ss, vp, vp, vp, vp, ss] + [ci] * 8
        r.xref_get_ref.argtypes = [vp, ss, vp, vp, vp, vp, ss] + [ci] * 8
        r.xref_avg.argtypes = [ci, vp, ss, vp, ss, vp, ss, ci]
        r.xref_frame_lowres.argtypes = [vp, vp, ss, vp]
        r.xref_frame_hpel.argtypes = [vp, vp, ss, vp, vp]


class OrcWeight(C.Structure):
    _fields_ = [("enabled", C.c_int), ("scale", C.c_int), ("denom", C.c_int), ("offset", C.c_int)]


def synth_luma(width, height, seed, kind="texture"):
    """deterministic synthetic luma: smooth low-pass texture + noise, or pure noise"""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (height, width), dtype=np.uint8)
    base = rng.integers(0, 256, (height // 8 + 3, width // 8 + 3)).astype(np.float32)
    up = np.kron(base, np.ones((8, 8), np.float32))
    k = np.ones(9, np.float32) / 9
    up = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 0, up)
    up = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 1, up)
    img = up[4:4 + height, 4:4 + width] + rng.normal(0, 2.0, (height, width))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


class XrefMeArgs(C.Structure):
    _fields_ = [("i_pixel", C.c_int), ("me_method", C.c_int), ("subpel_refine", C.c_int), ("me_range", C.c_int),
                ("qp", C.c_int), ("mv_min_spel", C.c_int * 2), ("mv_max_spel", C.c_int * 2), ("mvp", C.c_int16 * 2),
                ("i_mvc", C.c_int), ("mvc", (C.c_int16 * 2) * 16),
                ("wt_en", C.c_int), ("wt_scale", C.c_int), ("wt_denom", C.c_int), ("wt_offset", C.c_int),
                ("use_thresh", C.c_int), ("halfpel_thresh", C.c_int),
                ("mv", C.c_int16 * 2), ("cost", C.c_int), ("cost_mv", C.c_int), ("thresh_out", C.c_int)]


class XrefChroma(C.Structure):
    _fields_ = [("fenc_uv", C.c_void_p), ("fenc_uv_stride", C.c_ssize_t), ("fref_uv", C.c_void_p), ("fref_uv_stride", C.c_ssize_t),
                ("wt", (C.c_int * 4) * 2)]


class OrcMeCtx(C.Structure):
    _fields_ = [("me_method", C.c_int), ("subpel_refine", C.c_int), ("me_range", C.c_int), ("mbcmp_is_satd", C.c_int),
                ("mv_min_spel", C.c_int * 2), ("mv_max_spel", C.c_int * 2), ("mv_limit_fpel", (C.c_int * 2) * 2),
                ("chroma_me", C.c_int)]


class OrcMe(C.Structure):
    _fields_ = [("i_pixel", C.c_int), ("p_cost_mv", C.c_void_p), ("p_fref", C.c_void_p * 4), ("p_fref_w", C.c_void_p),
                ("p_fenc", C.c_void_p), ("fenc_stride", C.c_ssize_t), ("stride", C.c_ssize_t), ("weight", OrcWeight),
                ("mvp", C.c_int16 * 2), ("cost_mv", C.c_int), ("cost", C.c_int), ("mv", C.c_int16 * 2),
                ("p_fref_uv", C.c_void_p), ("stride_uv", C.c_ssize_t), ("p_fenc_uv", C.c_void_p), ("fenc_uv_stride", C.c_ssize_t),
                ("weight_uv", OrcWeight * 2)]


def _bind_me():
    o = oracle()
    o.orc_me_search_ref.argtypes = [C.POINTER(OrcMeCtx), C.POINTER(OrcMe), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    if have_ref():
        r = ref()
        vp = C.c_void_p
        r.xref_me_search.argtypes = [vp, C.POINTER(XrefMeArgs), vp, C.c_ssize_t, vp, vp, vp, vp, vp, C.c_ssize_t]
        r.xref_me_search_chroma.argtypes = [vp, C.POINTER(XrefMeArgs), vp, C.c_ssize_t, vp, vp, vp, vp, vp, C.c_ssize_t, C.POINTER(XrefChroma)]
        r.xref_cost_mv_table_qp.argtypes = [vp, C.c_int, u16p, C.c_int]


def make_ref_planes(luma, stride=None):
    """F plane + H,V,C half-pel planes (all padded, border filled) of a luma picture, via the oracle"""
    _bind_mc()
    h, w = luma.shape
    F = PaddedPlane(w, h, stride=stride)
    F.inner()[:] = luma
    F.fill_border()
    H, V, Cc = PaddedPlane(w, h, stride=F.stride), PaddedPlane(w, h, stride=F.stride), PaddedPlane(w, h, stride=F.stride)
    src = np.ascontiguousarray(luma)
    oracle().orc_hpel_filter_plane(ptr(src), w, w, h, ptr(H.buf, H.origin), ptr(V.buf, V.origin), ptr(Cc.buf, Cc.origin),
                                   F.stride, PAD)
    return [F, H, V, Cc]


class OrcLaParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "mb_width", "mb_height", "subpel_refine", "me_method", "me_range",
                                       "mv_range", "bframes", "bframe_bias", "weighted_bipred", "aq_mode", "do_edges", "vbv", "weighted_pred")]


def _bind_la():
    o = oracle()
    vp, ci = C.c_void_p, C.c_int
    o.orc_la_frame_new.restype = vp
    o.orc_la_frame_new.argtypes = [C.POINTER(OrcLaParams), vp, C.c_ssize_t]
    o.orc_la_frame_delete.argtypes = [vp]
    o.orc_la_frame_cost.argtypes = [C.POINTER(OrcLaParams), vp, C.POINTER(vp), ci, ci, ci]
    o.orc_la_frame_get.argtypes = [vp, ci, ci, ci, vp]
    o.orc_la_frame_set_qscale.argtypes = [vp, u16p]
    o.orc_la_frame_plane.restype = vp
    o.orc_la_frame_plane.argtypes = [vp, ci]
    o.orc_la_frame_stride.argtypes = [vp]
    o.orc_la_mbtree_propagate.argtypes = [C.POINTER(OrcLaParams), C.POINTER(vp), ci, ci, ci, ci, C.c_float]
    o.orc_la_mbtree_finish.argtypes = [vp, ci, ci, C.c_float]
    o.orc_la_mbtree_reset.argtypes = [vp]
    o.orc_la_frame_set_qp_offset_aq.argtypes = [vp, vp]
    o.orc_la_frame_get_mbtree.argtypes = [vp, ci, ci, vp]
    o.orc_log2.argtypes = [C.c_uint32]
    o.orc_log2.restype = C.c_float
    if have_ref():
        r = ref()
        r.xref_la_new.restype = vp
        r.xref_la_new.argtypes = [vp, ci]
        r.xref_la_set_frame.argtypes = [vp, ci, vp, C.c_ssize_t, vp]
        r.xref_la_frame_cost.argtypes = [vp, ci, ci, ci]
        r.xref_la_get.argtypes = [vp, ci, ci, ci, ci, vp]
        r.xref_la_get_lowres.argtypes = [vp, ci, ci, vp]
        r.xref_la_free.argtypes = [vp]
        r.xref_la_set_type.argtypes = [vp, ci, ci, C.c_float]
        r.xref_la_mbtree_reset.argtypes = [vp, ci]
        r.xref_la_mbtree_propagate.argtypes = [vp, C.c_float, ci, ci, ci, ci]
        r.xref_la_mbtree_finish.argtypes = [vp, ci, C.c_float, ci]
        r.xref_la_mbtree.argtypes = [vp, ci, ci]
        r.xref_la_get_mbtree.argtypes = [vp, ci, ci, ci, vp]
        r.xref_la_set_qp_offset_aq.argtypes = [vp, ci, vp]


def la_params_from_ref(hnd, width, height):
    """orc_la_params_t mirroring an opened reference encoder (x264_t fields)"""
    r = ref()
    g = lambda n: r.xref_param(hnd, n.encode())
    p = OrcLaParams()
    p.width, p.height = width, height
    p.mb_width, p.mb_height = g("mb_width"), g("mb_height")
    p.subpel_refine, p.me_method, p.me_range, p.mv_range = g("subme"), g("me"), g("merange"), g("mvrange")
    p.bframes, p.bframe_bias, p.weighted_bipred = g("bframes"), g("b_bias"), g("weightb")
    p.aq_mode = int(g("aq_mode") != 0)
    p.vbv = int(g("vbv") != 0)
    p.do_edges = int(g("mbtree") != 0 or g("vbv") != 0)
    p.weighted_pred = -1 if g("weightp") < 0 else int(g("weightp") != 0)      # X264_WEIGHTP_FAKE = -1 (encoder.c:1316-1317)
    return p


def synth_sequence(width, height, n, seed, cut_at=None):
    """moving low-pass texture with per-frame global motion (integer + half-pel), optional scene cut"""
    rng = np.random.default_rng(seed)
    master = synth_luma(2 * width + 128, 2 * height + 128, seed).astype(np.float32)
    master2 = synth_luma(2 * width + 128, 2 * height + 128, seed + 1).astype(np.float32)
    frames = []
    x = y = 32.0
    for i in range(n):
        if cut_at is not None and i == cut_at:
            master = master2
        x += rng.integers(-6, 7) / 1.0
        y += rng.integers(-4, 5) / 1.0
        x = float(np.clip(x, 0, 120))
        y = float(np.clip(y, 0, 120))
        xi, yi = int(x), int(y)
        crop = master[yi:yi + 2 * height:2, xi:xi + 2 * width:2]
        img = crop + rng.normal(0, 1.5, crop.shape)
        frames.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return frames


_st_oracle = None


def slicetype_oracle_lib():
    """the product's host slice-type logic (x264_b200/csrc/slicetype.c) linked against the CPU oracle instead of the GPU
    lookahead, so that its control flow can be checked against the reference's encoder without a device"""
    global _st_oracle
    if _st_oracle is None:
        oracle()
        out_dir = os.path.join(ROOT, "tests", "_build")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libslicetype_oracle.so")
        srcs = [os.path.join(ROOT, "x264_b200", "csrc", "slicetype.c"), os.path.join(ROOT, "tests", "csrc", "slicetype_oracle_glue.c")]
        deps = srcs + [ORACLE_SO, os.path.join(ROOT, "include", "x264_b200.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu99", "-o", so] + srcs +
                                  ["-L" + os.path.dirname(ORACLE_SO), "-loracle", "-Wl,-rpath," + os.path.dirname(ORACLE_SO), "-lm"])
        L = C.CDLL(so)
        from x264_b200.binding_ext import bind
        _st_oracle = L
    return _st_oracle


_st_ref = None


def slicetype_ref_lib():
    """host slice-type logic linked against the compiled reference's own slicetype_frame_cost (bench.py's CPU arm)"""
    global _st_ref
    if _st_ref is None:
        out_dir = os.path.join(ROOT, "tests", "_build")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libslicetype_ref.so")
        srcs = [os.path.join(ROOT, "x264_b200", "csrc", "slicetype.c"), os.path.join(ROOT, "tests", "csrc", "slicetype_ref_glue.c")]
        deps = srcs + [REF_SO, os.path.join(ROOT, "include", "x264_b200.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu99", "-o", so] + srcs +
                                  ["-L" + os.path.dirname(REF_SO), "-lx264ref", "-Wl,-rpath," + os.path.dirname(REF_SO), "-lm"])
        _st_ref = C.CDLL(so)
        _st_ref.slicetype_ref_glue_config.argtypes = [C.c_char_p, C.c_char_p]
    return _st_ref
