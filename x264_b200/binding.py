"""ctypes binding of include/x264_b200.h.  Plumbing only -- all compute is in the CUDA library."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "x264_b200.h")

PIXEL_NAMES = ["16x16", "16x8", "8x16", "8x8", "8x4", "4x8", "4x4", "4x16"]
PIXEL_W = [16, 16, 8, 8, 8, 4, 4, 4]
PIXEL_H = [16, 8, 16, 8, 4, 8, 4, 16]
SAD, SSD, SATD, SA8D = 0, 1, 2, 3
PAD = 32

cand_dtype = np.dtype([("fenc_off", np.uint32), ("ref_off", np.uint32)])
cand_x4_dtype = np.dtype([("fenc_off", np.uint32), ("ref_off", np.uint32, (4,))])


class X264CUError(RuntimeError):
    pass


class Planes(C.Structure):
    _fields_ = [("d_origin", C.c_void_p), ("stride", C.c_ssize_t), ("plane_pitch", C.c_ssize_t),
                ("width", C.c_int), ("height", C.c_int), ("n_planes", C.c_int)]


_lib = None


def lib_path():
    from . import build as _b
    return _b.LIB


def lib():
    """Load (building first if the sources are newer) the CUDA library.  Raises if it cannot be built/loaded:
    there is deliberately no other implementation to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _b
    path = _b.LIB
    if os.environ.get("X264CU_LIB"):          # tuning hook: a variant built by `build.py --variant` (tools/)
        path = os.environ["X264CU_LIB"]
        if not os.path.exists(path):
            raise X264CUError("X264CU_LIB=%s does not exist" % path)
    elif _b.needs_build():
        if os.path.exists(_b.NVCC):
            path = _b.build()
        elif not os.path.exists(path):
            raise X264CUError("libx264_b200.so is missing and nvcc is not available to build it")
    L = C.CDLL(path)
    vp, ci, ss, sz = C.c_void_p, C.c_int, C.c_ssize_t, C.c_size_t
    L.x264cu_open.argtypes = [C.POINTER(vp), ci]
    L.x264cu_close.argtypes = [vp]
    L.x264cu_strerror.argtypes = [vp]
    L.x264cu_strerror.restype = C.c_char_p
    L.x264cu_device_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(sz)]
    L.x264cu_stream.argtypes = [vp]
    L.x264cu_stream.restype = vp
    L.x264cu_sync.argtypes = [vp]
    L.x264cu_timer_start.argtypes = [vp]
    L.x264cu_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.x264cu_launch_count.argtypes = [vp]
    L.x264cu_launch_count.restype = C.c_uint64
    L.x264cu_malloc.argtypes = [vp, sz]
    L.x264cu_malloc.restype = vp
    L.x264cu_free.argtypes = [vp, vp]
    L.x264cu_malloc_host.argtypes = [vp, sz]
    L.x264cu_malloc_host.restype = vp
    L.x264cu_free_host.argtypes = [vp, vp]
    L.x264cu_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.x264cu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.x264cu_memset.argtypes = [vp, vp, ci, sz]
    L.x264cu_pixel_cmp_batch.argtypes = [vp, ci, ci, vp, ss, vp, ss, vp, ci, vp]
    L.x264cu_pixel_cmp_x4_batch.argtypes = [vp, ci, ci, ci, vp, ss, vp, ss, vp, ci, vp]
    L.x264cu_pixel_cmp_batch_host.argtypes = [vp, ci, ci, vp, sz, ss, vp, sz, ss, vp, ci, vp]
    L.x264cu_pixel_cmp_mvfield.argtypes = [vp, ci, ci, C.POINTER(Planes), C.POINTER(Planes), ci, vp, vp]
    L.x264cu_pixel_cmp_mvfield_host.argtypes = [vp, ci, ci, vp, vp, ss, ss, ci, ci, ci, ci, vp, vp]
    _bind_optional(L)
    _lib = L
    return L


def _bind_optional(L):
    """entry points added by later translation units (frame preparation, motion search, lookahead)"""
    from . import binding_ext
    binding_ext.bind(L)


def header_symbols():
    """every function name declared in include/x264_b200.h"""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(x264cu_[a-z0-9_]+)\s*\(", txt)))


def exported_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path()], text=True)
    return sorted(l.split()[-1] for l in out.splitlines() if " T " in l)


def _addr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):          # torch tensor (device or host)
        return a.data_ptr()
    return int(a)


class Context:
    """x264cu_ctx_t wrapper.  Mirrors the life cycle of the reference's OpenCL context
    (x264_opencl_lookahead_init / _delete, common/opencl.c:411, :596)."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        if self.L.x264cu_open(C.byref(h), device) != 0:
            raise X264CUError(self.L.x264cu_strerror(None).decode())
        self.h = h
        self._bufs = []

    def close(self):
        if self.h:
            for p in self._bufs:
                self.L.x264cu_free(self.h, p)
            self._bufs = []
            self.L.x264cu_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc != 0:
            raise X264CUError(self.L.x264cu_strerror(self.h).decode())

    # -- memory ------------------------------------------------------------------------------------
    def malloc(self, nbytes):
        p = self.L.x264cu_malloc(self.h, nbytes)
        if not p:
            raise X264CUError(self.L.x264cu_strerror(self.h).decode())
        self._bufs.append(p)
        return p

    def free(self, p):
        self._bufs.remove(p)
        self.L.x264cu_free(self.h, p)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 16) + 256)
        self.check(self.L.x264cu_memcpy_h2d(self.h, p, arr.ctypes.data, arr.nbytes))
        return p

    def h2d(self, d, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.L.x264cu_memcpy_h2d(self.h, _addr(d), arr.ctypes.data, arr.nbytes))

    def download(self, p, shape, dtype):
        out = np.empty(shape, dtype)
        self.check(self.L.x264cu_memcpy_d2h(self.h, out.ctypes.data, _addr(p), out.nbytes))
        return out

    def sync(self):
        self.check(self.L.x264cu_sync(self.h))

    def timer_start(self):
        self.check(self.L.x264cu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.L.x264cu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def malloc_host(self, nbytes, dtype=np.uint8):
        """page-locked host buffer as a numpy array (kept alive until close)"""
        p = self.L.x264cu_malloc_host(self.h, nbytes)
        if not p:
            raise X264CUError(self.L.x264cu_strerror(self.h).decode())
        buf = (C.c_uint8 * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=np.uint8).view(dtype)
        self._host = getattr(self, "_host", []) + [p]
        return arr

    @property
    def stream(self):
        return self.L.x264cu_stream(self.h)

    @property
    def launches(self):
        return int(self.L.x264cu_launch_count(self.h))

    def device_info(self):
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self.check(self.L.x264cu_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hb)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    # -- B1: pixel table twins ---------------------------------------------------------------------
    def pixel_cmp_batch(self, metric, i_pixel, d_fenc, fenc_stride, d_ref, ref_stride, d_cand, n, d_out):
        self.check(self.L.x264cu_pixel_cmp_batch(self.h, metric, i_pixel, _addr(d_fenc), fenc_stride, _addr(d_ref),
                                                 ref_stride, _addr(d_cand), n, _addr(d_out)))

    def pixel_cmp_x4_batch(self, metric, i_pixel, n_refs, d_fenc, fenc_stride, d_ref, ref_stride, d_cand, n, d_out):
        self.check(self.L.x264cu_pixel_cmp_x4_batch(self.h, metric, i_pixel, n_refs, _addr(d_fenc), fenc_stride,
                                                    _addr(d_ref), ref_stride, _addr(d_cand), n, _addr(d_out)))

    def pixel_cmp_batch_host(self, metric, i_pixel, fenc, fenc_stride, ref, ref_stride, cand):
        """host numpy planes in, numpy costs out (the e2e form)"""
        cand = np.ascontiguousarray(cand, dtype=cand_dtype)
        out = np.empty(len(cand), np.int32)
        self.check(self.L.x264cu_pixel_cmp_batch_host(self.h, metric, i_pixel, fenc.ctypes.data, fenc.nbytes, fenc_stride,
                                                      ref.ctypes.data, ref.nbytes, ref_stride, cand.ctypes.data,
                                                      len(cand), out.ctypes.data))
        return out

    def pixel_cmp_mvfield(self, metric, i_pixel, fenc_planes, ref_planes, k_cands, d_mv, d_out):
        """fenc_planes / ref_planes: (d_origin, stride, plane_pitch, width, height, n_planes)"""
        pf, pr = Planes(*fenc_planes), Planes(*ref_planes)
        self.check(self.L.x264cu_pixel_cmp_mvfield(self.h, metric, i_pixel, C.byref(pf), C.byref(pr), k_cands,
                                                   _addr(d_mv), _addr(d_out)))

    def pixel_cmp_mvfield_host(self, metric, i_pixel, fenc_base, ref_base, stride, plane_pitch, width, height,
                               n_planes, k_cands, mv):
        mv = np.ascontiguousarray(mv, dtype=np.int16)
        n = k_cands * n_planes * (width // PIXEL_W[i_pixel]) * (height // PIXEL_H[i_pixel])
        assert mv.size == 2 * n
        out = np.empty(n, np.int32)
        self.check(self.L.x264cu_pixel_cmp_mvfield_host(self.h, metric, i_pixel, _addr(fenc_base), _addr(ref_base), stride,
                                                        plane_pitch, width, height, n_planes, k_cands, mv.ctypes.data,
                                                        out.ctypes.data))
        return out
