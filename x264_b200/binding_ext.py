"""argtypes for the entry points beyond the pixel table (frame preparation, lookahead)"""
import ctypes as C

import numpy as np


class LookaheadParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "subpel_refine", "me_method", "me_range", "mv_range", "bframes",
                                       "bframe_bias", "weighted_bipred", "aq_mode", "mb_tree", "vbv", "n_slots", "weighted_pred")]


class MeJob(C.Structure):
    _fields_ = [("i_pixel", C.c_int32), ("fenc_off", C.c_uint32), ("ref_off", C.c_uint32), ("mvp", C.c_int16 * 2),
                ("mvc", (C.c_int16 * 2) * 9), ("i_mvc", C.c_int32), ("mv_min_spel", C.c_int16 * 2), ("mv_max_spel", C.c_int16 * 2),
                ("halfpel_thresh", C.c_int32)]


class MeResult(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("cost", C.c_int32), ("cost_mv", C.c_int32), ("halfpel_thresh", C.c_int32)]


class MeParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("me_method", "subpel_refine", "me_range", "mbcmp_satd", "lambda_", "mv_range",
                                       "weight_enabled", "weight_scale", "weight_denom", "weight_offset", "fpel_border")]


me_job_dtype = np.dtype([("i_pixel", np.int32), ("fenc_off", np.uint32), ("ref_off", np.uint32), ("mvp", np.int16, (2,)),
                         ("mvc", np.int16, (9, 2)), ("i_mvc", np.int32), ("mv_min_spel", np.int16, (2,)),
                         ("mv_max_spel", np.int16, (2,)), ("halfpel_thresh", np.int32)])
me_frame_job_dtype = np.dtype([("job", me_job_dtype), ("i_ref", np.int16), ("i_lambda", np.int16)])


class MeRef(C.Structure):
    _fields_ = [("d_fref", C.c_void_p * 4), ("d_fref_w", C.c_void_p), ("d_fref_uv", C.c_void_p), ("weight", (C.c_int * 4) * 3)]


class MeFrame(C.Structure):
    _fields_ = [("d_fenc", C.c_void_p), ("fenc_stride", C.c_ssize_t), ("d_fenc_uv", C.c_void_p), ("fenc_uv_stride", C.c_ssize_t),
                ("ref_stride", C.c_ssize_t), ("ref_uv_stride", C.c_ssize_t), ("refs", C.POINTER(MeRef)), ("n_refs", C.c_int),
                ("lambdas", C.POINTER(C.c_int)), ("n_lambdas", C.c_int), ("chroma_me", C.c_int)]


me_result_dtype = np.dtype([("mv", np.int16, (2,)), ("cost", np.int32), ("cost_mv", np.int32), ("halfpel_thresh", np.int32)])
bidir_job_dtype = np.dtype([("i_pixel", np.int32), ("fenc_off", np.uint32), ("ref0_off", np.uint32), ("ref1_off", np.uint32),
                            ("mv", np.int16, (4,)), ("mvp", np.int16, (4,)), ("mv_min_spel", np.int16, (2,)),
                            ("mv_max_spel", np.int16, (2,)), ("i_weight", np.int32)])
bidir_result_dtype = np.dtype([("mv", np.int16, (4,)), ("cost", np.int32)])
me_refine_job_dtype = np.dtype([("i_pixel", np.int32), ("fenc_off", np.uint32), ("ref_off", np.uint32), ("mvp", np.int16, (2,)),
                                ("mv", np.int16, (2,)), ("cost", np.int32), ("i_ref_cost", np.int32), ("mv_min_spel", np.int16, (2,)),
                                ("mv_max_spel", np.int16, (2,)), ("halfpel_thresh", np.int32)])


class SlicetypeParams(C.Structure):
    _fields_ = [("la", LookaheadParams)] + [(n, C.c_int) for n in ("keyint_max", "keyint_min", "scenecut_threshold", "b_adapt",
                                                                  "b_pyramid", "rc_lookahead", "psy", "frame_reference", "rc_cqp",
                                                                  "fps_num", "fps_den")] + [("qcompress", C.c_float), ("aq_strength", C.c_float),
                                                                                                  ("open_gop", C.c_int), ("intra_refresh", C.c_int)]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        # negative = the reference's defaults (0.6 / 1.0); zero is a legal value of both
        if len(a) <= 12 and "qcompress" not in kw:
            self.qcompress = -1.0
        if len(a) <= 13 and "aq_strength" not in kw:
            self.aq_strength = -1.0


TYPE_NAMES = {0: "AUTO", 1: "IDR", 2: "I", 3: "P", 4: "BREF", 5: "B"}


def bind(L):
    vp, ci, ss = C.c_void_p, C.c_int, C.c_ssize_t
    L.x264cu_me_search_batch.argtypes = [vp, C.POINTER(MeParams), vp, ss, C.POINTER(vp), vp, ss, vp, ci, vp]
    L.x264cu_me_search_frame.argtypes = [vp, C.POINTER(MeParams), C.POINTER(MeFrame), vp, ci, vp]
    L.x264cu_me_refine_bidir_batch.argtypes = [vp, C.POINTER(MeParams), vp, ss, C.POINTER(vp), C.POINTER(vp), ss, vp, ci, vp]
    L.x264cu_me_refine_qpel_batch.argtypes = [vp, C.POINTER(MeParams), ci, vp, ss, C.POINTER(vp), ss, vp, ci, vp]
    L.x264cu_slicetype_open.argtypes = [vp, C.POINTER(SlicetypeParams), C.POINTER(vp)]
    L.x264cu_slicetype_close.argtypes = [vp]
    L.x264cu_slicetype_step.argtypes = [vp, vp, ss, vp, C.POINTER(ci), C.POINTER(ci)]
    L.x264cu_slicetype_step_device.argtypes = [vp, vp, ss, vp, C.POINTER(ci), C.POINTER(ci)]
    L.x264cu_slicetype_set_prefetch.argtypes = [vp, ci]
    L.x264cu_slicetype_set_run_ahead.argtypes = [vp, ci]
    L.x264cu_slicetype_set_speculation.argtypes = [vp, ci]
    L.x264cu_slicetype_set_prefetch_group.argtypes = [vp, ci]
    L.x264cu_lookahead_speculation_stats.restype = C.c_long
    L.x264cu_lookahead_speculation_stats.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.x264cu_lookahead_finalize_batch.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.x264cu_slicetype_set_async_upload.argtypes = [vp, ci]
    L.x264cu_slicetype_get_qp_offset.argtypes = [vp, ci, vp]
    L.x264cu_slicetype_set_shard.argtypes = [vp, ci, ci, vp, vp]
    L.x264cu_slicetype_set_next_type.argtypes = [vp, ci]
    L.x264cu_slicetype_rc_analyse_slice.argtypes = [vp, ci, vp, vp, vp]
    L.x264cu_slicetype_get_planned.argtypes = [vp, ci, vp, vp, ci]
    L.x264cu_slicetype_step_i420.argtypes = [vp, vp, ss, vp, vp, ss, C.POINTER(ci), C.POINTER(ci)]
    L.x264cu_slicetype_lookahead.argtypes = [vp]
    L.x264cu_slicetype_lookahead.restype = vp
    L.x264cu_lookahead_search_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.x264cu_lookahead_frame_set_qp_offset_aq.argtypes = [vp, ci, vp]
    L.x264cu_lookahead_mbtree_reset.argtypes = [vp, ci]
    L.x264cu_lookahead_mbtree_swap.argtypes = [vp, ci, ci]
    L.x264cu_lookahead_mbtree_propagate.argtypes = [vp, C.POINTER(ci), ci, ci, ci, ci, C.c_float]
    L.x264cu_lookahead_mbtree_finish.argtypes = [vp, ci, ci, ci, C.c_float]
    L.x264cu_lookahead_frame_cost_recalculate.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    L.x264cu_lookahead_get_qp_offset.argtypes = [vp, ci, vp]
    L.x264cu_lookahead_get_propagate_cost.argtypes = [vp, ci, vp]
    L.x264cu_lookahead_get_weighted_cost_delta.argtypes = [vp, ci, ci]
    L.x264cu_lookahead_get_weighted_cost_delta.restype = C.c_float
    L.x264cu_lookahead_set_async_upload.argtypes = [vp, ci]
    L.x264cu_slicetype_lookahead.argtypes = [vp]
    L.x264cu_slicetype_lookahead.restype = vp
    L.x264cu_slicetype_slot_of.argtypes = [vp, ci]
    L.x264cu_slicetype_cost_requests.argtypes = [vp]
    L.x264cu_slicetype_cost_requests.restype = C.c_long
    L.x264cu_frame_init_lowres.argtypes = [vp, vp, ss, ci, ci, C.POINTER(vp), ss]
    L.x264cu_hpel_filter.argtypes = [vp, vp, ss, ci, ci, vp, vp, vp, ci]
    L.x264cu_lookahead_open.argtypes = [vp, C.POINTER(LookaheadParams), C.POINTER(vp)]
    L.x264cu_lookahead_close.argtypes = [vp]
    L.x264cu_lookahead_frame_put.argtypes = [vp, ci, vp, ss, vp]
    L.x264cu_lookahead_frame_put_device.argtypes = [vp, ci, vp, ss, vp]
    L.x264cu_lookahead_frame_cost.argtypes = [vp, C.POINTER(ci), ci, ci, ci, C.POINTER(ci)]
    L.x264cu_lookahead_search_batch.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    L.x264cu_lookahead_join.argtypes = [vp]
    L.x264cu_lookahead_get_mvs.argtypes = [vp, ci, ci, ci, vp, vp]
    L.x264cu_lookahead_get_costs.argtypes = [vp, ci, ci, ci, vp]
    L.x264cu_lookahead_get_intra.argtypes = [vp, ci, vp]
    L.x264cu_lookahead_get_row_satds.argtypes = [vp, ci, ci, ci, vp]
    L.x264cu_lookahead_get_weight.argtypes = [vp, ci, C.POINTER(ci)]
    L.x264cu_lookahead_get_cost_est.argtypes = [vp, ci, ci, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    L.x264cu_lookahead_get_lowres_plane.argtypes = [vp, ci, ci, vp, C.POINTER(ss)]


class Lookahead:
    """Mirror of the reference's lookahead offload hooks (x264_opencl_lowres_init / _motionsearch / _finalize_cost,
    encoder/slicetype-cl.c) as one object: frame_put() == lowres_init, frame_cost() == slicetype_frame_cost."""

    def __init__(self, ctx, width, height, subpel_refine=7, me_method=1, me_range=16, mv_range=512, bframes=3,
                 bframe_bias=0, weighted_bipred=1, aq_mode=1, mb_tree=1, vbv=0, n_slots=8, weighted_pred=0):
        self.ctx = ctx
        self.L = ctx.L
        self.p = LookaheadParams(width, height, subpel_refine, me_method, me_range, mv_range, bframes, bframe_bias,
                                 weighted_bipred, aq_mode, mb_tree, vbv, n_slots, weighted_pred)
        h = C.c_void_p()
        ctx.check(self.L.x264cu_lookahead_open(ctx.h, C.byref(self.p), C.byref(h)))
        self.h = h
        self.mb_w, self.mb_h = (width + 15) // 16, (height + 15) // 16
        self.mb_count = self.mb_w * self.mb_h

    def close(self):
        if self.h:
            self.L.x264cu_lookahead_close(self.h)
            self.h = None

    def frame_put(self, slot, luma, inv_qscale=None):
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        q = None
        if inv_qscale is not None:
            q = np.ascontiguousarray(inv_qscale, dtype=np.uint16)
        self.ctx.check(self.L.x264cu_lookahead_frame_put(self.h, slot, luma.ctypes.data, luma.shape[1],
                                                         q.ctypes.data if q is not None else None))

    def frame_put_device(self, slot, d_luma, stride, inv_qscale=None):
        q = None
        if inv_qscale is not None:
            q = np.ascontiguousarray(inv_qscale, dtype=np.uint16)
        self.ctx.check(self.L.x264cu_lookahead_frame_put_device(self.h, slot, int(d_luma), stride,
                                                                q.ctypes.data if q is not None else None))

    def frame_cost(self, frames, p0, p1, b):
        arr = (C.c_int * len(frames))(*frames)
        sc = C.c_int()
        self.ctx.check(self.L.x264cu_lookahead_frame_cost(self.h, arr, p0, p1, b, C.byref(sc)))
        return sc.value

    def search_batch(self, jobs):
        """jobs: list of (fenc_slot, ref_slot, list, dist)"""
        n = len(jobs)
        cols = [(C.c_int * n)(*[j[k] for j in jobs]) for k in range(4)]
        self.ctx.check(self.L.x264cu_lookahead_search_batch(self.h, n, *cols))

    def join(self):
        self.ctx.check(self.L.x264cu_lookahead_join(self.h))

    def get_mvs(self, slot, lst, dist_minus1):
        mv = np.zeros((self.mb_count, 2), np.int16)
        co = np.zeros(self.mb_count, np.int32)
        self.ctx.check(self.L.x264cu_lookahead_get_mvs(self.h, slot, lst, dist_minus1, mv.ctypes.data, co.ctypes.data))
        return mv, co

    def get_costs(self, slot, i0, i1):
        out = np.zeros(self.mb_count, np.uint16)
        self.ctx.check(self.L.x264cu_lookahead_get_costs(self.h, slot, i0, i1, out.ctypes.data))
        return out

    def get_intra(self, slot):
        out = np.zeros(self.mb_count, np.int32)
        self.ctx.check(self.L.x264cu_lookahead_get_intra(self.h, slot, out.ctypes.data))
        return out

    def get_row_satds(self, slot, i0, i1):
        out = np.zeros(self.mb_h, np.int32)
        self.ctx.check(self.L.x264cu_lookahead_get_row_satds(self.h, slot, i0, i1, out.ctypes.data))
        return out

    def get_cost_est(self, slot, i0, i1):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.ctx.check(self.L.x264cu_lookahead_get_cost_est(self.h, slot, i0, i1, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # MB-tree (slicetype.c:1029-1184)
    def set_qp_offset_aq(self, slot, aq):
        a = None if aq is None else np.ascontiguousarray(aq, dtype=np.float32)
        self.ctx.check(self.L.x264cu_lookahead_frame_set_qp_offset_aq(self.h, slot, a.ctypes.data if a is not None else None))

    def mbtree_reset(self, slot):
        self.ctx.check(self.L.x264cu_lookahead_mbtree_reset(self.h, slot))

    def mbtree_propagate(self, frames, p0, p1, b, referenced, fps_factor):
        arr = (C.c_int * len(frames))(*frames)
        self.ctx.check(self.L.x264cu_lookahead_mbtree_propagate(self.h, arr, p0, p1, b, int(referenced), float(fps_factor)))

    def mbtree_finish(self, slot, fps_factor, ref0_distance, strength):
        self.ctx.check(self.L.x264cu_lookahead_mbtree_finish(self.h, slot, int(fps_factor), int(ref0_distance), float(strength)))

    def frame_cost_recalculate(self, slot, dist0, dist1, b_type):
        """slicetype_frame_cost_recalculate -> (cost, row satds)"""
        score = C.c_int()
        rows = np.zeros(self.mb_h, np.int32)
        self.ctx.check(self.L.x264cu_lookahead_frame_cost_recalculate(self.h, slot, dist0, dist1, int(b_type), C.addressof(score), rows.ctypes.data))
        return score.value, rows

    def get_qp_offset(self, slot):
        out = np.zeros(self.mb_count, np.float32)
        self.ctx.check(self.L.x264cu_lookahead_get_qp_offset(self.h, slot, out.ctypes.data))
        return out

    def get_propagate_cost(self, slot):
        out = np.zeros(self.mb_count, np.uint16)
        self.ctx.check(self.L.x264cu_lookahead_get_propagate_cost(self.h, slot, out.ctypes.data))
        return out

    def get_weighted_cost_delta(self, slot, dist_minus1):
        return float(self.L.x264cu_lookahead_get_weighted_cost_delta(self.h, slot, dist_minus1))

    def get_weight(self, slot):
        w = (C.c_int * 4)()
        self.ctx.check(self.L.x264cu_lookahead_get_weight(self.h, slot, w))
        return tuple(w)

    def get_lowres_plane(self, slot, plane):
        st = C.c_ssize_t()
        self.ctx.check(self.L.x264cu_lookahead_get_lowres_plane(self.h, slot, plane, None, C.byref(st)))
        rows = self.mb_h * 8 + 64
        out = np.zeros(rows * st.value, np.uint8)
        self.ctx.check(self.L.x264cu_lookahead_get_lowres_plane(self.h, slot, plane, out.ctypes.data, C.byref(st)))
        return out.reshape(rows, st.value)


class Slicetype:
    """x264_lookahead_put_frame / x264_lookahead_get_frames + x264_slicetype_decide on the GPU lookahead.
    decide(frames) runs a whole sequence the way x264_encoder_encode would and returns [(display_index, type)] in coded order."""

    def __init__(self, ctx, width, height, keyint_max=250, keyint_min=25, scenecut_threshold=40, b_adapt=1, b_pyramid=2,
                 rc_lookahead=40, psy=1, frame_reference=3, rc_cqp=0, fps_num=0, fps_den=0, qcompress=-1.0, aq_strength=-1.0, open_gop=0, intra_refresh=0, **la_kwargs):
        self.ctx, self.L = ctx, ctx.L
        la = dict(subpel_refine=7, me_method=1, me_range=16, mv_range=512, bframes=3, bframe_bias=0, weighted_bipred=1,
                  aq_mode=1, mb_tree=1, vbv=0, n_slots=0, weighted_pred=0)
        la.update(la_kwargs)
        lp = LookaheadParams(width, height, la["subpel_refine"], la["me_method"], la["me_range"], la["mv_range"], la["bframes"],
                             la["bframe_bias"], la["weighted_bipred"], la["aq_mode"], la["mb_tree"], la["vbv"], la["n_slots"], la["weighted_pred"])
        self.p = SlicetypeParams(lp, keyint_max, keyint_min, scenecut_threshold, b_adapt, b_pyramid, rc_lookahead, psy,
                                 frame_reference, rc_cqp, fps_num, fps_den, qcompress, aq_strength, open_gop, intra_refresh)
        h = C.c_void_p()
        if self.L.x264cu_slicetype_open(ctx.h, C.byref(self.p), C.byref(h)) != 0:
            from .binding import X264CUError
            raise X264CUError("x264cu_slicetype_open failed: " + ctx.L.x264cu_strerror(ctx.h).decode())
        self.h = h
        self.mb_count = ((width + 15) // 16) * ((height + 15) // 16)
        self.mb_h = (height + 15) // 16

    @classmethod
    def from_params(cls, ctx, p):
        """from a filled SlicetypeParams (e.g. mirrored from an opened reference encoder by the tests)"""
        return cls(ctx, p.la.width, p.la.height, keyint_max=p.keyint_max, keyint_min=p.keyint_min, scenecut_threshold=p.scenecut_threshold,
                   b_adapt=p.b_adapt, b_pyramid=p.b_pyramid, rc_lookahead=p.rc_lookahead, psy=p.psy, frame_reference=p.frame_reference,
                   rc_cqp=p.rc_cqp, fps_num=p.fps_num, fps_den=p.fps_den, qcompress=p.qcompress, aq_strength=p.aq_strength,
                   open_gop=p.open_gop, intra_refresh=p.intra_refresh,
                   subpel_refine=p.la.subpel_refine, me_method=p.la.me_method, me_range=p.la.me_range, mv_range=p.la.mv_range,
                   bframes=p.la.bframes, bframe_bias=p.la.bframe_bias, weighted_bipred=p.la.weighted_bipred, aq_mode=p.la.aq_mode,
                   mb_tree=p.la.mb_tree, vbv=p.la.vbv, weighted_pred=p.la.weighted_pred)

    def close(self):
        if self.h:
            self.L.x264cu_slicetype_close(self.h)
            self.h = None

    def step(self, luma=None, inv_qscale=None):
        fr, ty = C.c_int(), C.c_int()
        if luma is not None:
            luma = np.ascontiguousarray(luma, dtype=np.uint8)
            q = None if inv_qscale is None else np.ascontiguousarray(inv_qscale, dtype=np.uint16)
            rc = self.L.x264cu_slicetype_step(self.h, luma.ctypes.data, luma.shape[1], q.ctypes.data if q is not None else None,
                                              C.byref(fr), C.byref(ty))
        else:
            rc = self.L.x264cu_slicetype_step(self.h, None, 0, None, C.byref(fr), C.byref(ty))
        self.ctx.check(rc)
        return fr.value, ty.value

    def step_i420(self, luma, cb, cr):
        """one picture with its chroma planes: adaptive quantisation runs on the device (la.aq_mode, aq_strength)"""
        fr, ty = C.c_int(), C.c_int()
        luma, cb, cr = (np.ascontiguousarray(a, dtype=np.uint8) for a in (luma, cb, cr))
        self.ctx.check(self.L.x264cu_slicetype_step_i420(self.h, luma.ctypes.data, luma.shape[1], cb.ctypes.data, cr.ctypes.data, cb.shape[1],
                                                         C.byref(fr), C.byref(ty)))
        return fr.value, ty.value

    def step_device(self, d_luma, stride):
        fr, ty = C.c_int(), C.c_int()
        self.ctx.check(self.L.x264cu_slicetype_step_device(self.h, int(d_luma), stride, None, C.byref(fr), C.byref(ty)))
        return fr.value, ty.value

    def set_prefetch(self, on):
        self.L.x264cu_slicetype_set_prefetch(self.h, int(on))

    def set_run_ahead(self, k):
        self.L.x264cu_slicetype_set_run_ahead(self.h, int(k))

    def set_speculation(self, on):
        self.L.x264cu_slicetype_set_speculation(self.h, int(on))

    def set_prefetch_group(self, k):
        self.L.x264cu_slicetype_set_prefetch_group(self.h, int(k))

    def speculation_stats(self):
        """(triples computed ahead of time, cost requests served from them, cost requests computed on demand)"""
        hits, misses = C.c_long(), C.c_long()
        n = self.L.x264cu_lookahead_speculation_stats(self.L.x264cu_slicetype_lookahead(self.h), C.byref(hits), C.byref(misses))
        return int(n), hits.value, misses.value

    def set_async_upload(self, on):
        """page-locked pictures passed to step() are read in place; keep them unmodified until four more pictures have been queued"""
        self.L.x264cu_slicetype_set_async_upload(self.h, int(on))

    def set_shard(self, rank, world, exchange):
        """one stream over `world` GPUs: exchange = x264_b200.dist.ShardExchange (kept alive by this object)"""
        self._exchange = exchange
        fn = C.cast(exchange.cb, C.c_void_p) if exchange is not None else None
        self.ctx.check(self.L.x264cu_slicetype_set_shard(self.h, int(rank), int(world), fn, None))

    def search_stats(self):
        """(device ms of the search launches so far, launches, searches) of the lookahead underneath"""
        b, l, n = C.c_double(), C.c_long(), C.c_long()
        self.ctx.check(self.L.x264cu_lookahead_search_stats(self.L.x264cu_slicetype_lookahead(self.h), C.byref(b), C.byref(l), C.byref(n)))
        return b.value, l.value, n.value

    def get_qp_offset(self, frame):
        """f_qp_offset (MB-tree) of a non-B picture just returned by step()"""
        out = np.zeros(self.mb_count, np.float32)
        self.ctx.check(self.L.x264cu_slicetype_get_qp_offset(self.h, int(frame), out.ctypes.data))
        return out

    def set_next_type(self, t):
        self.ctx.check(self.L.x264cu_slicetype_set_next_type(self.h, int(t)))

    def rc_analyse_slice(self, frame):
        """x264_rc_analyse_slice of the picture step() just returned -> (cost, row satds)"""
        cost = C.c_int()
        rows = np.zeros(self.mb_h, np.int32)
        self.ctx.check(self.L.x264cu_slicetype_rc_analyse_slice(self.h, int(frame), C.addressof(cost), rows.ctypes.data, None))
        return cost.value, rows

    def get_planned(self, frame, max_entries=32):
        """i_planned_type / i_planned_satd (VBV lookahead) of a non-B picture step() just returned"""
        t, sd = np.zeros(max_entries, np.int32), np.zeros(max_entries, np.int32)
        k = self.L.x264cu_slicetype_get_planned(self.h, int(frame), t.ctypes.data, sd.ctypes.data, max_entries)
        if k < 0:
            raise RuntimeError("x264cu_slicetype_get_planned failed")
        return list(t[:k]), list(sd[:k])

    def decide(self, frames, qp_out=None, chroma=None, forced=None, rc_out=None, vbv=False):
        """qp_out: dict filled with frame -> f_qp_offset for every non-B picture; chroma: [(cb, cr)] per picture -> step_i420;
        forced: pic_in.i_type per picture (0 = auto)"""
        out = []

        def note(fr, ty):
            out.append((fr, ty))
            if rc_out is not None:      # (cost, rows, planned types, planned costs) as tests/test_slicetype_host.py collects them
                cost, rows = (-1, np.zeros(self.mb_h, np.int32)) if ty in (4, 5) and not vbv else self.rc_analyse_slice(fr)
                pt, ps = self.get_planned(fr) if vbv and ty not in (4, 5) else ([], [])
                rc_out[fr] = (cost, rows, pt, ps)
            if qp_out is not None and ty not in (4, 5):
                qp_out[fr] = self.get_qp_offset(fr)

        for i, f in enumerate(frames):
            if forced is not None and forced[i]:
                self.set_next_type(forced[i])
            fr, ty = self.step(f) if chroma is None else self.step_i420(f, chroma[i][0], chroma[i][1])
            if fr >= 0:
                note(fr, ty)
        while True:
            fr, ty = self.step(None)
            if fr < 0:
                break
            note(fr, ty)
        return out

    @property
    def cost_requests(self):
        return int(self.L.x264cu_slicetype_cost_requests(self.h))


def me_search_batch(ctx, params, d_fenc, fenc_stride, d_fref, d_fref_w, ref_stride, jobs):
    """jobs: numpy array of me_job_dtype (host) -> numpy array of me_result_dtype.  d_* are device addresses."""
    assert jobs.dtype == me_job_dtype and C.sizeof(MeJob) == me_job_dtype.itemsize and C.sizeof(MeResult) == me_result_dtype.itemsize
    n = len(jobs)
    d_jobs = ctx.upload(jobs)
    d_res = ctx.malloc(max(n, 1) * me_result_dtype.itemsize)
    arr = (C.c_void_p * 4)(*[int(p) for p in d_fref])
    ctx.check(ctx.L.x264cu_me_search_batch(ctx.h, C.byref(params), int(d_fenc), fenc_stride, arr,
                                           int(d_fref_w) if d_fref_w else None, ref_stride, d_jobs, n, d_res))
    out = ctx.download(d_res, (n,), me_result_dtype)
    ctx.free(d_jobs)
    ctx.free(d_res)
    return out


def make_me_frame(d_fenc, fenc_stride, d_fenc_uv, fenc_uv_stride, ref_stride, ref_uv_stride, refs, lambdas, chroma_me):
    """refs: [(d_fref[4], d_fref_w or None, d_fref_uv or None, weights[3][4])] -> (MeFrame, keep-alive objects)"""
    arr = (MeRef * len(refs))()
    for i, (pl, w, uv, wt) in enumerate(refs):
        for k in range(4):
            arr[i].d_fref[k] = int(pl[k])
        arr[i].d_fref_w = int(w) if w else None
        arr[i].d_fref_uv = int(uv) if uv else None
        for k in range(3):
            for j in range(4):
                arr[i].weight[k][j] = int(wt[k][j])
    lam = (C.c_int * len(lambdas))(*[int(v) for v in lambdas])
    f = MeFrame(int(d_fenc), fenc_stride, int(d_fenc_uv) if d_fenc_uv else None, fenc_uv_stride, ref_stride, ref_uv_stride,
                arr, len(refs), lam, len(lambdas), int(chroma_me))
    return f, (arr, lam)


def me_search_frame(ctx, params, frame, jobs):
    """jobs: numpy array of me_frame_job_dtype (host) -> numpy array of me_result_dtype"""
    assert jobs.dtype == me_frame_job_dtype and me_frame_job_dtype.itemsize == C.sizeof(MeJob) + 4
    n = len(jobs)
    d_jobs = ctx.upload(jobs)
    d_res = ctx.malloc(max(n, 1) * me_result_dtype.itemsize)
    ctx.check(ctx.L.x264cu_me_search_frame(ctx.h, C.byref(params), C.byref(frame), d_jobs, n, d_res))
    out = ctx.download(d_res, (n,), me_result_dtype)
    ctx.free(d_jobs)
    ctx.free(d_res)
    return out


def me_refine_bidir_batch(ctx, params, d_fenc, fenc_stride, d_fref0, d_fref1, ref_stride, jobs):
    """jobs: numpy array of bidir_job_dtype (host) -> numpy array of bidir_result_dtype.  d_* are device addresses."""
    assert jobs.dtype == bidir_job_dtype and bidir_job_dtype.itemsize == 44 and bidir_result_dtype.itemsize == 12
    n = len(jobs)
    d_jobs = ctx.upload(jobs)
    d_res = ctx.malloc(max(n, 1) * bidir_result_dtype.itemsize)
    a0 = (C.c_void_p * 4)(*[int(p) for p in d_fref0])
    a1 = (C.c_void_p * 4)(*[int(p) for p in d_fref1])
    ctx.check(ctx.L.x264cu_me_refine_bidir_batch(ctx.h, C.byref(params), int(d_fenc), fenc_stride, a0, a1, ref_stride, d_jobs, n, d_res))
    out = ctx.download(d_res, (n,), bidir_result_dtype)
    ctx.free(d_jobs)
    ctx.free(d_res)
    return out


def me_refine_qpel_batch(ctx, params, refdupe, d_fenc, fenc_stride, d_fref, ref_stride, jobs):
    """jobs: numpy array of me_refine_job_dtype (host) -> numpy array of me_result_dtype.  d_* are device addresses."""
    assert jobs.dtype == me_refine_job_dtype and me_refine_job_dtype.itemsize == 40
    n = len(jobs)
    d_jobs = ctx.upload(jobs)
    d_res = ctx.malloc(max(n, 1) * me_result_dtype.itemsize)
    arr = (C.c_void_p * 4)(*[int(p) for p in d_fref])
    ctx.check(ctx.L.x264cu_me_refine_qpel_batch(ctx.h, C.byref(params), int(refdupe), int(d_fenc), fenc_stride, arr, ref_stride, d_jobs, n, d_res))
    out = ctx.download(d_res, (n,), me_result_dtype)
    ctx.free(d_jobs)
    ctx.free(d_res)
    return out
