"""x264_b200 -- B200-native (sm_100a) backend for x264's motion-estimation / lookahead cost path.

The product is the C-ABI shared library x264_b200/csrc/libx264_b200.so (include/x264_b200.h).  This package
is the thin Python binding used by the tests and bench.py: it loads the library with ctypes and mirrors the
reference's interfaces for this path (x264_pixel_function_t table -> PixelFunctions, x264_mc_functions_t ->
McFunctions, the lookahead hooks -> Lookahead).  There is no CPU fallback: importing works without a GPU (so
that the build and the symbol table can be checked), but every compute call needs a CUDA device.
"""
from .binding import (lib, lib_path, Context, X264CUError, PIXEL_W, PIXEL_H, PIXEL_NAMES, cand_dtype,
                      cand_x4_dtype, SAD, SSD, SATD, SA8D, PAD, exported_symbols, header_symbols)

from .binding_ext import (Lookahead, LookaheadParams, Slicetype, SlicetypeParams, TYPE_NAMES, MeParams, me_job_dtype,
                          me_result_dtype, me_search_batch, bidir_job_dtype, bidir_result_dtype, me_refine_bidir_batch,
                          me_refine_job_dtype, me_refine_qpel_batch, me_frame_job_dtype, MeRef, MeFrame, make_me_frame,
                          me_search_frame)

__all__ = ["me_frame_job_dtype", "MeRef", "MeFrame", "make_me_frame", "me_search_frame", "me_refine_job_dtype", "me_refine_qpel_batch", "bidir_job_dtype", "bidir_result_dtype", "me_refine_bidir_batch", "MeParams", "me_job_dtype", "me_result_dtype", "me_search_batch", "Lookahead", "LookaheadParams", "Slicetype", "SlicetypeParams", "TYPE_NAMES", "lib", "lib_path", "Context", "X264CUError", "PIXEL_W", "PIXEL_H", "PIXEL_NAMES", "cand_dtype",
           "cand_x4_dtype", "SAD", "SSD", "SATD", "SA8D", "PAD", "exported_symbols", "header_symbols"]
