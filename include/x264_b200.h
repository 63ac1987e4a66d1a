/*
 * x264_b200 -- B200-native (sm_100a) backend for x264's motion-estimation / lookahead cost path.
 *
 * C ABI (plain pointers and sizes, no C++/torch types).  These are the entry points a host written in C
 * (the x264 encoder itself) binds; INTEGRATION.md shows the glue on the reference side.
 * Every entry cites the reference interface it replaces (paths relative to jpsdr/x264).
 *
 * Conventions
 *   - all functions return 0 on success, -1 on failure; x264cu_strerror() gives the reason.  There is no
 *     CPU fallback anywhere: a missing device or a failed launch is an error.
 *   - "d_" pointers are device (HBM) pointers, "h_" pointers are host pointers.
 *   - planes are 8-bit luma in the reference's padded frame layout (common/frame.c:75-130): pixel (x,y)
 *     lives at origin + y*stride + x and a border of X264CU_PAD pixels around the picture is readable.
 *   - block-size indices are the reference's PIXEL_16x16 .. PIXEL_4x16 (common/pixel.h:37-52).
 *   - calls on one context are serialised on that context's stream (the reference calls its OpenCL hooks
 *     from exactly one thread too, encoder/lookahead.c:90).
 */
#ifndef X264_B200_H
#define X264_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X264CU_PAD 32                      /* PADH / PADV, common/frame.h:32-33 */
#define X264CU_BFRAME_MAX 16               /* X264_BFRAME_MAX, common/base.h:136 */

enum x264cu_pixel_e                        /* common/pixel.h:37-52 */
{
    X264CU_PIXEL_16x16 = 0, X264CU_PIXEL_16x8 = 1, X264CU_PIXEL_8x16 = 2, X264CU_PIXEL_8x8 = 3,
    X264CU_PIXEL_8x4 = 4, X264CU_PIXEL_4x8 = 5, X264CU_PIXEL_4x4 = 6, X264CU_PIXEL_4x16 = 7,
    X264CU_PIXEL_NB = 8
};

enum x264cu_metric_e                       /* members of x264_pixel_function_t, common/pixel.h:78-144 */
{
    X264CU_SAD = 0,                        /* pixf.sad[]   common/pixel.c:55-80   */
    X264CU_SSD = 1,                        /* pixf.ssd[]   common/pixel.c:85-110  */
    X264CU_SATD = 2,                       /* pixf.satd[]  common/pixel.c:242-332 */
    X264CU_SA8D = 3                        /* pixf.sa8d[]  common/pixel.c:334-381 (16x16 and 8x8 only) */
};

enum x264cu_me_e { X264CU_ME_DIA = 0, X264CU_ME_HEX = 1, X264CU_ME_UMH = 2, X264CU_ME_ESA = 3, X264CU_ME_TESA = 4 };   /* x264.h X264_ME_* */

typedef struct x264cu_ctx x264cu_ctx_t;

/* ------------------------------------------------------------------------------------------------
 * context -- replaces x264_opencl_load_library / x264_opencl_lookahead_init / _delete
 * (common/opencl.h:796-804, common/opencl.c:53, :411, :596).
 * ---------------------------------------------------------------------------------------------- */
int         x264cu_open( x264cu_ctx_t **ctx, int device );
void        x264cu_close( x264cu_ctx_t *ctx );
const char *x264cu_strerror( x264cu_ctx_t *ctx );            /* ctx may be NULL: last open() failure */
int         x264cu_device_info( x264cu_ctx_t *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes );
void       *x264cu_stream( x264cu_ctx_t *ctx );              /* the cudaStream_t all work is enqueued on */
int         x264cu_sync( x264cu_ctx_t *ctx );                /* x264_opencl_flush, encoder/slicetype-cl.c:58 */
/* number of kernels this context has launched since open (bench.py's gpu_launches) */
uint64_t    x264cu_launch_count( x264cu_ctx_t *ctx );

/* device-side stopwatch on the context's stream (CUDA events); used by bench.py for kernel timing */
int         x264cu_timer_start( x264cu_ctx_t *ctx );
int         x264cu_timer_stop( x264cu_ctx_t *ctx, float *elapsed_ms );    /* synchronises */

/* device memory, so that a C host needs no CUDA headers */
void *x264cu_malloc( x264cu_ctx_t *ctx, size_t bytes );
void  x264cu_free( x264cu_ctx_t *ctx, void *d_ptr );
void *x264cu_malloc_host( x264cu_ctx_t *ctx, size_t bytes );  /* page-locked staging (opencl.h:718) */
void  x264cu_free_host( x264cu_ctx_t *ctx, void *h_ptr );
int   x264cu_memcpy_h2d( x264cu_ctx_t *ctx, void *d_dst, const void *h_src, size_t bytes );
int   x264cu_memcpy_d2h( x264cu_ctx_t *ctx, void *h_dst, const void *d_src, size_t bytes );
int   x264cu_memset( x264cu_ctx_t *ctx, void *d_dst, int value, size_t bytes );

/* ------------------------------------------------------------------------------------------------
 * B1: batched twins of the x264_pixel_function_t table (common/pixel.h:78-144).
 * A GPU cannot sit behind int cmp(pixel*,intptr_t,pixel*,intptr_t) one 256-byte block at a time, so
 * each table entry is exported as "the same function over an array of candidates".
 * ---------------------------------------------------------------------------------------------- */

/* one candidate = one call of x264_pixel_cmp_t (common/pixel.h:33): byte offsets of the two blocks */
typedef struct { uint32_t fenc_off, ref_off; } x264cu_cand_t;
/* one x4 call (common/pixel.h:35): one fenc block against four reference blocks; x3 uses n_refs = 3 */
typedef struct { uint32_t fenc_off, ref_off[4]; } x264cu_cand_x4_t;

/* out[i] = pixf.<metric>[i_pixel]( d_fenc + cand[i].fenc_off, fenc_stride, d_ref + cand[i].ref_off, ref_stride ) */
int x264cu_pixel_cmp_batch( x264cu_ctx_t *ctx, int metric, int i_pixel,
                            const uint8_t *d_fenc, intptr_t fenc_stride,
                            const uint8_t *d_ref, intptr_t ref_stride,
                            const x264cu_cand_t *d_cand, int n, int32_t *d_out );

/* pixf.sad_x3/x4, satd_x3/x4 (common/pixel.c:441-496): out[i*4+j], j < n_refs */
int x264cu_pixel_cmp_x4_batch( x264cu_ctx_t *ctx, int metric, int i_pixel, int n_refs,
                               const uint8_t *d_fenc, intptr_t fenc_stride,
                               const uint8_t *d_ref, intptr_t ref_stride,
                               const x264cu_cand_x4_t *d_cand, int n, int32_t *d_out );

/* Host-buffer form of x264cu_pixel_cmp_batch: copies both planes and the candidate list to the device,
 * runs the kernel, copies the costs back (what an encoder thread holding host frames would call). */
int x264cu_pixel_cmp_batch_host( x264cu_ctx_t *ctx, int metric, int i_pixel,
                                 const uint8_t *h_fenc, size_t fenc_bytes, intptr_t fenc_stride,
                                 const uint8_t *h_ref, size_t ref_bytes, intptr_t ref_stride,
                                 const x264cu_cand_t *h_cand, int n, int32_t *h_out );

/* A stack of equally sized padded planes in HBM (e.g. the luma planes of consecutive frames). */
typedef struct
{
    const uint8_t *d_origin;   /* pixel (0,0) of plane 0 */
    intptr_t stride;           /* bytes per row, multiple of 16 */
    intptr_t plane_pitch;      /* bytes between consecutive planes, multiple of 16 */
    int width, height;         /* picture size; X264CU_PAD pixels around it are readable */
    int n_planes;
} x264cu_planes_t;

/* MV-field form: the picture is tiled into blocks of size i_pixel; block (bx,by) of plane f is compared with the
 * reference block displaced by the full-pel vector mv = d_mv[((k*n_planes + f)*blocks_y + by)*blocks_x + bx]
 * (int16 x, int16 y), |mv| <= X264CU_PAD.  out has the same indexing.  This is the shape in which
 * x264_me_search_ref evaluates candidates (encoder/me.c:63-141 COST_MV*): every block of the frame, one
 * displaced reference block each, k_cands times.  Reference tiles (+halo) are staged through shared
 * memory with TMA, so each byte of both planes crosses HBM once. */
int x264cu_pixel_cmp_mvfield( x264cu_ctx_t *ctx, int metric, int i_pixel,
                              const x264cu_planes_t *fenc, const x264cu_planes_t *ref,
                              int k_cands, const int16_t *d_mv, int32_t *d_out );

/* host-buffer form (planes and mv field on the host, padded layout as above; h_*_base = first byte of the
 * allocation, i.e. origin - X264CU_PAD*stride - X264CU_PAD) */
int x264cu_pixel_cmp_mvfield_host( x264cu_ctx_t *ctx, int metric, int i_pixel,
                                   const uint8_t *h_fenc_base, const uint8_t *h_ref_base,
                                   intptr_t stride, intptr_t plane_pitch, int width, int height, int n_planes,
                                   int k_cands, const int16_t *h_mv, int32_t *h_out );

/* ------------------------------------------------------------------------------------------------
 * B1 (mc table): frame preparation twins of x264_mc_functions_t entries (common/mc.h:267-340).
 * ---------------------------------------------------------------------------------------------- */

/* x264_frame_init_lowres (common/mc.c:458-507) + x264_frame_expand_border_lowres (common/frame.c:627-631):
 * d_luma is the picture as the user supplied it (width x height, any stride); it is treated as expanded to
 * the next multiple of 16 by edge replication (x264_frame_expand_border_mod16, common/frame.c:640-665).
 * d_lowres[i] are the origins of the F,H,V,C planes (lowres_stride bytes per row, X264CU_PAD border each side,
 * width_lowres = 8*ceil(width/16), lines_lowres = 8*ceil(height/16)). */
int x264cu_frame_init_lowres( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                              uint8_t *const d_lowres[4], intptr_t lowres_stride );

/* the same for a stack of n_pictures pictures in one launch: picture p's luma at d_luma + p * luma_pitch, its four planes at
 * d_lowres[i] + p * lowres_pitch */
int x264cu_frame_init_lowres_batch( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, intptr_t luma_pitch, int n_pictures,
                                    int width, int height, uint8_t *const d_lowres[4], intptr_t lowres_stride, intptr_t lowres_pitch );

/* hpel_filter as driven over a whole frame by x264_frame_filter + x264_frame_expand_border_filtered
 * (common/mc.c:172-196, :704-746; common/frame.c:596-625): fills the H, V and C half-pel planes (origins given,
 * same stride, X264CU_PAD border) from a width x height luma plane; also (re)writes the border of d_src itself
 * when expand_src != 0 (x264_frame_expand_border, frame.c:562-594). */
int x264cu_hpel_filter( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, int width, int height,
                        uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src );

/* the same over a stack of n_planes pictures in one launch (plane p of each of the four stacks at + p * plane_pitch bytes): what a
 * caller holding several reconstructed pictures -- or a measurement that wants inputs larger than the L2 -- uses */
int x264cu_hpel_filter_batch( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, intptr_t plane_pitch, int n_planes, int width, int height,
                              uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src );

/* mc_luma / get_ref (common/mc.c:198-249, tables x264_hpel_ref0/1 common/tables.c:183-184): job i produces the w x h block
 * (i_pixel) of the reference at quarter-pel vector (mvx, mvy) -- one half-pel plane or the rounded mean of two, then the
 * optional explicit weight (mc_weight, mc.c:117-137) -- into d_dst + i*w*h (rows packed, stride w).  d_src = F,H,V,C planes
 * (x264cu_hpel_filter output, padded); src_off = byte offset of the block at vector 0, the same in all four planes.
 * weight = { enabled, i_scale, i_denom, i_offset } (x264_weight_t, common/mc.h:235-245) or NULL.  get_ref returns the same
 * pixels (it only avoids the copy when no interpolation is needed, mc.c:244-248). */
typedef struct { uint32_t src_off; int16_t mvx, mvy; } x264cu_mc_job_t;
int x264cu_mc_luma_batch( x264cu_ctx_t *ctx, const uint8_t *const d_src[4], intptr_t src_stride, int i_pixel,
                          const x264cu_mc_job_t *d_jobs, int n, const int weight[4], uint8_t *d_dst );

/* mc.avg[i_pixel] (pixel_avg_WxH, common/mc.c:49-111): n packed w x h blocks a, b -> dst; weight 32 is the rounded mean,
 * anything else clip( (a*weight + b*(64-weight) + 32) >> 6 ) (bi-prediction, i_bipred_weight) */
int x264cu_pixel_avg_batch( x264cu_ctx_t *ctx, int i_pixel, const uint8_t *d_a, const uint8_t *d_b, int n, int weight, uint8_t *d_dst );

/* x264_weight_scale_plane (common/frame.c:825-841) = mc.weight over a whole plane: dst = clip( ((src*scale + 2^(denom-1)) >> denom)
 * + offset ), or src*scale + offset when denom is 0 (mc_weight, common/mc.c:117-137).  weight = { -, scale, denom, offset } */
int x264cu_weight_scale_plane( x264cu_ctx_t *ctx, const uint8_t *d_src, uint8_t *d_dst, intptr_t stride, int width, int height,
                               const int weight[4] );

/* x264_pixel_ssd_wxh (common/pixel.c:112-151): sum of squared differences of two width x height planes (PSNR); synchronises */
int x264cu_pixel_ssd_wxh( x264cu_ctx_t *ctx, const uint8_t *d_pix1, intptr_t stride1, const uint8_t *d_pix2, intptr_t stride2,
                          int width, int height, uint64_t *h_ssd );

/* Successive elimination (a7).  x264cu_integral_init = what integral_init4h / 8h / 4v / 8v leave in frame->integral once
 * x264_frame_filter has run over a reference frame (common/mc.c:424-456, :748-783): d_sum8[y*stride + x] = the sum of the 8x8 pixel
 * box with its top-left corner at (x, y), for every position of the padded plane where the box fits ((-PAD .. width+PAD-8) x
 * (-PAD .. height+PAD-8)); d_sum4 (may be NULL; the reference keeps it with sub-8x8 partitions) the same for 4x4 boxes.
 * d_plane / d_sum8 / d_sum4 point at position (0,0); stride in ELEMENTS, the pixel plane's. */
int x264cu_integral_init( x264cu_ctx_t *ctx, const uint8_t *d_plane, intptr_t stride, int width, int height,
                          uint16_t *d_sum8, uint16_t *d_sum4 );
/* pixf.ads[i_pixel] (x264_pixel_ads1 / 2 / 4, common/pixel.c:759-803) over many rows of candidates: job = one call --
 * ( enc_dc, sums = d_sums + sums_off, delta, cost_mvx = d_cost_mvx + cost_off, width, thresh ).  For every job the indices i <
 * width with  sum_k |enc_dc[k] - sums[i + offset_k]| + cost_mvx[i] < thresh  are written in ascending order to d_mvs + out_off
 * (room for `width` entries) and their number to d_counts[job]: mvs[] and the return value of the reference's function. */
typedef struct
{
    int32_t  enc_dc[4];              /* DC of the block's 8x8 (4x4) sub-blocks: me.c:644-651 */
    uint32_t sums_off;               /* element offset of the row's first position in the sums plane */
    int32_t  delta;                  /* me.c:639-649: 8 or 4, times the stride for 16x16, 8x16 and 4x8 */
    uint32_t cost_off;               /* offset of cost_mvx[0] in d_cost_mvx */
    int32_t  width, thresh;
    uint32_t out_off;                /* where this job's list starts in d_mvs */
} x264cu_ads_job_t;
int x264cu_pixel_ads_batch( x264cu_ctx_t *ctx, int i_pixel, const uint16_t *d_sums, const uint16_t *d_cost_mvx,
                            const x264cu_ads_job_t *d_jobs, int n, int32_t *d_counts, int16_t *d_mvs );

/* Input staging (N4): the plane-copy entries of the mc table (common/mc.c:294-339, x264_mc_functions_t.plane_copy_swap /
 * _interleave / _deinterleave; a plain plane_copy is cudaMemcpy2D) on planes in HBM: w x h byte PAIRS. */
int x264cu_plane_copy_interleave( x264cu_ctx_t *ctx, uint8_t *d_dst, intptr_t dst_stride, const uint8_t *d_srcu, intptr_t srcu_stride,
                                  const uint8_t *d_srcv, intptr_t srcv_stride, int w, int h );
int x264cu_plane_copy_deinterleave( x264cu_ctx_t *ctx, uint8_t *d_dsta, intptr_t dsta_stride, uint8_t *d_dstb, intptr_t dstb_stride,
                                    const uint8_t *d_src, intptr_t src_stride, int w, int h );
int x264cu_plane_copy_swap( x264cu_ctx_t *ctx, uint8_t *d_dst, intptr_t dst_stride, const uint8_t *d_src, intptr_t src_stride, int w, int h );
/* x264_frame_copy_picture (common/frame.c:363-480) for the 8-bit 4:2:0 colour spaces: a picture in host memory (x264_image_t:
 * i_csp = X264_CSP_I420 2 / YV12 3 / NV12 4 / NV21 5, optionally | X264_CSP_VFLIP 0x1000; plane pointers and strides) becomes the
 * reference's internal frame in HBM: the luma plane and ONE interleaved Cb/Cr plane (width/2 pairs x height/2 rows; d_chroma may
 * be NULL).  Enqueued on the context's stream; the host planes may be reused after x264cu_sync. */
int x264cu_frame_copy_picture( x264cu_ctx_t *ctx, int i_csp, const uint8_t *const h_plane[3], const int stride[3], int width, int height,
                               uint8_t *d_luma, intptr_t luma_stride, uint8_t *d_chroma, intptr_t chroma_stride );

/* x264_adaptive_quant_frame( h, frame, NULL ) (encoder/ratecontrol.c:305-420), aq-mode 0 - 3: per macroblock the AC energy of
 * the 16x16 luma block and the two 8x8 chroma blocks (ac_energy_mb), f_qp_offset_aq = aq_strength * 1.0397 * (x264_log2(energy)
 * - 14.427) and i_inv_qscale_factor = x264_exp2fix8 of it -- the two per-macroblock inputs of the lookahead (frame_put's
 * h_inv_qscale, frame_set_qp_offset_aq) -- from an I420 picture in HBM (luma width x height, Cb / Cr (width+1)/2 x (height+1)/2),
 * treated as edge-replicated to the macroblock grid.  Outputs stay on the device (mb_count entries each); h_stats (optional,
 * synchronises) = i_pixel_sum[3], i_pixel_ssd[3] as the function leaves them.  aq-mode 0 gives offsets 0 / factors 256, which is
 * what the reference initialises for MB-tree.  Modes 2 / 3 (auto-variance): qp = (energy+1)^(1/8) recentred on its frame mean,
 * which is summed in single precision in the reference's raster order (one device thread) so that the result is bit-identical. */
int x264cu_adaptive_quant_frame( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, const uint8_t *d_cb, const uint8_t *d_cr,
                                 intptr_t chroma_stride, int width, int height, int aq_mode, float aq_strength,
                                 float *d_qp_offset_aq, uint16_t *d_inv_qscale, uint64_t *h_stats );

/* ------------------------------------------------------------------------------------------------
 * B2: the lowres lookahead, the seam where common/opencl.c + encoder/slicetype-cl.c sit today
 * (hooks called from slicetype_frame_cost, encoder/slicetype.c:878-897).  Results are those of the
 * reference's CPU path (slicetype_mb_cost, slicetype.c:514-791) with one lookahead thread, bit for bit.
 * ---------------------------------------------------------------------------------------------- */
typedef struct x264cu_lookahead x264cu_lookahead_t;

typedef struct
{
    int width, height;            /* h->param.i_width / i_height */
    int subpel_refine;            /* h->param.analyse.i_subpel_refine */
    int me_method;                /* h->param.analyse.i_me_method (X264CU_ME_*) */
    int me_range;                 /* h->param.analyse.i_me_range */
    int mv_range;                 /* h->param.analyse.i_mv_range (after validation) */
    int bframes;                  /* h->param.i_bframe */
    int bframe_bias;              /* h->param.i_bframe_bias */
    int weighted_bipred;          /* h->param.analyse.b_weighted_bipred */
    int aq_mode;                  /* h->param.rc.i_aq_mode != 0 */
    int mb_tree;                  /* h->param.rc.b_mb_tree */
    int vbv;                      /* h->param.rc.i_vbv_buffer_size != 0: row SATDs kept, edge macroblocks costed, VBV lookahead in slicetype */
    int n_slots;                  /* frames resident in HBM at once (>= lookahead + bframes + 3) */
    int weighted_pred;            /* h->param.analyse.i_weighted_pred: non-zero runs the lookahead weight analysis (slicetype.c:284-501)
                                     on first P-type searches; -1 = X264_WEIGHTP_FAKE (encoder.c:1316-1317), which also records
                                     f_weighted_cost_delta for MB-tree (slicetype.c:462-463) */
} x264cu_lookahead_params_t;

/* x264_opencl_lookahead_init / _delete (common/opencl.c:411, :596) */
int  x264cu_lookahead_open( x264cu_ctx_t *ctx, const x264cu_lookahead_params_t *params, x264cu_lookahead_t **out );
void x264cu_lookahead_close( x264cu_lookahead_t *la );

/* x264_opencl_lowres_init (encoder/slicetype-cl.c:82): upload the luma of one frame into `slot`, build its
 * lowres planes on the device and reset its memoised costs / vectors (x264_frame_init_lowres, mc.c:458-482).
 * h_inv_qscale: fenc->i_inv_qscale_factor (u16 per MB) or NULL for 256 (AQ off). */
int x264cu_lookahead_frame_put( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride,
                                const uint16_t *h_inv_qscale );
/* The same from an I420 picture, with x264_adaptive_quant_frame (encoder.c:3417, ratecontrol.c:305-420) run on the device:
 * i_inv_qscale_factor and f_qp_offset_aq / f_qp_offset of the slot come from the picture itself (aq_mode 0..3,
 * aq_strength = h->param.rc.f_aq_strength).  Cb / Cr: (width+1)/2 x (height+1)/2, chroma_stride bytes per row. */
int x264cu_lookahead_frame_put_i420( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride,
                                     const uint8_t *h_cb, const uint8_t *h_cr, intptr_t chroma_stride, int aq_mode, float aq_strength );
/* How h_luma is read.  Pageable memory is staged through the library's own pinned ring (the caller's buffer is free on
 * return).  Page-locked memory (x264cu_malloc_host; the reference stages through page-locked buffers too, opencl.h:718) is
 * read in place by the copy engine on the lookahead's upload stream: with async_upload = 0 (default) the call returns once
 * that copy has finished; with async_upload = N > 0 it returns at once, up to N copies are in flight (1 means 4, at most 32)
 * and the buffer must stay unmodified until N more pictures have been queued on this lookahead (or x264cu_sync) -- x264 itself
 * holds a queued x264_frame_t much longer: its planes are untouched until the frame leaves the lookahead. */
void x264cu_lookahead_set_async_upload( x264cu_lookahead_t *la, int on );
/* same with the luma already in HBM */
int x264cu_lookahead_frame_put_device( x264cu_lookahead_t *la, int slot, const uint8_t *d_luma, intptr_t luma_stride,
                                       const uint16_t *h_inv_qscale );

/* slicetype_frame_cost (encoder/slicetype.c:836-995): frames[] maps the reference's frame indices to slots
 * (frames[p0], frames[b], frames[p1] are read).  Returns the frame score in *score. */
int x264cu_lookahead_frame_cost( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int *score );

/* x264_opencl_slicetype_prep (encoder/slicetype-cl.c:653): run a batch of lowres motion searches in ONE launch so
 * that the GPU is filled; job i searches frame slot fenc[i] against slot ref[i] as list[i] (0/1) at distance
 * dist[i] >= 1.  Searches are pure functions of the two frames (no weighted prediction), so doing them ahead of
 * time does not change any later x264cu_lookahead_frame_cost result.  Already-searched jobs are skipped. */
int x264cu_lookahead_search_batch( x264cu_lookahead_t *la, int n_jobs, const int *fenc, const int *ref,
                                   const int *list, const int *dist );
/* ---- one picture stream sharded over several GPUs: a lowres search is a pure function of two pictures, so the
 * searches of a window can be split between GPUs that all hold the pictures; what has to travel is the result of each search --
 * lowres_mvs[list][dist-1] and lowres_mv_costs[list][dist-1] of the searched picture, x264cu_lookahead_search_bytes() =
 * 8 bytes per macroblock -- in ONE all-gather per window.  export packs a locally searched result into d_dst, import installs
 * a result searched elsewhere (the pair then counts as searched here; the reference-order state of x264cu_lookahead_frame_cost
 * is untouched); both are enqueued on x264cu_lookahead_exchange_stream(), the stream the caller's collective must run on;
 * import_done publishes everything imported since the last call to later cost requests. ---- */
size_t x264cu_lookahead_search_bytes( x264cu_lookahead_t *la );
void  *x264cu_lookahead_exchange_stream( x264cu_lookahead_t *la );
int    x264cu_lookahead_export_search( x264cu_lookahead_t *la, int slot, int list, int dist, void *d_dst );
int    x264cu_lookahead_import_search( x264cu_lookahead_t *la, int slot, int list, int dist, const void *d_src );
int    x264cu_lookahead_import_done( x264cu_lookahead_t *la );

/* Cost requests answered ahead of time: finalize -- everything slicetype_mb_cost does once the vectors exist (slicetype.c:579-652,
 * :706-790) -- of n triples in ONE launch, their result records read back in one copy.  Triple i = (picture in b_slot[i], list-0
 * reference in p0_slot[i] at distance d0[i] >= 1, list-1 reference in p1_slot[i] at distance d1[i]; d1 = 0: a P cost, p1_slot =
 * b_slot).  Computed in the variant that uses the later reference's list-0 vectors for the temporal-direct candidate
 * (slicetype.c:629-642), which is what the reference's request order yields nearly always; x264cu_lookahead_frame_cost uses the
 * record when the request it serves is of that variant and falls back to computing on demand otherwise, so results never depend
 * on what was speculated.  Triples whose searches have not been launched (or were weighted), that are already answered, and every
 * triple when row sums are kept (vbv) are skipped.  Runs on its own low-priority stream behind the searches it reads. */
int x264cu_lookahead_finalize_batch( x264cu_lookahead_t *la, int n, const int *b_slot, const int *p0_slot, const int *p1_slot,
                                     const int *d0, const int *d1 );
/* triples computed ahead of time so far; *hits = cost requests served from them, *misses = computed on demand */
long x264cu_lookahead_speculation_stats( x264cu_lookahead_t *la, long *hits, long *misses );

/* 1 if x264_weights_analyse( fenc, ref ) would leave at its early exit (means and variances of the two pictures agree:
 * no weight, slicetype.c:316-330) -- then the list-0 search of that pair is the same whether it is first requested as a P or
 * as a B cost, and may be run ahead of time with x264cu_lookahead_search_batch.  0 if a weight may be chosen (fade): that
 * search must wait for its first cost request.  Always 1 without weighted prediction.  -1 on error.  Waits for the uploads
 * of the two pictures only (their luma statistics travel back with them). */
int x264cu_lookahead_weight_trivial( x264cu_lookahead_t *la, int fenc_slot, int ref_slot );
/* device time of the search launches so far (CUDA events on the stream each was launched on; waits for those in flight), their
 * number and the number of searches they carried: bench.py's live timing of the dominant kernel */
int x264cu_lookahead_search_stats( x264cu_lookahead_t *la, double *busy_ms, long *launches, long *searches );
/* x264_opencl_flush without the host block (encoder/slicetype-cl.c:100-127): order the context's stream after every
 * search queued by x264cu_lookahead_search_batch, so that work (or a timer event) queued next sees them finished */
int x264cu_lookahead_join( x264cu_lookahead_t *la );

/* read back per-MB results of one slot (the arrays x264_frame_t holds, common/frame.h:97-141) */
int x264cu_lookahead_get_mvs( x264cu_lookahead_t *la, int slot, int list, int dist_minus1, int16_t *h_mvs, int32_t *h_mv_costs );
int x264cu_lookahead_get_costs( x264cu_lookahead_t *la, int slot, int b_minus_p0, int p1_minus_b, uint16_t *h_lowres_costs );
int x264cu_lookahead_get_intra( x264cu_lookahead_t *la, int slot, int32_t *h_intra_costs );
int x264cu_lookahead_get_row_satds( x264cu_lookahead_t *la, int slot, int b_minus_p0, int p1_minus_b, int32_t *h_rows );
/* cost_est / cost_est_aq / intra_mbs as memoised on the host side (i_cost_est[18][18] etc.); -1 = not computed */
int x264cu_lookahead_get_cost_est( x264cu_lookahead_t *la, int slot, int b_minus_p0, int p1_minus_b,
                                   int *cost_est, int *cost_est_aq, int *intra_mbs );
/* the luma weight x264_weights_analyse left in fenc->weight[0][0] (slicetype.c:284-501, common/mc.h:30-46) when this
 * slot's picture was last costed as a P frame: out[4] = { enabled, i_scale, i_denom, i_offset } */
int x264cu_lookahead_get_weight( x264cu_lookahead_t *la, int slot, int *out4 );
int x264cu_lookahead_get_lowres_plane( x264cu_lookahead_t *la, int slot, int plane, uint8_t *h_out, intptr_t *stride );

/* ---- MB-tree (encoder/slicetype.c:1029-1184): propagation of each macroblock's "how much later pictures depend on it" back
 * along the lowres vectors, and the quantiser offsets it yields.  Device twins of macroblock_tree_propagate
 * (mbtree_propagate_cost / mbtree_propagate_list, common/mc.c:511-598) and macroblock_tree_finish; the sequence of calls is
 * macroblock_tree's own (x264cu_slicetype_step issues it when mb_tree is set).  i_propagate_cost is integer and matches the
 * reference exactly; f_qp_offset is float, evaluated operation by operation in the reference's order. ---- */
/* fenc->f_qp_offset_aq (and the initial f_qp_offset) of the picture in `slot`: the AQ offsets of x264_adaptive_quant_frame
 * (ratecontrol.c:225-420), NULL = all zero (AQ off).  Call after frame_put. */
int x264cu_lookahead_frame_set_qp_offset_aq( x264cu_lookahead_t *la, int slot, const float *h_qp_offset_aq );
/* memset( frame->i_propagate_cost, 0, ... ) */
int x264cu_lookahead_mbtree_reset( x264cu_lookahead_t *la, int slot );
/* XCHG( uint16_t*, a->i_propagate_cost, b->i_propagate_cost ) of the lookahead-less tree (slicetype.c:1125, :1177) */
int x264cu_lookahead_mbtree_swap( x264cu_lookahead_t *la, int slot_a, int slot_b );
/* macroblock_tree_propagate( h, frames, average_duration, p0, p1, b, referenced ); the cost (p0,p1,b) must have been requested.
 * fps_factor = CLIP_DURATION(frames[b]->f_duration) / (CLIP_DURATION(average_duration) * 256) * MBTREE_PRECISION (slicetype.c:1063) */
int x264cu_lookahead_mbtree_propagate( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int referenced, float fps_factor );
/* macroblock_tree_finish( h, frame, average_duration, ref0_distance ): fps_factor = round( CLIP_DURATION(average_duration) /
 * CLIP_DURATION(frame->f_duration) * 256 / MBTREE_PRECISION ), strength = 5 * (1 - rc.f_qcompress) (slicetype.c:1031-1038);
 * fps_factor = 0: f_qp_offset = f_qp_offset_aq (the lookahead-less intra case, slicetype.c:1121-1123) */
int x264cu_lookahead_mbtree_finish( x264cu_lookahead_t *la, int slot, int fps_factor, int ref0_distance, float strength );
/* slicetype_frame_cost_recalculate( h, frames, p0, p1, b ) (slicetype.c:999-1024), used by x264_rc_analyse_slice and the VBV
 * lookahead: the cost of a requested (slot = frames[b], dist0 = b-p0, dist1 = p1-b) with every macroblock's lowres cost scaled by
 * x264_exp2fix8 of its quantiser offset -- f_qp_offset (MB-tree's), or f_qp_offset_aq when b_type (a B picture).  Rewrites
 * i_row_satds[dist0][dist1] on the device; *h_score = the returned cost, h_row_satd (mb_height ints, may be NULL) = the rows. */
int x264cu_lookahead_frame_cost_recalculate( x264cu_lookahead_t *la, int slot, int dist0, int dist1, int b_type, int *h_score, int32_t *h_row_satd );
int x264cu_lookahead_get_qp_offset( x264cu_lookahead_t *la, int slot, float *h_qp_offset );          /* f_qp_offset, one per MB */
int x264cu_lookahead_get_propagate_cost( x264cu_lookahead_t *la, int slot, uint16_t *h_propagate_cost );
/* f_weighted_cost_delta[dist_minus1] (slicetype.c:462-463; set only with weighted_pred < 0 = X264_WEIGHTP_FAKE) */
float x264cu_lookahead_get_weighted_cost_delta( x264cu_lookahead_t *la, int slot, int dist_minus1 );

/* ------------------------------------------------------------------------------------------------
 * Slice-type decision on top of the GPU lookahead: the host control flow of x264_slicetype_decide /
 * x264_slicetype_analyse / scenecut / slicetype_path (encoder/slicetype.c:1288-1974) and the frame queue of
 * encoder/lookahead.c:192-250 (synchronous lookahead: param.i_sync_lookahead == 0), calling
 * x264cu_lookahead_frame_cost exactly where the reference calls slicetype_frame_cost -- MB-tree's and the
 * rate control's cost requests included, because the memoised B costs depend on request order
 * (slicetype.c:629-642).  Plain C; no device code.
 * Not covered: 2-pass statistics, blu-ray compatible open-GOP, the planned cpb durations of the VBV plan and the intra-refresh
 * column correction of x264_rc_analyse_slice (la.vbv together with intra_refresh is rejected at open).
 * ---------------------------------------------------------------------------------------------- */
typedef struct x264cu_slicetype x264cu_slicetype_t;

typedef struct
{
    x264cu_lookahead_params_t la; /* la.n_slots is ignored (sized internally) */
    int keyint_max, keyint_min;   /* h->param.i_keyint_max / i_keyint_min (after validation) */
    int scenecut_threshold;       /* h->param.i_scenecut_threshold */
    int b_adapt;                  /* h->param.i_bframe_adaptive: 0 none, 1 fast, 2 trellis */
    int b_pyramid;                /* h->param.i_bframe_pyramid: 0 none, 1 strict, 2 normal */
    int rc_lookahead;             /* h->param.rc.i_lookahead */
    int psy;                      /* h->param.analyse.b_psy */
    int frame_reference;          /* h->param.i_frame_reference */
    int rc_cqp;                   /* h->param.rc.i_rc_method == X264_RC_CQP */
    int fps_num, fps_den;         /* h->param.i_fps_num / i_fps_den (constant frame rate); 0 = 25/1.  MB-tree's duration factors */
    float qcompress;              /* h->param.rc.f_qcompress (0..1; 0 is legal: MB-tree strength = 5 * (1 - qcompress)); negative = the default 0.6 */
    float aq_strength;            /* h->param.rc.f_aq_strength; 0 switches adaptive quantisation off as in the reference (encoder.c:1094-1097);
                                     negative = the default 1.0.  Used by x264cu_slicetype_step_i420 (la.aq_mode = the mode, 0..3) */
    int open_gop;                 /* h->param.b_open_gop (without b_bluray_compat) */
    int intra_refresh;            /* h->param.b_intra_refresh: no keyframes but the first, scene cuts become I pictures */
} x264cu_slicetype_params_t;

enum { X264CU_TYPE_AUTO = 0, X264CU_TYPE_IDR = 1, X264CU_TYPE_I = 2, X264CU_TYPE_P = 3, X264CU_TYPE_BREF = 4,
       X264CU_TYPE_B = 5, X264CU_TYPE_KEYFRAME = 6 };       /* x264.h:274-281 */

int  x264cu_slicetype_open( x264cu_ctx_t *ctx, const x264cu_slicetype_params_t *params, x264cu_slicetype_t **out );
void x264cu_slicetype_close( x264cu_slicetype_t *st );
/* One x264_encoder_encode step as far as the lookahead is concerned (encoder.c:3323-3450): h_luma != NULL queues a
 * picture (x264_lookahead_put_frame), NULL flushes.  When the encoder would have picked a frame to encode,
 * *out_frame is its display index and *out_type its decided type (coded order); otherwise *out_frame = -1.
 * Returns 0, or -1 on error.  While flushing, *out_frame == -1 means the stream is drained. */
int  x264cu_slicetype_step( x264cu_slicetype_t *st, const uint8_t *h_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                            int *out_frame, int *out_type );
/* same from an I420 picture: adaptive quantisation (la.aq_mode, aq_strength) runs on the device too, so that the whole of
 * x264_encoder_encode's preparation for the lookahead -- x264_adaptive_quant_frame, x264_frame_init_lowres, put_frame -- is here */
int  x264cu_slicetype_step_i420( x264cu_slicetype_t *st, const uint8_t *h_luma, intptr_t luma_stride, const uint8_t *h_cb, const uint8_t *h_cr,
                                 intptr_t chroma_stride, int *out_frame, int *out_type );
/* same with the picture already in HBM (8-byte aligned base and stride) */
int  x264cu_slicetype_step_device( x264cu_slicetype_t *st, const uint8_t *d_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                                   int *out_frame, int *out_type );
/* prefetch = 1 (default): when a picture is queued, its lowres searches against the previous bframes+1 pictures are
 * launched at once on a second stream (the x264_opencl_slicetype_prep idea, encoder/slicetype-cl.c:653); 0: every search
 * runs on demand inside the cost request that needs it.  The decisions are identical either way. */
void x264cu_slicetype_set_prefetch( x264cu_slicetype_t *st, int prefetch );
/* pictures queued beyond the lookahead before a decision is taken (0..64; default 24 for lookaheads >= 12, else 0; must be
 * set before the first picture).  It is the synchronous twin of param.i_sync_lookahead (encoder.c:1137-1141, :1611): the
 * analysis never looks at more than i_slicetype_length+1 pictures (b_deterministic, slicetype.c:1480-1485), so the
 * decisions do not change -- only the searches of the newest pictures get time to finish on the second stream. */
void x264cu_slicetype_set_run_ahead( x264cu_slicetype_t *st, int pictures );
/* speculate = 1: with every prefetch launch, all cost requests the decision can make about the new pictures are computed too
 * (x264cu_lookahead_finalize_batch) -- a request then costs the calling thread a table lookup instead of a launch, a copy and a
 * wait.  0: every request is computed when it is made.  Default: on for a sharded stream (where it is what splits the cost
 * requests between the GPUs) and, on a single GPU, for the trellis over a B pyramid (b_adapt 2, bframes > 1: a new GOP's first
 * analysis asks for thousands of triples at once; +12 % at 8K / bframes 16); off otherwise (b_adapt 1's short windows: the
 * searches' throughput is the bound and the extra triples cost 5 %).  The decisions are identical either way. */
void x264cu_slicetype_set_speculation( x264cu_slicetype_t *st, int speculate );
/* pictures whose searches are gathered into one prefetch launch (1..32; default 12 for lookaheads >= 12, else 1; before the first
 * picture).  A launch needs several dozen independent searches to fill the GPU; the decisions do not depend on it. */
void x264cu_slicetype_set_prefetch_group( x264cu_slicetype_t *st, int pictures );
/* pic_in->i_type of the NEXT picture queued with x264cu_slicetype_step* (forced frame types: a qpfile, an application's keyframe
 * request; X264CU_TYPE_KEYFRAME = IDR, or I with open-GOP); AUTO again afterwards.  The analysis respects it exactly as
 * x264_slicetype_analyse / x264_slicetype_decide do (i_forced_type, slicetype.c:1534-1539, :1656, :1690-1738, :1803-1828). */
int  x264cu_slicetype_set_next_type( x264cu_slicetype_t *st, int type );
/* x264cu_lookahead_set_async_upload for the lookahead underneath: page-locked pictures are read in place and must stay
 * unmodified until `on` more pictures have been queued (1 means 4) */
void x264cu_slicetype_set_async_upload( x264cu_slicetype_t *st, int on );
/* Sharded stream: this process is rank `rank` of `world` GPUs that are all fed the SAME pictures.  Every rank runs the same
 * decisions; of each prefetch group it searches only the pairs whose searched picture has display index % world == rank, and
 * the groups' results are exchanged with one all-gather each, one group late so that nobody waits for a search.  The callback
 * does the collective (the caller owns the communicator: torch.distributed / NCCL):
 *   phase 0: provide two device buffers, *d_send of bytes_per_rank and *d_recv of world * bytes_per_rank bytes
 *   phase 1: all-gather d_send of every rank into d_recv (rank r's block at r * bytes_per_rank), enqueued on `stream`
 *   phases 2 / 3: the same for a second, independent pair of buffers (the exchange of cost-request results)
 * and returns 0 / -1.  Frame types and MB-tree offsets are those of a single GPU (tested). */
typedef int (*x264cu_exchange_fn)( void *user, int phase, size_t bytes_per_rank, void **d_send, void **d_recv, void *stream );
int  x264cu_slicetype_set_shard( x264cu_slicetype_t *st, int rank, int world, x264cu_exchange_fn fn, void *user );
/* x264cu_lookahead_finalize_batch with the triples split between the GPUs of a sharded stream: triple i is computed by rank
 * owner[i] (x264cu_slicetype uses the display index of its picture b, modulo world -- the rank that also ran b's searches), and
 * every rank receives every result -- the 32-byte record and lowres_costs[b-p0][p1-b] of each triple -- in ONE all-gather per
 * window: the same callback with phases 2 (buffers) and 3 (collective), so that it can keep this exchange's buffers apart from
 * the search exchange's.  All ranks must call it with the same list. */
int  x264cu_lookahead_finalize_batch_sharded( x264cu_lookahead_t *la, int n, const int *b_slot, const int *p0_slot, const int *p1_slot,
                                              const int *d0, const int *d1, const int *owner, int rank, int world,
                                              x264cu_exchange_fn exchange, void *user );
/* A ready-made exchange for hosts written in C: ncclAllGather over NVLink / NVSwitch.  libnccl.so.2 is loaded at run time (dlopen;
 * X264CU_NCCL_LIB names another file), so linking libx264_b200 never pulls NCCL in.  One process per GPU:
 *   x264cu_nccl_open: rank 0 creates the ncclUniqueId and publishes it in the file `id_path` (any path all ranks can read, e.g. on
 *                     /dev/shm); the other ranks wait for it (timeout_s, 0 = 60) -- or
 *   x264cu_nccl_wrap: use a communicator (ncclComm_t) the application already has;
 * then  x264cu_slicetype_set_shard( st, rank, world, x264cu_exchange_nccl, nc ). */
typedef struct x264cu_nccl x264cu_nccl_t;
int  x264cu_nccl_open( x264cu_ctx_t *ctx, int rank, int world, const char *id_path, int timeout_s, x264cu_nccl_t **out );
int  x264cu_nccl_wrap( x264cu_ctx_t *ctx, void *nccl_comm, int rank, int world, x264cu_nccl_t **out );
void x264cu_nccl_close( x264cu_nccl_t *nc );
int  x264cu_exchange_nccl( void *user, int phase, size_t bytes_per_rank, void **d_send, void **d_recv, void *stream );
long x264cu_nccl_calls( x264cu_nccl_t *nc, unsigned long long *bytes );      /* exchanges so far, bytes gathered */

/* the lookahead object underneath (for reading per-MB results) and the slot a display index currently occupies (-1 if gone) */
x264cu_lookahead_t *x264cu_slicetype_lookahead( x264cu_slicetype_t *st );
int  x264cu_slicetype_slot_of( x264cu_slicetype_t *st, int frame );
/* f_qp_offset of a picture still held by the lookahead (a frame just returned by x264cu_slicetype_step is, until the next
 * call): MB-tree's quantiser offsets when mb_tree is set.  mb_count floats.  Non-B pictures only: B pictures are coded with
 * f_qp_offset_aq (slicetype.c:1003), which the caller supplied, and have left the lookahead by then (-1). */
int  x264cu_slicetype_get_qp_offset( x264cu_slicetype_t *st, int frame, float *h_qp_offset );
/* x264_rc_analyse_slice (slicetype.c:1976-2030) for the picture the last x264cu_slicetype_step returned (it stays in the
 * lookahead until the next step): *cost = fdec->i_satd -- the memoised cost of the picture against the references it will be
 * coded with, recalculated with MB-tree's quantiser offsets when mb_tree is set (slicetype_frame_cost_recalculate), the AQ
 * weighted cost otherwise when aq_mode is set --, h_row_satd (mb_height ints, may be NULL) = fdec->i_row_satd, h_row_satd_intra
 * (may be NULL; non-I pictures) = i_row_satds[0][0].  The rows are defined with la.vbv or mb_tree, as in the reference.  B
 * pictures need la.vbv (slicetype.c:1916); rc_cqp: -1 (the reference does not call it).  Not built: the intra-refresh column
 * correction (slicetype.c:2015-2036; la.vbv with intra_refresh is rejected at open). */
int  x264cu_slicetype_rc_analyse_slice( x264cu_slicetype_t *st, int frame, int *cost, int *h_row_satd, int *h_row_satd_intra );
/* VBV lookahead (vbv_lookahead, slicetype.c:1225-1286; la.vbv and rc_lookahead > 0): i_planned_type[] / i_planned_satd[] of a
 * non-B picture the last step returned = the types and costs of the pictures coded after it, as far as the lookahead has
 * decided them.  Returns the number of entries (the reference terminates the list with X264_TYPE_AUTO), -1 on error.  The
 * f_planned_cpb_duration[] of the reference (picture durations, no analysis) is not produced. */
int  x264cu_slicetype_get_planned( x264cu_slicetype_t *st, int frame, int *h_type, int *h_satd, int max_entries );
/* number of slicetype_frame_cost requests issued so far (memo hits included) */
long x264cu_slicetype_cost_requests( x264cu_slicetype_t *st );
/* the largest distance p1-b any of those requests named for a B picture (p0 < b < p1).  The prefetcher launches list-1 searches up
 * to half a mini-GOP under a B pyramid (nobody asks further: tests/test_slicetype_host.py sweeps it); farther ones run on demand. */
int x264cu_slicetype_farthest_list1( x264cu_slicetype_t *st );

/* ------------------------------------------------------------------------------------------------
 * Batched twin of x264_me_search_ref + refine_subpel (encoder/me.h:58-60, encoder/me.c:182-992): one job = one call.
 * DIA / HEX / UMH / ESA / TESA (the exhaustive searches of me.c:618-771, with their successive-
 * elimination prefilter where it decides the result), every partition size and sub-pel level.  One warp runs one search with the
 * reference's control flow; jobs are independent (their predictors are inputs), which is how the full-resolution
 * motion-estimation stage is replayed from recorded x264_me_t inputs (BASELINE config 3).  This entry is luma only and takes one
 * reference picture and one lambda per call; x264cu_me_search_frame below takes a whole picture's searches, chroma ME included.
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    int32_t  i_pixel;                /* m->i_pixel, PIXEL_16x16 .. PIXEL_4x4 */
    uint32_t fenc_off;               /* byte offset of the block in the fenc plane */
    uint32_t ref_off;                /* byte offset of the co-located block in each reference plane */
    int16_t  mvp[2];                 /* m->mvp */
    int16_t  mvc[9][2];              /* candidate predictors (mvc argument): up to 9 in B slices (x264_mb_predict_mv_ref16x16, common/mvpred.c:519-610) */
    int32_t  i_mvc;                  /* 0..9 */
    int16_t  mv_min_spel[2], mv_max_spel[2];   /* h->mb.mv_min_spel / mv_max_spel; mv_limit_fpel follows from them and params.fpel_border */
    int32_t  halfpel_thresh;         /* *p_halfpel_thresh, or -1 for NULL */
} x264cu_me_job_t;

typedef struct
{
    int16_t mv[2];                   /* m->mv */
    int32_t cost;                    /* m->cost */
    int32_t cost_mv;                 /* m->cost_mv (undefined after a half-pel early exit, as in the reference) */
    int32_t halfpel_thresh;          /* updated *p_halfpel_thresh */
} x264cu_me_result_t;

typedef struct
{
    int me_method;                   /* h->mb.i_me_method (X264CU_ME_DIA .. X264CU_ME_TESA; esa: me_range <= 120, tesa: <= 64) */
    int subpel_refine;               /* h->mb.i_subpel_refine */
    int me_range;                    /* h->param.analyse.i_me_range */
    int mbcmp_satd;                  /* encoder.c:1409-1427: mbcmp is SATD iff the encoder's subme > 1; under TESA fpelcmp follows it */
    int lambda;                      /* a->i_lambda: cost_mv = lambda * bits (analyse.c:143-157) */
    int mv_range;                    /* h->param.analyse.i_mv_range: sizes the cost table */
    int weight_enabled, weight_scale, weight_denom, weight_offset;   /* m->weight[0] (common/mc.h:235-245) */
    int fpel_border;                 /* h->mb.mv_limit_fpel = (mv_min_spel >> 2) + fpel_border .. (mv_max_spel >> 2) - fpel_border: 6 in the
                                        encoder's analysis (i_fpel_border, analyse.c:333-349), 0 in the lookahead (slicetype.c:550-562) */
} x264cu_me_params_t;

/* d_fref[4]: F,H,V,C plane origins of the reference (x264cu_hpel_filter output), d_fref_w: the weighted full-pel plane or
 * d_fref[0].  All planes share ref_stride and are padded by X264CU_PAD. */
int x264cu_me_search_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *params,
                            const uint8_t *d_fenc, intptr_t fenc_stride,
                            const uint8_t *const d_fref[4], const uint8_t *d_fref_w, intptr_t ref_stride,
                            const x264cu_me_job_t *d_jobs, int n, x264cu_me_result_t *d_results );

/* The motion searches of one coded picture in ONE launch (BASELINE config 3: the x264_me_t stream x264_mb_analyse_inter_* feeds
 * to x264_me_search_ref, encoder/analyse.c:1287-1938): every job names the reference picture it searches -- any entry of either
 * list, weighted duplicates included (h->fref[list][i_ref], h->sh.weight[i_ref][0..2]) -- and its lambda (a->i_lambda follows the
 * macroblock's quantiser: AQ, MB-tree, VBV).  With chroma_me (h->mb.b_chroma_me, common/macroblock.c:507: P slices at subme >= 5,
 * B slices at subme >= 9) the quarter-pel refinement of partitions >= 8x8 adds the two chroma planes' costs (COST_MV_SATD's chroma
 * branch, me.c:826-857: mc_chroma, common/mc.c:251-283, the planes' explicit weights, mbcmp of the half-size block; 4:2:0,
 * progressive).  Plane pointers are pixel (0,0) of each plane; fenc_off / ref_off of the jobs are byte offsets from there (the
 * chroma origin of a block is derived from them: row y/2, byte 2*(x/2) of the interleaved plane). */
typedef struct
{
    const uint8_t *d_fref[4];        /* fref->filtered[0][0..3]: F,H,V,C (x264cu_hpel_filter output) */
    const uint8_t *d_fref_w;         /* the weighted full-pel plane (fref->weighted[0]) or NULL = d_fref[0] */
    const uint8_t *d_fref_uv;        /* fref->plane[1]: NV12 chroma (padded), or NULL without chroma ME */
    int weight[3][4];                /* m->weight[0..2] = { enabled, i_scale, i_denom, i_offset } (common/mc.h:235-245) */
} x264cu_me_ref_t;

typedef struct
{
    const uint8_t *d_fenc; intptr_t fenc_stride;          /* fenc->plane[0] */
    const uint8_t *d_fenc_uv; intptr_t fenc_uv_stride;    /* fenc->plane[1] (NV12), or NULL without chroma ME */
    intptr_t ref_stride, ref_uv_stride;                   /* shared by all references */
    const x264cu_me_ref_t *refs; int n_refs;              /* host array, <= 64 */
    const int *lambdas; int n_lambdas;                    /* host array of the distinct lambdas in use, <= 128 */
    int chroma_me;                                        /* h->mb.b_chroma_me */
} x264cu_me_frame_t;

typedef struct
{
    x264cu_me_job_t job;
    int16_t i_ref;                   /* index into x264cu_me_frame_t.refs */
    int16_t i_lambda;                /* index into x264cu_me_frame_t.lambdas */
} x264cu_me_frame_job_t;

/* params: me_method, subpel_refine, me_range, mbcmp_satd and mv_range are read (lambda and the weight come with each job) */
int x264cu_me_search_frame( x264cu_ctx_t *ctx, const x264cu_me_params_t *params, const x264cu_me_frame_t *frame,
                            const x264cu_me_frame_job_t *d_jobs, int n, x264cu_me_result_t *d_results );

/* Batched twin of x264_me_refine_qpel (refdupe = 0; encoder/me.h:59, me.c:800-809) and x264_me_refine_qpel_refdupe (refdupe = 1;
 * me.h:60, me.c:811-814): the sub-pel refinement continued from a stored vector and cost.  params: subpel_refine =
 * h->mb.i_subpel_refine, mbcmp_satd, lambda, mv_range and the weight are read.  Results as x264cu_me_result_t (halfpel_thresh is
 * the updated *p_halfpel_thresh of the refdupe form, -1 otherwise). */
typedef struct
{
    int32_t  i_pixel;                /* m->i_pixel */
    uint32_t fenc_off, ref_off;      /* byte offsets of the block in the fenc plane / of the co-located block in the reference planes */
    int16_t  mvp[2];                 /* m->mvp */
    int16_t  mv[2];                  /* m->mv on entry */
    int32_t  cost;                   /* m->cost on entry */
    int32_t  i_ref_cost;             /* m->i_ref_cost: taken off the cost of 8x8 and larger partitions by x264_me_refine_qpel */
    int16_t  mv_min_spel[2], mv_max_spel[2];
    int32_t  halfpel_thresh;         /* refdupe: *p_halfpel_thresh, or -1 for NULL */
} x264cu_me_refine_job_t;

int x264cu_me_refine_qpel_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *params, int refdupe,
                                 const uint8_t *d_fenc, intptr_t fenc_stride, const uint8_t *const d_fref[4], intptr_t ref_stride,
                                 const x264cu_me_refine_job_t *d_jobs, int n, x264cu_me_result_t *d_results );

/* Batched twin of x264_me_refine_bidir_satd (encoder/me.h:63, encoder/me.c:1027-1183): joint +-1 quarter-pel refinement of
 * the two vectors of a bi-predicted partition, one job = one call.  params: mbcmp_satd, lambda and mv_range are read. */
typedef struct
{
    int32_t  i_pixel;                /* m0->i_pixel */
    uint32_t fenc_off;               /* byte offset of the block in the fenc plane */
    uint32_t ref0_off, ref1_off;     /* byte offsets of the co-located block in the list-0 / list-1 reference planes */
    int16_t  mv[4];                  /* m0->mv[0..1], m1->mv[0..1] */
    int16_t  mvp[4];                 /* m0->mvp, m1->mvp */
    int16_t  mv_min_spel[2], mv_max_spel[2];   /* h->mb.mv_min_spel / mv_max_spel */
    int32_t  i_weight;               /* bipred weight of list 0 (32 = plain average; pixel_avg_weight_wxh otherwise) */
} x264cu_bidir_job_t;

typedef struct
{
    int16_t mv[4];                   /* refined m0->mv, m1->mv (the inputs when the pair is within 8 of the window edge) */
    int32_t cost;                    /* best mbcmp + mv costs seen (not an output of the reference; COST_MAX = 1<<28 on early return) */
} x264cu_bidir_result_t;

int x264cu_me_refine_bidir_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *params,
                                  const uint8_t *d_fenc, intptr_t fenc_stride,
                                  const uint8_t *const d_fref0[4], const uint8_t *const d_fref1[4], intptr_t ref_stride,
                                  const x264cu_bidir_job_t *d_jobs, int n, x264cu_bidir_result_t *d_results );

#ifdef __cplusplus
}
#endif
#endif
