/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_pixel.c header).  CPU restatement, in plain C, of the
 * reference's ME / lookahead cost path (jpsdr/x264 core 165, 8-bit).  Every function cites the reference
 * file:line it follows.  Parity status: pinned against oracle/_ref (the compiled, unmodified reference).
 */
#ifndef X264_B200_ORACLE_H
#define X264_B200_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same indices as common/pixel.h:37-52 */
enum { ORC_PIXEL_16x16, ORC_PIXEL_16x8, ORC_PIXEL_8x16, ORC_PIXEL_8x8, ORC_PIXEL_8x4, ORC_PIXEL_4x8,
       ORC_PIXEL_4x4, ORC_PIXEL_4x16, ORC_PIXEL_NB };
enum { ORC_SAD, ORC_SSD, ORC_SATD, ORC_SA8D };
enum { ORC_ME_DIA, ORC_ME_HEX, ORC_ME_UMH, ORC_ME_ESA, ORC_ME_TESA };   /* x264.h X264_ME_DIA/HEX/UMH/ESA/TESA */
#define ORC_COST_MAX (1<<28)                             /* encoder/me.h:30 */
#define ORC_PAD 32                                       /* common/frame.h:32-33 PADH/PADV */

extern const int orc_pixel_w[ORC_PIXEL_NB], orc_pixel_h[ORC_PIXEL_NB];

typedef struct { uint32_t fenc_off, ref_off; } orc_cand_t;

/* ---- pixel metrics (oracle_pixel.c) ---- */
int  orc_sad ( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h );
int  orc_ssd ( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h );
int  orc_satd( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h );
int  orc_sa8d( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h );
int  orc_pixel_cmp( int metric, int i_pixel, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb );
void orc_pixel_cmp_batch( int metric, int i_pixel, const uint8_t *fenc, intptr_t fenc_stride,
                          const uint8_t *ref, intptr_t ref_stride, const orc_cand_t *cand, int n, int32_t *out );
void orc_pixel_cmp_mvfield( int metric, int i_pixel, const uint8_t *fenc, intptr_t fenc_stride,
                            const uint8_t *ref, intptr_t ref_stride, int blocks_x, int blocks_y, int k_cands,
                            const int16_t *mv, int32_t *out );

/* ---- motion compensation / frame preparation (oracle_mc.c) ---- */
/* weight: scale/denom/offset as in x264_weight_t (common/mc.h:235-245); enabled==0 means "no weightfn" */
typedef struct { int enabled, scale, denom, offset; } orc_weight_t;

void orc_plane_expand_border( uint8_t *pix, intptr_t stride, int width, int height, int padh, int padv );
void orc_frame_init_lowres( const uint8_t *src, intptr_t src_stride, int width, int height,
                            uint8_t *lowres[4], intptr_t dst_stride, int width_lowres, int lines_lowres );
void orc_hpel_filter_plane( const uint8_t *src, intptr_t stride, int width, int height,
                            uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, intptr_t dst_stride, int pad );
void orc_mc_luma( uint8_t *dst, intptr_t dst_stride, const uint8_t *const src[4], intptr_t src_stride,
                  int mvx, int mvy, int w, int h, const orc_weight_t *wt );
void orc_pixel_avg( uint8_t *dst, intptr_t sd, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb,
                    int w, int h, int weight );
void orc_mc_weight( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, const orc_weight_t *wt, int w, int h );
void orc_weight_scale_plane( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, int w, int h, const orc_weight_t *wt );
/* mc_chroma (common/mc.c:251-283) for one component (0 = U, 1 = V) of an NV12 plane; w x h chroma pixels */
void orc_mc_chroma( uint8_t *dst, intptr_t dst_stride, const uint8_t *src_uv, intptr_t src_stride,
                    int mvx, int mvy, int w, int h, int comp );
/* successive elimination: pixf.ads[] (common/pixel.c:759-803), integral planes (common/mc.c:424-456, :748-783) */
int  orc_pixel_ads( int k, const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx, int16_t *mvs, int width, int thresh );
void orc_integral_init( const uint8_t *plane, intptr_t stride, int width, int height, int pad, uint16_t *sum8, uint16_t *sum4 );
/* input staging: common/mc.c:294-339 and x264_frame_copy_picture (common/frame.c:363-480), 8-bit 4:2:0 colour spaces */
void orc_plane_copy_interleave( uint8_t *dst, intptr_t sd, const uint8_t *u, intptr_t su, const uint8_t *v, intptr_t sv, int w, int h );
void orc_plane_copy_deinterleave( uint8_t *a, intptr_t sa, uint8_t *b, intptr_t sb, const uint8_t *src, intptr_t ss, int w, int h );
void orc_plane_copy_swap( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, int w, int h );
int  orc_frame_copy_picture( int i_csp, const uint8_t *const plane[3], const int stride[3], int w, int h,
                             uint8_t *luma, intptr_t sl, uint8_t *chroma, intptr_t sc );
/* table[0 .. 2*len] with the zero-mvd entry at table[len]; len = 2*4*mv_range */
void orc_cost_mv_table( uint16_t *table, int len, int lambda );

/* ---- motion search (oracle_me.c) ---- */
typedef struct
{
    /* search configuration (x264_t fields read by me.c) */
    int me_method;          /* h->mb.i_me_method */
    int subpel_refine;      /* h->mb.i_subpel_refine */
    int me_range;           /* h->param.analyse.i_me_range */
    int mbcmp_is_satd;      /* encoder.c:1409-1427: mbcmp = SATD iff param subme > 1 */
    int mv_min_spel[2], mv_max_spel[2];     /* h->mb.mv_min_spel / mv_max_spel */
    int mv_limit_fpel[2][2];                /* h->mb.mv_limit_fpel */
    int chroma_me;          /* h->mb.b_chroma_me (common/macroblock.c:507); 4:2:0, progressive */
} orc_me_ctx_t;

typedef struct
{
    int i_pixel;
    const uint16_t *p_cost_mv;     /* centred table: p_cost_mv[mvd] */
    const uint8_t *p_fref[4];      /* F,H,V,C planes at the block origin */
    const uint8_t *p_fref_w;       /* weighted full-pel plane (== p_fref[0] when unweighted) */
    const uint8_t *p_fenc;         /* block origin, stride fenc_stride */
    intptr_t fenc_stride;
    intptr_t stride;
    orc_weight_t weight;
    int16_t mvp[2];
    /* out */
    int cost_mv, cost;
    int16_t mv[2];
    /* chroma ME (c->chroma_me, partitions >= 8x8; me.c:826-857): NV12 planes at the block's chroma origin */
    const uint8_t *p_fref_uv;      /* m->p_fref[4] */
    intptr_t stride_uv;            /* m->i_stride[1] */
    const uint8_t *p_fenc_uv;      /* the source picture's interleaved chroma (the reference reads de-interleaved copies, p_fenc[1..2]) */
    intptr_t fenc_uv_stride;
    orc_weight_t weight_uv[2];     /* m->weight[1], m->weight[2] */
} orc_me_t;

void orc_me_search_ref( const orc_me_ctx_t *c, orc_me_t *m, const int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh );
/* x264_me_refine_qpel (mode 0) / x264_me_refine_qpel_refdupe (mode 1), me.c:800-814: continue from m->mv / m->cost */
void orc_me_refine_qpel( const orc_me_ctx_t *c, orc_me_t *m, int mode, int i_ref_cost, int *p_halfpel_thresh );
/* x264_me_refine_bidir_satd (me.c:1027-1183): refines m0->mv / m1->mv jointly; reads c->mv_min_spel / mv_max_spel / mbcmp_is_satd */
void orc_me_refine_bidir_satd( const orc_me_ctx_t *c, orc_me_t *m0, orc_me_t *m1, int i_weight );

/* ---- lowres lookahead (oracle_lookahead.c) ---- */
typedef struct
{
    int width, height;                 /* full-res luma size as given by the user (h->param.i_width/height) */
    int mb_width, mb_height;           /* ceil/16 */
    int subpel_refine;                 /* h->param.analyse.i_subpel_refine */
    int me_method;                     /* h->param.analyse.i_me_method */
    int me_range;                      /* h->param.analyse.i_me_range */
    int mv_range;                      /* h->param.analyse.i_mv_range */
    int bframes;                       /* h->param.i_bframe */
    int bframe_bias;                   /* h->param.i_bframe_bias */
    int weighted_bipred;               /* h->param.analyse.b_weighted_bipred */
    int aq_mode;                       /* h->param.rc.i_aq_mode != 0 */
    int do_edges;                      /* mbtree || vbv || tiny frame (slicetype.c:823) */
    int vbv;                           /* row satds requested */
    int weighted_pred;                 /* h->param.analyse.i_weighted_pred != 0 (incl. X264_WEIGHTP_FAKE, encoder.c:1316) */
} orc_la_params_t;

typedef struct orc_la_frame orc_la_frame_t;
struct orc_la_frame
{
    uint8_t *lowres_buf[4];            /* allocations */
    uint8_t *lowres[4];                /* plane origins */
    intptr_t stride_lowres;
    int width_lowres, lines_lowres;
    int mb_count, bframes;
    int16_t (*lowres_mvs[2][17])[2];
    int     *lowres_mv_costs[2][17];
    uint16_t *lowres_costs[18][18];
    int     *row_satds[18][18];
    int     *intra_cost;
    uint16_t *inv_qscale_factor;
    int cost_est[18][18], cost_est_aq[18][18];
    int intra_mbs[18];
    int b_intra_calculated;
    uint64_t pixel_sum, pixel_ssd;     /* i_pixel_sum[0] / i_pixel_ssd[0] as left by x264_adaptive_quant_frame (ratecontrol.c:304-415) */
    int i_frame;
    orc_weight_t weight;               /* fenc->weight[0][0] after the last lookahead analysis */
    uint8_t *weighted_buf;             /* fenc->weighted[0]: weighted copy of the reference's padded F plane */
    /* MB-tree (slicetype.c:1029-1184) */
    uint16_t *propagate_cost;          /* i_propagate_cost */
    float *qp_offset, *qp_offset_aq;   /* f_qp_offset / f_qp_offset_aq */
    float weighted_cost_delta[17];     /* f_weighted_cost_delta (slicetype.c:462-463: X264_WEIGHTP_FAKE only = weighted_pred < 0) */
    int mb_width;
};

orc_la_frame_t *orc_la_frame_new( const orc_la_params_t *p, const uint8_t *luma, intptr_t luma_stride );
void orc_la_frame_delete( orc_la_frame_t *f );
int  orc_la_frame_cost( const orc_la_params_t *p, const uint16_t *cost_mv_centre,
                        orc_la_frame_t **frames, int p0, int p1, int b );
/* macroblock_tree_propagate (slicetype.c:1050-1089) with mbtree_propagate_cost / _list (common/mc.c:511-598);
 * fps_factor = CLIP_DURATION(f_duration) / (CLIP_DURATION(average_duration) * 256) * MBTREE_PRECISION */
void orc_la_mbtree_propagate( const orc_la_params_t *p, orc_la_frame_t **frames, int p0, int p1, int b, int referenced, float fps_factor );
/* macroblock_tree_finish (slicetype.c:1029-1048): fps_factor = round( CLIP(avg) / CLIP(dur) * 256 / MBTREE_PRECISION ),
 * strength = 5 * (1 - qcompress) */
void orc_la_mbtree_finish( orc_la_frame_t *f, int fps_factor, int ref0_distance, float strength );
void orc_la_mbtree_reset( orc_la_frame_t *f );
/* slicetype_frame_cost_recalculate (slicetype.c:999-1024) */
int  orc_la_frame_cost_recalculate( const orc_la_params_t *p, orc_la_frame_t **frames, int p0, int p1, int b, int b_is_b_type );
void orc_la_frame_set_qp_offset_aq( orc_la_frame_t *f, const float *aq );
void orc_la_frame_get_mbtree( orc_la_frame_t *f, int what, int i, void *out );
float orc_log2( uint32_t x );           /* x264_log2, common/base.h:226-230 */
float orc_log2_frac( uint32_t x );      /* its mantissa part alone: x264_log2_lut[(x<<lz>>24)&0x7f] */

/* ---- adaptive quantisation (oracle_aq.c): x264_adaptive_quant_frame, aq-mode 0 / 1 ---- */
void orc_adaptive_quant_frame( const uint8_t *luma, intptr_t stride, const uint8_t *cb, const uint8_t *cr, intptr_t cstride,
                               int width, int height, int aq_mode, float aq_strength, float *qp_offset_aq, uint16_t *inv_qscale,
                               uint64_t *stats );

#ifdef __cplusplus
}
#endif
#endif
