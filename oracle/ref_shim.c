/*
 * TEST INFRASTRUCTURE ONLY -- thin C-ABI exports around the UNMODIFIED reference (jpsdr/x264), compiled by
 * oracle/Makefile.ref straight from /root/reference into oracle/_ref/libx264ref.so.  This file contains no
 * reference code: it #includes the reference's encoder/analyse.c (which itself #includes slicetype.c) so that
 * the file-static slicetype_* functions can be driven directly, and calls the reference's own functions.
 * Used by tests/ to pin oracle/ and the CUDA path, and by bench.py --impl reference as the CPU arm.
 */
/* every x264_me_search_ref / x264_me_search call of the analysis goes through a recorder (xref_me_trace_*, below): the macro of
 * encoder/me.h is re-pointed before the reference's analyse.c is compiled; the recorder calls the real function */
#include "common/common.h"
#include "encoder/macroblock.h"
#include "encoder/me.h"
#undef x264_me_search_ref
#define x264_me_search_ref xref_traced_me_search_ref
static void xref_traced_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh );
#include "encoder/analyse.c"          /* brings in common/common.h, me.h, slicetype.c, rdo.c ... */
#undef x264_me_search_ref
#define x264_me_search_ref x264_template(me_search_ref)

#include <stdio.h>
#include <string.h>
#include <stdlib.h>

x264_t *x264_encoder_open( x264_param_t *, void * );
int     x264_encoder_encode( x264_t *, x264_nal_t **pp_nal, int *pi_nal, x264_picture_t *pic_in, x264_picture_t *pic_out );
void    x264_encoder_close( x264_t * );
int     x264_encoder_delayed_frames( x264_t * );

#define XREF_API __attribute__((visibility("default")))

XREF_API int xref_build( void ) { return X264_BUILD; }

/* ------------------------------------------------------------------ encoder handle ------------------ */
static void quiet_log( void *p, int level, const char *fmt, va_list ap ) { (void)p; (void)level; (void)fmt; (void)ap; }

/* opts: "key=value:key=value" passed one by one to x264_param_parse (common/base.c:886) */
XREF_API void *xref_open( int width, int height, const char *preset, const char *opts, int verbose )
{
    x264_param_t param;
    if( x264_param_default_preset( &param, preset && preset[0] ? preset : "medium", NULL ) < 0 )
        return NULL;
    param.i_width = width;
    param.i_height = height;
    param.i_csp = X264_CSP_I420;
    param.i_threads = 1;
    param.i_lookahead_threads = 1;
    param.i_fps_num = 25; param.i_fps_den = 1;
    param.b_vfr_input = 0;
    param.cpu = 0;                        /* C path: identical results to asm (checkasm) and deterministic */
    if( !verbose )
        param.pf_log = quiet_log;
    else
        param.i_log_level = X264_LOG_DEBUG;
    if( opts && opts[0] )
    {
        char *dup = strdup( opts );
        for( char *tok = strtok( dup, ":" ); tok; tok = strtok( NULL, ":" ) )
        {
            char *eq = strchr( tok, '=' );
            if( eq ) *eq = 0;
            int r = x264_param_parse( &param, tok, eq ? eq+1 : NULL );
            if( r < 0 )
            {
                fprintf( stderr, "xref_open: bad option %s=%s (%d)\n", tok, eq ? eq+1 : "", r );
                free( dup );
                return NULL;
            }
        }
        free( dup );
    }
    return x264_encoder_open( &param, NULL );
}

static x264_frame_t *me_frame;
static x264_t *me_frame_h;
XREF_API void xref_close( void *hv )
{
    if( !hv ) return;
    if( me_frame_h == hv ) { me_frame = NULL; me_frame_h = NULL; }      /* the cached frame dies with its encoder */
    x264_encoder_close( (x264_t*)hv );
}

XREF_API int xref_param( void *hv, const char *name )
{
    x264_t *h = hv;
#define P(n,v) if( !strcmp( name, n ) ) return (v);
    P( "mb_width", h->mb.i_mb_width ) P( "mb_height", h->mb.i_mb_height )
    P( "subme", h->param.analyse.i_subpel_refine ) P( "me", h->param.analyse.i_me_method )
    P( "merange", h->param.analyse.i_me_range ) P( "mvrange", h->param.analyse.i_mv_range )
    P( "bframes", h->param.i_bframe ) P( "b_bias", h->param.i_bframe_bias ) P( "b_adapt", h->param.i_bframe_adaptive )
    P( "weightb", h->param.analyse.b_weighted_bipred ) P( "weightp", h->param.analyse.i_weighted_pred )
    P( "aq_mode", h->param.rc.i_aq_mode ) P( "mbtree", h->param.rc.b_mb_tree ) P( "vbv", h->param.rc.i_vbv_buffer_size )
    P( "lookahead", h->param.rc.i_lookahead ) P( "lookahead_threads", h->param.i_lookahead_threads )
    P( "scenecut", h->param.i_scenecut_threshold ) P( "keyint_max", h->param.i_keyint_max ) P( "keyint_min", h->param.i_keyint_min )
    P( "sync_lookahead", h->param.i_sync_lookahead ) P( "threads", h->param.i_threads )
    P( "stride", h->fdec->i_stride[0] ) P( "stride_lowres", h->fdec->i_stride_lowres )
    P( "width_lowres", h->fdec->i_width_lowres ) P( "lines_lowres", h->fdec->i_lines_lowres )
    P( "psy", h->param.analyse.b_psy ) P( "ref", h->param.i_frame_reference )
    P( "open_gop", h->param.b_open_gop ) P( "intra_refresh", h->param.b_intra_refresh )
    P( "b_pyramid", h->param.i_bframe_pyramid ) P( "open_gop", h->param.b_open_gop ) P( "intra_refresh", h->param.b_intra_refresh )
#undef P
    return -9999;
}

/* ------------------------------------------------------------------ pixel table (B1) ---------------- */
static x264_pixel_function_t g_pf;
static x264_mc_functions_t   g_mc;
static int g_tables_ready;
static void tables_init( void )
{
    if( g_tables_ready ) return;
    x264_pixel_init( 0, &g_pf );
    x264_mc_init( 0, &g_mc, 1 );
    g_tables_ready = 1;
}

/* metric: 0 sad, 1 ssd, 2 satd, 3 sa8d (sa8d: i_pixel 0 = 16x16, 3 = 8x8 only) */
XREF_API int xref_pixel_cmp( int metric, int i_pixel, uint8_t *a, intptr_t sa, uint8_t *b, intptr_t sb )
{
    tables_init();
    switch( metric )
    {
        case 0: return g_pf.sad[i_pixel]( a, sa, b, sb );
        case 1: return g_pf.ssd[i_pixel]( a, sa, b, sb );
        case 2: return g_pf.satd[i_pixel]( a, sa, b, sb );
        case 3: return g_pf.sa8d[i_pixel]( a, sa, b, sb );
    }
    return -1;
}

typedef struct { uint32_t fenc_off, ref_off; } xref_cand_t;

XREF_API void xref_pixel_cmp_batch( int metric, int i_pixel, uint8_t *fenc, intptr_t fenc_stride,
                                    uint8_t *ref, intptr_t ref_stride, const xref_cand_t *cand, int n, int32_t *out )
{
    tables_init();
    x264_pixel_cmp_t fn = metric == 0 ? g_pf.sad[i_pixel] : metric == 1 ? g_pf.ssd[i_pixel]
                        : metric == 2 ? g_pf.satd[i_pixel] : g_pf.sa8d[i_pixel];
    for( int i = 0; i < n; i++ )
        out[i] = fn( fenc + cand[i].fenc_off, fenc_stride, ref + cand[i].ref_off, ref_stride );
}

/* x3/x4 entries: fenc is a FENC_STRIDE(16) block as in common/pixel.h:34-35 */
XREF_API void xref_pixel_cmp_x4( int metric, int i_pixel, uint8_t *fenc16, uint8_t *p0, uint8_t *p1, uint8_t *p2, uint8_t *p3,
                                 intptr_t stride, int *scores, int n_refs )
{
    tables_init();
    if( n_refs == 3 )
        ( metric == 0 ? g_pf.sad_x3[i_pixel] : g_pf.satd_x3[i_pixel] )( fenc16, p0, p1, p2, stride, scores );
    else
        ( metric == 0 ? g_pf.sad_x4[i_pixel] : g_pf.satd_x4[i_pixel] )( fenc16, p0, p1, p2, p3, stride, scores );
}

/* ------------------------------------------------------------------ mc table ------------------------ */
static void make_weight( x264_weight_t *w, int enabled, int scale, int denom, int offset )
{
    memset( w, 0, sizeof(*w) );
    if( enabled )
    {
        w->i_scale = scale; w->i_denom = denom; w->i_offset = offset;
        w->weightfn = g_mc.weight;
    }
}

XREF_API void xref_mc_luma( uint8_t *dst, intptr_t dst_stride, uint8_t *src0, uint8_t *src1, uint8_t *src2, uint8_t *src3,
                            intptr_t src_stride, int mvx, int mvy, int w, int h, int wt_en, int scale, int denom, int offset )
{
    tables_init();
    x264_weight_t wt; make_weight( &wt, wt_en, scale, denom, offset );
    pixel *src[4] = { src0, src1, src2, src3 };
    g_mc.mc_luma( dst, dst_stride, src, src_stride, mvx, mvy, w, h, &wt );
}

/* get_ref may return a pointer into the source: always copy the result into dst (stride dst_stride) */
XREF_API void xref_get_ref( uint8_t *dst, intptr_t dst_stride, uint8_t *src0, uint8_t *src1, uint8_t *src2, uint8_t *src3,
                            intptr_t src_stride, int mvx, int mvy, int w, int h, int wt_en, int scale, int denom, int offset )
{
    tables_init();
    x264_weight_t wt; make_weight( &wt, wt_en, scale, denom, offset );
    pixel *src[4] = { src0, src1, src2, src3 };
    ALIGNED_ARRAY_64( pixel, tmp,[32*32] );
    intptr_t st = 32;
    pixel *r = g_mc.get_ref( tmp, &st, src, src_stride, mvx, mvy, w, h, &wt );
    for( int y = 0; y < h; y++ )
        memcpy( dst + y*dst_stride, r + y*st, w );
}

XREF_API void xref_avg( int i_pixel, uint8_t *dst, intptr_t sd, uint8_t *a, intptr_t sa, uint8_t *b, intptr_t sb, int weight )
{
    tables_init();
    g_mc.avg[i_pixel]( dst, sd, a, sa, b, sb, weight );
}

XREF_API void xref_frame_init_lowres_core( uint8_t *src, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc,
                                           intptr_t src_stride, intptr_t dst_stride, int width, int height )
{
    tables_init();
    g_mc.frame_init_lowres_core( src, d0, dh, dv, dc, src_stride, dst_stride, width, height );
}

/* Run the reference's real per-frame lowres preparation (mc.c:458-482) on a luma picture of the encoder's
 * size; copies out the four padded lowres planes (stride_lowres x (lines_lowres+2*PADV)), origin included. */
XREF_API int xref_frame_lowres( void *hv, const uint8_t *luma, intptr_t luma_stride, uint8_t *out4 )
{
    x264_t *h = hv;
    x264_frame_t *f = x264_frame_pop_unused( h, 0 );
    if( !f ) return -1;
    for( int y = 0; y < h->param.i_height; y++ )
        memcpy( f->plane[0] + y*f->i_stride[0], luma + y*luma_stride, h->param.i_width );
    x264_frame_expand_border_mod16( h, f );
    x264_frame_init_lowres( h, f );
    intptr_t st = f->i_stride_lowres;
    size_t plane_bytes = (size_t)st * (f->i_lines_lowres + 2*PADV);
    for( int i = 0; i < 4; i++ )
        for( int y = 0; y < f->i_lines_lowres + 2*PADV; y++ )
            memcpy( out4 + i*plane_bytes + y*st, f->lowres[i] + (y-PADV)*st - PADH, f->i_width_lowres + 2*PADH );
    x264_frame_push_unused( h, f );
    return 0;
}

/* Run the reference's hpel pipeline the way fdec_filter_row does (encoder.c:2470-2480): per MB row
 * expand_border -> x264_frame_filter -> expand_border_filtered.  Copies out H,V,C padded planes
 * (stride x (lines+2*PADV)), starting at (-PADH,-PADV). */
XREF_API int xref_frame_hpel( void *hv, const uint8_t *luma, intptr_t luma_stride, uint8_t *out3, uint8_t *out_src )
{
    x264_t *h = hv;
    x264_frame_t *f = x264_frame_pop_unused( h, 1 );
    if( !f ) return -1;
    int W = h->mb.i_mb_width*16, H = h->mb.i_mb_height*16;
    for( int y = 0; y < H; y++ )
        memcpy( f->plane[0] + y*f->i_stride[0], luma + y*luma_stride, W );
    f->b_kept_as_ref = 1;
    h->i_threadslice_start = 0;
    h->i_threadslice_end = h->mb.i_mb_height;
    for( int mb_y = 0; mb_y < h->mb.i_mb_height; mb_y++ )
    {
        int end = mb_y == h->mb.i_mb_height - 1;
        x264_frame_expand_border( h, f, mb_y );
        x264_frame_filter( h, f, mb_y, end );
        x264_frame_expand_border_filtered( h, f, mb_y, end );
    }
    intptr_t st = f->i_stride[0];
    size_t plane_bytes = (size_t)st * (H + 2*PADV);
    for( int y = 0; y < H + 2*PADV; y++ )
    {
        for( int i = 1; i < 4; i++ )
            memcpy( out3 + (i-1)*plane_bytes + y*st, f->filtered[0][i] + (y-PADV)*st - PADH, W + 2*PADH );
        if( out_src )
            memcpy( out_src + y*st, f->plane[0] + (y-PADV)*st - PADH, W + 2*PADH );
    }
    x264_frame_push_unused( h, f );
    return 0;
}

XREF_API void xref_cost_mv_table( void *hv, uint16_t *out, int len )
{
    x264_t *h = hv;
    for( int i = -len; i <= len; i++ )
        out[len+i] = h->cost_mv[X264_LOOKAHEAD_QP][i];
}

/* ------------------------------------------------------------------ motion search ------------------- */
typedef struct
{
    int i_pixel, me_method, subpel_refine, me_range, qp;
    int mv_min_spel[2], mv_max_spel[2];
    int16_t mvp[2];
    int i_mvc;
    int16_t mvc[16][2];
    int wt_en, wt_scale, wt_denom, wt_offset;
    int use_thresh, halfpel_thresh;
    /* out */
    int16_t mv[2];
    int cost, cost_mv, thresh_out;
} xref_me_args_t;

/* chroma ME (h->mb.b_chroma_me, me.c:826-857): NV12 planes at the block's chroma origin and the weights of the two planes */
typedef struct
{
    uint8_t *fenc_uv; intptr_t fenc_uv_stride;
    uint8_t *fref_uv; intptr_t fref_uv_stride;
    int wt[2][4];                      /* m->weight[1], m->weight[2]: enabled, scale, denom, offset */
} xref_chroma_t;

/* Drives the reference's x264_me_search_ref (encoder/me.c:182) on caller-supplied planes.
 * fref[0..3] = F,H,V,C plane pointers at the block origin, fref_w = weighted full-pel plane (or fref[0]). */
static __thread int g_fpel_border;           /* i_fpel_border of the encoder's analysis (analyse.c:333-349); 0 = the lookahead's limits */
static void me_search_common( x264_t *h, xref_me_args_t *a, uint8_t *fenc, intptr_t fenc_stride,
                              uint8_t *f0, uint8_t *f1, uint8_t *f2, uint8_t *f3, uint8_t *fref_w, intptr_t stride, uint16_t *integral,
                              const xref_chroma_t *ch )
{
    tables_init();
    ALIGNED_ARRAY_64( pixel, fenc_buf,[16*16] );
    ALIGNED_ARRAY_64( pixel, fenc_c,[16*8] );          /* U at column 0, V at column 8, FENC_STRIDE: the layout of h->mb.pic.fenc_buf's chroma */
    int bw = x264_pixel_size[a->i_pixel].w, bh = x264_pixel_size[a->i_pixel].h;
    for( int y = 0; y < bh; y++ )
        memcpy( fenc_buf + y*FENC_STRIDE, fenc + y*fenc_stride, bw );
    x264_weight_t wt[3];
    make_weight( &wt[0], a->wt_en, a->wt_scale, a->wt_denom, a->wt_offset );
    make_weight( &wt[1], 0, 0, 0, 0 ); make_weight( &wt[2], 0, 0, 0, 0 );
    if( a->wt_en ) wt[0].weightfn = h->mc.weight;
    x264_me_t m;
    memset( &m, 0, sizeof(m) );
    if( ch )
    {
        for( int y = 0; y < bh/2; y++ )
            for( int x = 0; x < bw/2; x++ )
            {
                fenc_c[y*FENC_STRIDE + x]     = ch->fenc_uv[y*ch->fenc_uv_stride + 2*x];
                fenc_c[y*FENC_STRIDE + 8 + x] = ch->fenc_uv[y*ch->fenc_uv_stride + 2*x + 1];
            }
        m.p_fenc[1] = fenc_c; m.p_fenc[2] = fenc_c + 8;
        m.p_fref[4] = ch->fref_uv;
        m.i_stride[1] = ch->fref_uv_stride;
        for( int k = 0; k < 2; k++ )
        {
            make_weight( &wt[1+k], ch->wt[k][0], ch->wt[k][1], ch->wt[k][2], ch->wt[k][3] );
            if( ch->wt[k][0] ) wt[1+k].weightfn = h->mc.weight;
        }
    }
    m.i_pixel = a->i_pixel;
    m.p_cost_mv = h->cost_mv[a->qp];
    m.i_ref_cost = 0;
    m.i_ref = 0;
    m.weight = wt;
    m.p_fref[0] = f0; m.p_fref[1] = f1; m.p_fref[2] = f2; m.p_fref[3] = f3;
    m.p_fref_w = fref_w;
    m.integral = integral;
    m.p_fenc[0] = fenc_buf;
    m.i_stride[0] = stride;
    m.mvp[0] = a->mvp[0]; m.mvp[1] = a->mvp[1];
    int save_range = h->param.analyse.i_me_range;
    h->param.analyse.i_me_range = a->me_range;
    h->mb.i_me_method = a->me_method;
    h->mb.i_qp = a->qp;                                 /* ESA / TESA take the x mv costs from cost_mv_fpel[h->mb.i_qp] (me.c:639) */
    h->mb.i_subpel_refine = a->subpel_refine;
    h->mb.b_chroma_me = ch != NULL;
    h->mb.b_interlaced = 0;
    for( int i = 0; i < 2; i++ )
    {
        h->mb.mv_min_spel[i] = a->mv_min_spel[i];
        h->mb.mv_max_spel[i] = a->mv_max_spel[i];
        h->mb.mv_limit_fpel[0][i] = ( a->mv_min_spel[i] >> 2 ) + g_fpel_border;
        h->mb.mv_limit_fpel[1][i] = ( a->mv_max_spel[i] >> 2 ) - g_fpel_border;
    }
    ALIGNED_ARRAY_8( int16_t, mvc,[16],[2] );
    memcpy( mvc, a->mvc, sizeof(mvc) );
    int thresh = a->halfpel_thresh;
    x264_me_search_ref( h, &m, mvc, a->i_mvc, a->use_thresh ? &thresh : NULL );
    h->param.analyse.i_me_range = save_range;
    h->mb.b_chroma_me = 0;
    a->mv[0] = m.mv[0]; a->mv[1] = m.mv[1];
    a->cost = m.cost; a->cost_mv = m.cost_mv;
    a->thresh_out = thresh;
}

XREF_API void xref_me_search( void *hv, xref_me_args_t *a, uint8_t *fenc, intptr_t fenc_stride,
                              uint8_t *f0, uint8_t *f1, uint8_t *f2, uint8_t *f3, uint8_t *fref_w, intptr_t stride )
{
    me_search_common( hv, a, fenc, fenc_stride, f0, f1, f2, f3, fref_w, stride, NULL, NULL );
}

/* the same with chroma ME on (refine_subpel's COST_MV_SATD chroma branch, me.c:826-857; partitions of 8x8 and larger) */
XREF_API void xref_me_search_chroma( void *hv, xref_me_args_t *a, uint8_t *fenc, intptr_t fenc_stride,
                                     uint8_t *f0, uint8_t *f1, uint8_t *f2, uint8_t *f3, uint8_t *fref_w, intptr_t stride,
                                     const xref_chroma_t *ch )
{
    me_search_common( hv, a, fenc, fenc_stride, f0, f1, f2, f3, fref_w, stride, NULL, ch );
}

/* x264_me_refine_qpel (mode 0) / x264_me_refine_qpel_refdupe (mode 1) (encoder/me.c:800-814) on caller-supplied planes, from
 * a->mv / a->cost; a->qp, a->subpel_refine, limits, mvp, weights and the half-pel threshold as in xref_me_search */
XREF_API void xref_me_refine_qpel( void *hv, xref_me_args_t *a, int mode, int i_ref_cost, uint8_t *fenc, intptr_t fenc_stride,
                                   uint8_t *f0, uint8_t *f1, uint8_t *f2, uint8_t *f3, intptr_t stride )
{
    x264_t *h = hv;
    tables_init();
    ALIGNED_ARRAY_64( pixel, fenc_buf,[16*16] );
    int bw = x264_pixel_size[a->i_pixel].w, bh = x264_pixel_size[a->i_pixel].h;
    for( int y = 0; y < bh; y++ )
        memcpy( fenc_buf + y*FENC_STRIDE, fenc + y*fenc_stride, bw );
    x264_weight_t wt; make_weight( &wt, a->wt_en, a->wt_scale, a->wt_denom, a->wt_offset );
    if( a->wt_en ) wt.weightfn = h->mc.weight;
    x264_me_t m;
    memset( &m, 0, sizeof(m) );
    m.i_pixel = a->i_pixel;
    m.p_cost_mv = h->cost_mv[a->qp];
    m.i_ref_cost = i_ref_cost;
    m.weight = &wt;
    m.p_fref[0] = f0; m.p_fref[1] = f1; m.p_fref[2] = f2; m.p_fref[3] = f3;
    m.p_fref_w = f0;
    m.p_fenc[0] = fenc_buf;
    m.i_stride[0] = stride;
    m.mvp[0] = a->mvp[0]; m.mvp[1] = a->mvp[1];
    m.mv[0] = a->mv[0]; m.mv[1] = a->mv[1];
    m.cost = a->cost;
    h->mb.i_subpel_refine = a->subpel_refine;
    h->mb.b_chroma_me = 0;
    for( int i = 0; i < 2; i++ )
    {
        h->mb.mv_min_spel[i] = a->mv_min_spel[i];
        h->mb.mv_max_spel[i] = a->mv_max_spel[i];
    }
    int thresh = a->halfpel_thresh;
    if( mode == 0 ) x264_me_refine_qpel( h, &m );
    else            x264_me_refine_qpel_refdupe( h, &m, a->use_thresh ? &thresh : NULL );
    a->mv[0] = m.mv[0]; a->mv[1] = m.mv[1];
    a->cost = m.cost; a->cost_mv = m.cost_mv;
    a->thresh_out = thresh;
}

/* x264_me_refine_bidir_satd (encoder/me.c:1027-1183, rd = 0) on caller-supplied planes: f0[4] / f1[4] = F,H,V,C planes of the
 * list-0 / list-1 reference at the block origin.  mv0 / mv1 are updated in place. */
XREF_API void xref_me_refine_bidir_satd( void *hv, int i_pixel, int qp, uint8_t *fenc, intptr_t fenc_stride,
                                         uint8_t **f0, uint8_t **f1, intptr_t stride, int16_t *mv0, int16_t *mvp0,
                                         int16_t *mv1, int16_t *mvp1, int i_weight, const int *mv_min_spel, const int *mv_max_spel )
{
    x264_t *h = hv;
    tables_init();
    ALIGNED_ARRAY_64( pixel, fenc_buf,[16*16] );
    ALIGNED_ARRAY_64( pixel, fdec_buf,[32*FDEC_STRIDE] );
    int bw = x264_pixel_size[i_pixel].w, bh = x264_pixel_size[i_pixel].h;
    for( int y = 0; y < bh; y++ )
        memcpy( fenc_buf + y*FENC_STRIDE, fenc + y*fenc_stride, bw );
    x264_me_t m[2];
    memset( m, 0, sizeof(m) );
    for( int l = 0; l < 2; l++ )
    {
        m[l].i_pixel = i_pixel;
        m[l].p_cost_mv = h->cost_mv[qp];
        m[l].p_fenc[0] = fenc_buf;
        m[l].i_stride[0] = stride;
        for( int k = 0; k < 4; k++ ) m[l].p_fref[k] = ( l ? f1 : f0 )[k];
        m[l].p_fref_w = m[l].p_fref[0];
        m[l].weight = x264_weight_none;
        m[l].mvp[0] = ( l ? mvp1 : mvp0 )[0]; m[l].mvp[1] = ( l ? mvp1 : mvp0 )[1];
        m[l].mv[0] = ( l ? mv1 : mv0 )[0]; m[l].mv[1] = ( l ? mv1 : mv0 )[1];
    }
    pixel *save = h->mb.pic.p_fdec[0];
    h->mb.pic.p_fdec[0] = fdec_buf;
    for( int i = 0; i < 2; i++ )
    {
        h->mb.mv_min_spel[i] = mv_min_spel[i];
        h->mb.mv_max_spel[i] = mv_max_spel[i];
    }
    x264_me_refine_bidir_satd( h, &m[0], &m[1], i_weight );
    h->mb.pic.p_fdec[0] = save;
    mv0[0] = m[0].mv[0]; mv0[1] = m[0].mv[1];
    mv1[0] = m[1].mv[0]; mv1[1] = m[1].mv[1];
}

/* ESA / TESA read the reference frame's integral image (me.c:636-760), which x264_frame_filter builds when the encoder was
 * opened with me=esa|tesa: here the search runs against a reference frame the reference builds itself from ref_luma
 * (picture-sized, mod 16); (bx, by) = position of the block.  The frame is kept until ref_luma changes. */
static x264_frame_t *me_frame;           /* the reference frame xref_me_search_frame built last; dropped when its encoder closes */
static const uint8_t *me_frame_src;
static x264_t *me_frame_h;
static uint64_t me_frame_hash;

XREF_API int xref_me_search_frame( void *hv, xref_me_args_t *a, uint8_t *fenc, intptr_t fenc_stride,
                                   const uint8_t *ref_luma, intptr_t ref_stride, int bx, int by )
{
    x264_t *h = hv;
#define f me_frame
#define f_src me_frame_src
#define f_h me_frame_h
#define f_hash me_frame_hash
    /* the cached frame is identified by the content, not by the address (buffers get recycled at the same address) */
    uint64_t hash = 1469598103934665603ull;
    for( int y = 0; y < h->mb.i_mb_height*16; y++ )
        for( int x = 0; x < h->mb.i_mb_width*16; x++ )
            hash = ( hash ^ ref_luma[y*ref_stride + x] ) * 1099511628211ull;
    if( !f || f_src != ref_luma || f_h != h || f_hash != hash )
    {
        f_hash = hash;
        if( f && f_h == h ) x264_frame_push_unused( h, f );
        f = x264_frame_pop_unused( h, 1 );
        if( !f ) return -1;
        f_src = ref_luma; f_h = h;
        int W = h->mb.i_mb_width*16, H = h->mb.i_mb_height*16;
        for( int y = 0; y < H; y++ )
            memcpy( f->plane[0] + y*f->i_stride[0], ref_luma + y*ref_stride, W );
        f->b_kept_as_ref = 1;
        h->i_threadslice_start = 0;
        h->i_threadslice_end = h->mb.i_mb_height;
        for( int mb_y = 0; mb_y < h->mb.i_mb_height; mb_y++ )
        {
            int end = mb_y == h->mb.i_mb_height - 1;
            x264_frame_expand_border( h, f, mb_y );
            x264_frame_filter( h, f, mb_y, end );
            x264_frame_expand_border_filtered( h, f, mb_y, end );
        }
    }
    if( !f->integral ) return -2;                       /* the encoder was not opened with me=esa */
    intptr_t st = f->i_stride[0], off = by*st + bx;
    x264_frame_t *save = h->fenc;
    h->fenc = f;                                        /* me.c:645 reads h->fenc->i_lines[0] for the 4x4 plane of the integral */
    pixel *wbuf = NULL, *fref_w = f->filtered[0][0] + off;
    if( a->wt_en )
    {   /* the weighted full-pel plane of a weighted reference (encoder.c:2141-2166 builds it the same way): the whole padded plane */
        int lines = f->i_lines[0] + 2*PADV;
        wbuf = x264_malloc( st * lines );
        if( !wbuf ) return -1;
        x264_weight_t wt; make_weight( &wt, 1, a->wt_scale, a->wt_denom, a->wt_offset );
        wt.weightfn = h->mc.weight;
        x264_weight_scale_plane( h, wbuf, st, f->filtered[0][0] - PADV*st - PADH_ALIGN, st, st, lines, &wt );
        fref_w = wbuf + PADV*st + PADH_ALIGN + off;
    }
    me_search_common( h, a, fenc, fenc_stride, f->filtered[0][0] + off, f->filtered[0][1] + off, f->filtered[0][2] + off,
                      f->filtered[0][3] + off, fref_w, st, f->integral + off, NULL );
    if( wbuf ) x264_free( wbuf );
    h->fenc = save;
    return 0;
#undef f
#undef f_src
#undef f_h
#undef f_hash
}

XREF_API void xref_cost_mv_table_qp( void *hv, int qp, uint16_t *out, int len )
{
    x264_t *h = hv;
    for( int i = -len; i <= len; i++ )
        out[len+i] = h->cost_mv[qp][i];
}
XREF_API int xref_lambda( int qp ) { return x264_lambda_tab[qp]; }

/* ------------------------------------------------------------------ lowres lookahead ---------------- */
typedef struct
{
    x264_t *h;
    int n;
    x264_frame_t **frames;        /* the frames[] array handed to slicetype_frame_cost */
    x264_frame_t **slots;         /* when used through xref_la_remap: slot -> frame, frames[] is rebuilt per request */
    x264_mb_analysis_t a;
} xref_la_t;

XREF_API void *xref_la_new( void *hv, int n )
{
    x264_t *h = hv;
    xref_la_t *la = calloc( 1, sizeof(*la) );
    la->h = h; la->n = n;
    la->frames = calloc( n + 2, sizeof(x264_frame_t*) );
    la->slots = la->frames;
    lowres_context_init( h, &la->a );                    /* slicetype.c:45-61 */
    return la;
}

/* mc.c:458-482 on a fresh frame; inv_qscale (u16 per MB) optional, defaults to 256 */
XREF_API int xref_la_set_frame( void *lav, int idx, const uint8_t *luma, intptr_t luma_stride, const uint16_t *inv_qscale )
{
    xref_la_t *la = lav;
    x264_t *h = la->h;
    x264_frame_t *f = la->slots[idx];
    if( !f )
    {
        f = la->slots[idx] = x264_frame_pop_unused( h, 0 );
        if( !f ) return -1;
    }
    for( int y = 0; y < h->param.i_height; y++ )
        memcpy( f->plane[0] + y*f->i_stride[0], luma + y*luma_stride, h->param.i_width );
    x264_frame_expand_border_mod16( h, f );
    /* frame statistics the lookahead weight analysis reads (i_pixel_sum / i_pixel_ssd), exactly as the encoder fills them
     * before x264_frame_init_lowres (encoder.c:3417); chroma is flat 128 */
    for( int y = 0; y < f->i_lines[1]; y++ )
        memset( f->plane[1] + y*f->i_stride[1], 128, f->i_width[1] );
    x264_adaptive_quant_frame( h, f, NULL );
    x264_frame_init_lowres( h, f );
    f->b_intra_calculated = 0;
    f->i_frame = idx;
    if( f->i_inv_qscale_factor )
        for( int i = 0; i < h->mb.i_mb_count; i++ )
            f->i_inv_qscale_factor[i] = inv_qscale ? inv_qscale[i] : 256;
    /* stale vectors of a recycled frame: the reference's own frames start zeroed (frame.c:287-293) */
    for( int l = 0; l <= !!h->param.i_bframe; l++ )
        for( int d = 0; d <= h->param.i_bframe; d++ )
        {
            memset( f->lowres_mvs[l][d], 0, 2*h->mb.i_mb_count*sizeof(int16_t) );
            f->lowres_mvs[l][d][0][0] = 0x7FFF;
        }
    return 0;
}

XREF_API int xref_la_frame_cost( void *lav, int p0, int p1, int b )
{
    xref_la_t *la = lav;
    return slicetype_frame_cost( la->h, &la->a, la->frames, p0, p1, b );
}

/* what: 0 lowres_mvs[i][j] (int16 x2 per MB), 1 lowres_mv_costs[i][j] (int), 2 lowres_costs[i][j] (u16),
 *       3 i_intra_cost (int... stored as u16 in the reference: widened), 4 {cost_est, cost_est_aq, intra_mbs[i]}, 5 row_satds[i][j] */
XREF_API void xref_la_get( void *lav, int idx, int what, int i, int j, void *out )
{
    xref_la_t *la = lav;
    x264_t *h = la->h;
    x264_frame_t *f = idx >= 300 ? la->slots[idx-300] : la->frames[idx];      /* >= 300: address by slot (xref_la_remap users) */
    int n = h->mb.i_mb_count;
    switch( what )
    {
        case 0: memcpy( out, f->lowres_mvs[i][j], n * 4 ); break;
        case 1: memcpy( out, f->lowres_mv_costs[i][j], n * sizeof(int) ); break;
        case 2: memcpy( out, f->lowres_costs[i][j], n * 2 ); break;
        case 3: for( int k = 0; k < n; k++ ) ((int*)out)[k] = f->i_intra_cost[k]; break;
        case 4: ((int*)out)[0] = f->i_cost_est[i][j]; ((int*)out)[1] = f->i_cost_est_aq[i][j]; ((int*)out)[2] = f->i_intra_mbs[i]; break;
        case 5: memcpy( out, f->i_row_satds[i][j], h->mb.i_mb_height * sizeof(int) ); break;
        case 6: ((int*)out)[0] = !!f->weight[0][0].weightfn; ((int*)out)[1] = f->weight[0][0].i_scale;
                ((int*)out)[2] = f->weight[0][0].i_denom; ((int*)out)[3] = f->weight[0][0].i_offset; break;
    }
}

/* MB-tree (slicetype.c:1029-1184), one call each, so that a test can replay macroblock_tree's sequence step by step */
XREF_API void xref_la_set_type( void *lav, int idx, int type, float duration )
{
    xref_la_t *la = lav;
    la->frames[idx]->i_type = type;
    la->frames[idx]->f_duration = duration;
}
#define XREF_LA_FRAME( la, idx ) ( (idx) >= 300 ? (la)->slots[(idx)-300] : (la)->frames[idx] )     /* >= 300: by slot, as in xref_la_get */
XREF_API void xref_la_mbtree_reset( void *lav, int idx )
{
    xref_la_t *la = lav;
    memset( XREF_LA_FRAME( la, idx )->i_propagate_cost, 0, la->h->mb.i_mb_count * sizeof(uint16_t) );
}
XREF_API void xref_la_mbtree_swap( void *lav, int a, int b )
{
    xref_la_t *la = lav;
    XCHG( uint16_t*, XREF_LA_FRAME( la, a )->i_propagate_cost, XREF_LA_FRAME( la, b )->i_propagate_cost );
}
XREF_API void xref_la_mbtree_propagate( void *lav, float average_duration, int p0, int p1, int b, int referenced )
{
    xref_la_t *la = lav;
    macroblock_tree_propagate( la->h, la->frames, average_duration, p0, p1, b, referenced );
}
XREF_API void xref_la_mbtree_finish( void *lav, int idx, float average_duration, int ref0_distance )
{
    xref_la_t *la = lav;
    macroblock_tree_finish( la->h, XREF_LA_FRAME( la, idx ), average_duration, ref0_distance );
}
/* the whole of macroblock_tree( h, a, frames, num_frames, b_intra ) on frames[0..num_frames] with the types set before */
XREF_API void xref_la_mbtree( void *lav, int num_frames, int b_intra )
{
    xref_la_t *la = lav;
    macroblock_tree( la->h, &la->a, la->frames, num_frames, b_intra );
}
/* slicetype_frame_cost_recalculate( h, frames, p0, p1, b ) (slicetype.c:999-1024); frames[b]->i_type picks the offsets */
XREF_API int xref_la_frame_cost_recalculate( void *lav, int p0, int p1, int b )
{
    xref_la_t *la = lav;
    return slicetype_frame_cost_recalculate( la->h, la->frames, p0, p1, b );
}
/* the same by frame (idx as in xref_la_get) and distances, with the picture type given: b_type != 0 = a B picture */
XREF_API int xref_la_frame_cost_recalculate_at( void *lav, int idx, int i0, int i1, int b_type, int *rows )
{
    xref_la_t *la = lav;
    x264_frame_t *fr[2*X264_BFRAME_MAX + 8] = { NULL };
    x264_frame_t *f = XREF_LA_FRAME( la, idx );
    int save = f->i_type;
    f->i_type = b_type ? X264_TYPE_B : X264_TYPE_P;
    fr[i0] = f;
    int score = slicetype_frame_cost_recalculate( la->h, fr, 0, i0 + i1, i0 );
    f->i_type = save;
    if( rows ) memcpy( rows, f->i_row_satds[i0][i1], la->h->mb.i_mb_height * sizeof(int) );
    return score;
}
/* what: 0 f_qp_offset, 1 f_qp_offset_aq (float per MB), 2 i_propagate_cost (u16 per MB), 3 f_weighted_cost_delta[i] (one float) */
XREF_API void xref_la_get_mbtree( void *lav, int idx, int what, int i, void *out )
{
    xref_la_t *la = lav;
    x264_frame_t *f = XREF_LA_FRAME( la, idx );
    int n = la->h->mb.i_mb_count;
    switch( what )
    {
        case 0: memcpy( out, f->f_qp_offset, n * sizeof(float) ); break;
        case 1: memcpy( out, f->f_qp_offset_aq, n * sizeof(float) ); break;
        case 2: memcpy( out, f->i_propagate_cost, n * sizeof(uint16_t) ); break;
        case 3: *(float*)out = f->f_weighted_cost_delta[i]; break;
    }
}
XREF_API void xref_la_set_qp_offset_aq( void *lav, int idx, const float *aq )
{
    xref_la_t *la = lav;
    memcpy( la->frames[idx]->f_qp_offset_aq, aq, la->h->mb.i_mb_count * sizeof(float) );
    memcpy( la->frames[idx]->f_qp_offset, aq, la->h->mb.i_mb_count * sizeof(float) );
}

XREF_API void xref_la_get_lowres( void *lav, int idx, int plane, uint8_t *out )
{
    xref_la_t *la = lav;
    x264_frame_t *f = la->frames[idx];
    intptr_t st = f->i_stride_lowres;
    for( int y = 0; y < f->i_lines_lowres + 2*PADV; y++ )
        memcpy( out + y*st, f->lowres[plane] + (y-PADV)*st - PADH, f->i_width_lowres + 2*PADH );
}

/* slot-addressed use (bench / slicetype glue): frames[i] = slot table[frames_in[i]] for i in p0..p1 */
XREF_API void xref_la_remap( void *lav, const int *frames_in, int p0, int p1 )
{
    xref_la_t *la = lav;
    if( la->slots == la->frames )
    {
        la->slots = calloc( la->n + 2, sizeof(x264_frame_t*) );
        memcpy( la->slots, la->frames, ( la->n + 2 ) * sizeof(x264_frame_t*) );
        la->frames = calloc( 512, sizeof(x264_frame_t*) );
    }
    for( int i = p0; i <= p1; i++ ) la->frames[i] = la->slots[frames_in[i]];
}

XREF_API void xref_la_free( void *lav )
{
    xref_la_t *la = lav;
    if( la->slots != la->frames ) { free( la->frames ); la->frames = la->slots; }
    /* deleted, not pushed back: the encoder's unused-frame list has a fixed capacity (encoder.c frames.unused) */
    for( int i = 0; i < la->n; i++ )
        if( la->frames[i] ) x264_frame_delete( la->frames[i] );
    free( la->frames );
    free( la );
}

/* ------------------------------------------------------------------ whole-encoder frame types ------- */
/* Encode n luma pictures (chroma = 128) with the opened encoder and report, in coded (output) order, the display
 * index (pts) and decided type (X264_TYPE_*) of every frame.  This is the reference's own answer to "which slice
 * types does the lookahead choose".  Returns the number of frames output. */
/* when set: xref_encode_types also stores, per coded frame, the f_qp_offset array the encoder used for it (MB-tree's output;
 * mb_count floats each, coded order) */
static float *xref_qp_capture;
XREF_API void xref_set_qp_capture( float *buf ) { xref_qp_capture = buf; }
#define XREF_CAPTURE_QP() do { if( xref_qp_capture && h->fenc ) \
    memcpy( xref_qp_capture + (size_t)n_out * h->mb.i_mb_count, h->fenc->f_qp_offset, h->mb.i_mb_count * sizeof(float) ); } while( 0 )

/* when set: xref_encode_types also stores, per coded frame, `stride` ints: [0] = the cost x264_rc_analyse_slice returns for it
 * (called again here right after the frame was coded: same inputs, same result; -1 where the encoder does not call it: B
 * pictures without VBV, CQP), [1 .. mb_height] = its i_row_satd, then n = the number of planned entries (VBV lookahead,
 * slicetype.c:1225-1286), i_planned_type[0..31] and i_planned_satd[0..31] */
static int *xref_rc_capture; static int xref_rc_stride;
XREF_API void xref_set_rc_capture( int *buf, int stride ) { xref_rc_capture = buf; xref_rc_stride = stride; }
static void capture_rc( x264_t *h, int n_out )
{
    if( !xref_rc_capture || !h->fenc ) return;
    int *o = xref_rc_capture + (size_t)n_out * xref_rc_stride;
    int mbh = h->mb.i_mb_height;
    o[0] = -1;
    if( h->param.rc.i_rc_method != X264_RC_CQP && ( !IS_X264_TYPE_B( h->fenc->i_type ) || h->param.rc.i_vbv_buffer_size ) )
    {
        o[0] = x264_rc_analyse_slice( h );
        memcpy( o + 1, h->fenc->i_row_satd, mbh * sizeof(int) );
    }
    int n = 0;
    if( !IS_X264_TYPE_B( h->fenc->i_type ) && h->param.rc.i_vbv_buffer_size && h->param.rc.i_lookahead )
        while( n < 32 && h->fenc->i_planned_type[n] != X264_TYPE_AUTO ) n++;
    o[1 + mbh] = n;
    for( int i = 0; i < n; i++ ) { o[2 + mbh + i] = h->fenc->i_planned_type[i]; o[2 + mbh + 32 + i] = h->fenc->i_planned_satd[i]; }
}

/* when set: pic_in.i_type of picture i (X264_TYPE_*; forced frame types as a qpfile / an application would give them) */
static const int *xref_forced_types;
XREF_API void xref_set_forced_types( const int *types ) { xref_forced_types = types; }

XREF_API int xref_encode_types( void *hv, const uint8_t *luma, int n, int *out_idx, int *out_type )
{
    x264_t *h = hv;
    int w = h->param.i_width, ht = h->param.i_height;
    int cw = ( w + 1 ) / 2, ch = ( ht + 1 ) / 2;
    uint8_t *chroma = malloc( (size_t)cw * ch );
    memset( chroma, 128, (size_t)cw * ch );
    int n_out = 0;
    x264_nal_t *nal; int i_nal;
    x264_picture_t pic_in, pic_out;
    for( int i = 0; i < n; i++ )
    {
        x264_picture_init( &pic_in );
        pic_in.img.i_csp = X264_CSP_I420;
        pic_in.img.i_plane = 3;
        pic_in.img.plane[0] = (uint8_t*)luma + (size_t)i * w * ht; pic_in.img.i_stride[0] = w;
        pic_in.img.plane[1] = chroma; pic_in.img.i_stride[1] = cw;
        pic_in.img.plane[2] = chroma; pic_in.img.i_stride[2] = cw;
        pic_in.i_pts = i;
        pic_in.i_type = xref_forced_types ? xref_forced_types[i] : X264_TYPE_AUTO;
        int sz = x264_encoder_encode( h, &nal, &i_nal, &pic_in, &pic_out );
        if( sz < 0 ) { free( chroma ); return -1; }
        if( sz > 0 ) { XREF_CAPTURE_QP(); capture_rc( h, n_out ); out_idx[n_out] = (int)pic_out.i_pts; out_type[n_out] = pic_out.i_type; n_out++; }
    }
    while( x264_encoder_delayed_frames( h ) > 0 )
    {
        int sz = x264_encoder_encode( h, &nal, &i_nal, NULL, &pic_out );
        if( sz < 0 ) { free( chroma ); return -1; }
        if( sz > 0 ) { XREF_CAPTURE_QP(); capture_rc( h, n_out ); out_idx[n_out] = (int)pic_out.i_pts; out_type[n_out] = pic_out.i_type; n_out++; }
    }
    free( chroma );
    return n_out;
}

/* ------------------------------------------------------------------ adaptive quantisation ------------ */
/* x264_adaptive_quant_frame (encoder/ratecontrol.c:305-420) on an I420 picture, the way x264_encoder_encode prepares it
 * (x264_frame_copy_picture -> x264_frame_expand_border_mod16 -> x264_adaptive_quant_frame, encoder.c:3382-3417).
 * out: f_qp_offset_aq (float per MB), i_inv_qscale_factor (u16 per MB), stats[6] = i_pixel_sum[3], i_pixel_ssd[3] */
XREF_API int xref_aq_frame( void *hv, const uint8_t *luma, const uint8_t *cb, const uint8_t *cr, float *qp_offset_aq,
                            uint16_t *inv_qscale, uint64_t *stats )
{
    x264_t *h = hv;
    x264_frame_t *f = x264_frame_pop_unused( h, 0 );
    if( !f ) return -1;
    int w = h->param.i_width, ht = h->param.i_height;
    x264_picture_t pic;
    x264_picture_init( &pic );
    pic.img.i_csp = X264_CSP_I420;
    pic.img.i_plane = 3;
    pic.img.plane[0] = (uint8_t*)luma; pic.img.i_stride[0] = w;
    pic.img.plane[1] = (uint8_t*)cb;   pic.img.i_stride[1] = ( w + 1 ) / 2;
    pic.img.plane[2] = (uint8_t*)cr;   pic.img.i_stride[2] = ( w + 1 ) / 2;
    if( x264_frame_copy_picture( h, f, &pic ) < 0 ) return -1;
    if( h->param.i_width != 16 * h->mb.i_mb_width || h->param.i_height != 16 * h->mb.i_mb_height )
        x264_frame_expand_border_mod16( h, f );
    x264_adaptive_quant_frame( h, f, NULL );
    int n = h->mb.i_mb_count;
    memcpy( qp_offset_aq, f->f_qp_offset_aq, n * sizeof(float) );
    if( f->i_inv_qscale_factor ) memcpy( inv_qscale, f->i_inv_qscale_factor, n * sizeof(uint16_t) );
    for( int i = 0; i < 3; i++ ) { stats[i] = f->i_pixel_sum[i]; stats[3+i] = f->i_pixel_ssd[i]; }
    x264_frame_push_unused( h, f );
    return 0;
}


/* ------------------------------------------------------------------ recorded x264_me_t stream ------- */
/* BASELINE config 3: every x264_me_t the reference's analysis hands to x264_me_search_ref while it encodes
 * (analyse.c:1287, :1392-1785, :1938, :2231-2491: all partition sizes, every reference of both lists), with the result the
 * reference got, plus copies of the planes those searches read -- so that the whole stream can be replayed elsewhere. */
typedef struct
{
    int32_t i_pixel, bx, by, ref_idx, qp, lambda, i_mvc, thresh_in;     /* ref_idx: index into the traced frame's reference table */
    int16_t mvp[2], mvc[9][2], lim[4];                                   /* lim = mv_min_spel[0..1], mv_max_spel[0..1] */
    int16_t mv[2];
    int32_t cost, cost_mv, thresh_out;
} xref_me_rec_t;

typedef struct
{
    int list, i_ref, display, weighted;
    int weight[3][4];
    const pixel *key;                    /* h->mb.pic.p_fref[list][i_ref][0] - MB offset: identifies the planes */
    const x264_weight_t *wkey;
    uint8_t *planes[4], *wplane, *uv;    /* copies, origin at PADV*stride + PADH (luma) / (PADV/2)*stride_uv + PADH (chroma) */
} xref_trace_ref_t;

typedef struct
{
    int coded, display, slice_type, chroma_me, me_method, subpel, me_range, mbcmp_satd, mv_range, fpel_border;
    int width, height, stride, lines, stride_uv, lines_uv;
    uint8_t *fenc, *fenc_uv;             /* plane[0] / plane[1] of the source picture, from pixel (0,0) */
    xref_trace_ref_t refs[40]; int n_refs;
    xref_me_rec_t *recs; int n_recs, cap;
} xref_trace_frame_t;

static xref_trace_frame_t *trace_frames; static int trace_n, trace_max, trace_skip, trace_on, trace_last_coded = -1;

static uint8_t *trace_copy_padded( const pixel *origin, int stride, int lines, int padv )
{
    size_t n = (size_t)stride * ( lines + 2*padv );
    uint8_t *p = calloc( n, 1 );
    if( p ) memcpy( p, origin - (intptr_t)padv*stride - PADH, n - 64 );
    return p;
}

static xref_trace_frame_t *trace_frame_for( x264_t *h )
{
    if( h->i_frame != trace_last_coded )
    {
        trace_last_coded = h->i_frame;
        if( h->i_frame < trace_skip || trace_n >= trace_max ) return NULL;
        xref_trace_frame_t *f = &trace_frames[trace_n++];
        memset( f, 0, sizeof(*f) );
        f->coded = h->i_frame; f->display = h->fenc->i_frame; f->slice_type = h->sh.i_type;
        f->chroma_me = h->mb.b_chroma_me; f->me_method = h->mb.i_me_method; f->subpel = h->mb.i_subpel_refine;
        f->me_range = h->param.analyse.i_me_range; f->mbcmp_satd = h->pixf.mbcmp[0] == h->pixf.satd[0];
        f->mv_range = h->param.analyse.i_mv_range; f->fpel_border = 6;                /* i_fpel_border, analyse.c:333 */
        f->width = h->param.i_width; f->height = h->param.i_height;
        f->stride = h->fenc->i_stride[0]; f->lines = h->fenc->i_lines[0];
        f->stride_uv = h->fenc->i_stride[1]; f->lines_uv = h->fenc->i_lines[1];
        f->fenc = malloc( (size_t)f->stride * f->lines ); f->fenc_uv = malloc( (size_t)f->stride_uv * f->lines_uv );
        memcpy( f->fenc, h->fenc->plane[0], (size_t)f->stride * f->lines - 64 );
        memcpy( f->fenc_uv, h->fenc->plane[1], (size_t)f->stride_uv * f->lines_uv - 64 );
    }
    else if( !trace_n || trace_frames[trace_n-1].coded != h->i_frame )
        return NULL;
    return &trace_frames[trace_n-1];
}

static void xref_traced_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh )
{
    xref_trace_frame_t *f = NULL;
    xref_me_rec_t rec;
    int list = -1;
    if( trace_on && h->fenc && m->i_stride[0] == h->mb.pic.i_stride[0] )
    {   /* a call of the full-resolution analysis reads one of the current macroblock's reference windows (LOAD_HPELS,
         * analyse.c:1217-1246); the lookahead's lowres searches (slicetype.c:694) do not and are left out */
        for( int l = 0; l < 2 && list < 0; l++ )
            if( m->i_ref >= 0 && m->i_ref < h->mb.pic.i_fref[l] )
            {
                intptr_t d = m->p_fref[0] - h->mb.pic.p_fref[l][m->i_ref][0];
                if( d >= 0 && d < 16*m->i_stride[0] && d % m->i_stride[0] < 16 ) list = l;
            }
        if( list >= 0 ) f = trace_frame_for( h );
    }
    if( f )
    {
        intptr_t d = m->p_fenc[0] - h->mb.pic.p_fenc[0];
        int xoff = d % FENC_STRIDE, yoff = d / FENC_STRIDE;
        const pixel *key = h->mb.pic.p_fref[list][m->i_ref][0] - ( 16*h->mb.i_mb_x + 16*h->mb.i_mb_y*(intptr_t)m->i_stride[0] );
        int r;
        for( r = 0; r < f->n_refs; r++ )
            if( f->refs[r].key == key && f->refs[r].wkey == m->weight && f->refs[r].list == list ) break;
        if( r == f->n_refs && r < 40 )
        {
            xref_trace_ref_t *t = &f->refs[f->n_refs++];
            x264_frame_t *fr = h->fref[list][m->i_ref];
            t->list = list; t->i_ref = m->i_ref; t->display = fr->i_frame; t->key = key; t->wkey = m->weight;
            for( int k = 0; k < 3; k++ )
            {
                t->weight[k][0] = m->weight[k].weightfn != NULL; t->weight[k][1] = m->weight[k].i_scale;
                t->weight[k][2] = m->weight[k].i_denom; t->weight[k][3] = m->weight[k].i_offset;
            }
            for( int k = 0; k < 4; k++ ) t->planes[k] = trace_copy_padded( fr->filtered[0][k], f->stride, f->lines, PADV );
            t->uv = trace_copy_padded( fr->plane[1], f->stride_uv, f->lines_uv, PADV >> 1 );
            t->weighted = m->p_fref_w != m->p_fref[0];
            if( t->weighted ) t->wplane = trace_copy_padded( h->fenc->weighted[m->i_ref], f->stride, f->lines, PADV );
        }
        if( r >= 40 ) f = NULL;
        else
        {
            memset( &rec, 0, sizeof(rec) );
            rec.i_pixel = m->i_pixel; rec.bx = 16*h->mb.i_mb_x + xoff; rec.by = 16*h->mb.i_mb_y + yoff; rec.ref_idx = r;
            rec.qp = -1;
            for( int q = 0; q <= QP_MAX; q++ ) if( h->cost_mv[q] == m->p_cost_mv ) rec.qp = q;
            rec.lambda = rec.qp >= 0 ? x264_lambda_tab[rec.qp] : -1;
            rec.i_mvc = i_mvc; rec.thresh_in = p_halfpel_thresh ? *p_halfpel_thresh : -1;
            rec.mvp[0] = m->mvp[0]; rec.mvp[1] = m->mvp[1];
            for( int i = 0; i < i_mvc && i < 9; i++ ) { rec.mvc[i][0] = mvc[i][0]; rec.mvc[i][1] = mvc[i][1]; }
            rec.lim[0] = h->mb.mv_min_spel[0]; rec.lim[1] = h->mb.mv_min_spel[1];
            rec.lim[2] = h->mb.mv_max_spel[0]; rec.lim[3] = h->mb.mv_max_spel[1];
        }
    }
    x264_me_search_ref( h, m, mvc, i_mvc, p_halfpel_thresh );
    if( f )
    {
        rec.mv[0] = m->mv[0]; rec.mv[1] = m->mv[1]; rec.cost = m->cost; rec.cost_mv = m->cost_mv;
        rec.thresh_out = p_halfpel_thresh ? *p_halfpel_thresh : -1;
        if( f->n_recs == f->cap )
        {
            f->cap = f->cap ? 2*f->cap : 1 << 16;
            f->recs = realloc( f->recs, (size_t)f->cap * sizeof(rec) );
        }
        f->recs[f->n_recs++] = rec;
    }
}

XREF_API void xref_me_trace_free( void )
{
    for( int i = 0; i < trace_n; i++ )
    {
        xref_trace_frame_t *f = &trace_frames[i];
        free( f->fenc ); free( f->fenc_uv ); free( f->recs );
        for( int r = 0; r < f->n_refs; r++ )
        {
            for( int k = 0; k < 4; k++ ) free( f->refs[r].planes[k] );
            free( f->refs[r].wplane ); free( f->refs[r].uv );
        }
    }
    free( trace_frames ); trace_frames = NULL; trace_n = trace_max = 0; trace_on = 0; trace_last_coded = -1;
}
/* record the searches of up to max_frames coded pictures, starting with coded picture number `skip` */
XREF_API int xref_me_trace_start( int max_frames, int skip )
{
    xref_me_trace_free();
    trace_frames = calloc( max_frames, sizeof(*trace_frames) );
    if( !trace_frames ) return -1;
    trace_max = max_frames; trace_skip = skip; trace_on = 1;
    return 0;
}
XREF_API void xref_me_trace_stop( void ) { trace_on = 0; }
XREF_API int xref_me_trace_frames( void ) { return trace_n; }
/* info[0..17] = coded, display, slice_type, chroma_me, me_method, subpel, me_range, mbcmp_satd, mv_range, fpel_border, width,
 * height, stride, lines, stride_uv, lines_uv, n_refs, n_recs */
XREF_API int xref_me_trace_frame_info( int i, int *info )
{
    if( i < 0 || i >= trace_n ) return -1;
    xref_trace_frame_t *f = &trace_frames[i];
    memcpy( info, &f->coded, 16 * sizeof(int) );
    info[16] = f->n_refs; info[17] = f->n_recs;
    return 0;
}
XREF_API const void *xref_me_trace_recs( int i ) { return i >= 0 && i < trace_n ? trace_frames[i].recs : NULL; }
/* which: -1 fenc luma, -2 fenc chroma (ref ignored); 0..3 F,H,V,C of reference `ref`, 4 its weighted plane (or NULL), 5 its chroma */
XREF_API const void *xref_me_trace_plane( int i, int ref, int which )
{
    if( i < 0 || i >= trace_n ) return NULL;
    xref_trace_frame_t *f = &trace_frames[i];
    if( which == -1 ) return f->fenc;
    if( which == -2 ) return f->fenc_uv;
    if( ref < 0 || ref >= f->n_refs ) return NULL;
    return which < 4 ? f->refs[ref].planes[which] : which == 4 ? f->refs[ref].wplane : f->refs[ref].uv;
}
/* info[0..15] = list, i_ref, display, weighted, weight[3][4] */
XREF_API int xref_me_trace_ref_info( int i, int ref, int *info )
{
    if( i < 0 || i >= trace_n || ref < 0 || ref >= trace_frames[i].n_refs ) return -1;
    memcpy( info, &trace_frames[i].refs[ref].list, 16 * sizeof(int) );
    return 0;
}

/* Encode n I420 pictures (planes packed: luma w*h, then Cb, then Cr, picture after picture) with the opened encoder; the
 * bitstream is discarded.  Returns the number of frames output. */
XREF_API int xref_encode_i420( void *hv, const uint8_t *yuv, int n )
{
    x264_t *h = hv;
    int w = h->param.i_width, ht = h->param.i_height;
    int cw = ( w + 1 ) / 2, ch = ( ht + 1 ) / 2;
    size_t fsz = (size_t)w*ht + 2*(size_t)cw*ch;
    int n_out = 0;
    x264_nal_t *nal; int i_nal;
    x264_picture_t pic_in, pic_out;
    for( int i = 0; i < n; i++ )
    {
        x264_picture_init( &pic_in );
        pic_in.img.i_csp = X264_CSP_I420;
        pic_in.img.i_plane = 3;
        pic_in.img.plane[0] = (uint8_t*)yuv + i*fsz;             pic_in.img.i_stride[0] = w;
        pic_in.img.plane[1] = pic_in.img.plane[0] + (size_t)w*ht; pic_in.img.i_stride[1] = cw;
        pic_in.img.plane[2] = pic_in.img.plane[1] + (size_t)cw*ch; pic_in.img.i_stride[2] = cw;
        pic_in.i_pts = i;
        int r = x264_encoder_encode( h, &nal, &i_nal, &pic_in, &pic_out );
        if( r < 0 ) return -1;
        if( r > 0 ) n_out++;
    }
    while( x264_encoder_delayed_frames( h ) )
    {
        int r = x264_encoder_encode( h, &nal, &i_nal, NULL, &pic_out );
        if( r < 0 ) return -1;
        if( r > 0 ) n_out++;
    }
    return n_out;
}


/* Re-run the reference's x264_me_search_ref on recorded searches (xref_me_rec_t) against caller-held copies of the planes: the
 * CPU arm of bench.py --workload me (one encoder handle per thread) and a self-check of the recorder.  info = the 18 ints of
 * xref_me_trace_frame_info; planes as xref_me_trace_plane returns them; out = 5 ints per search: mv, cost, cost_mv, threshold. */
typedef struct { uint8_t *planes[4], *wplane, *uv; int weight[3][4]; } xref_replay_ref_t;

XREF_API void xref_me_replay( void *hv, const int *info, uint8_t *fenc, uint8_t *fenc_uv, const xref_replay_ref_t *refs,
                              const xref_me_rec_t *recs, int n, int32_t *out )
{
    x264_t *h = hv;
    const int chroma_me = info[3], me_method = info[4], subpel = info[5], me_range = info[6];
    const int stride = info[12], stride_uv = info[14];
    const intptr_t lo = (intptr_t)PADV*stride + PADH, co = (intptr_t)(PADV>>1)*stride_uv + PADH;
    g_fpel_border = info[9];
    for( int i = 0; i < n; i++ )
    {
        const xref_me_rec_t *r = &recs[i];
        const xref_replay_ref_t *f = &refs[r->ref_idx];
        xref_me_args_t a;
        memset( &a, 0, sizeof(a) );
        a.i_pixel = r->i_pixel; a.me_method = me_method; a.subpel_refine = subpel; a.me_range = me_range; a.qp = r->qp;
        for( int k = 0; k < 2; k++ ) { a.mv_min_spel[k] = r->lim[k]; a.mv_max_spel[k] = r->lim[2+k]; a.mvp[k] = r->mvp[k]; }
        a.i_mvc = r->i_mvc;
        memcpy( a.mvc, r->mvc, sizeof(r->mvc) );
        a.wt_en = f->weight[0][0]; a.wt_scale = f->weight[0][1]; a.wt_denom = f->weight[0][2]; a.wt_offset = f->weight[0][3];
        a.use_thresh = r->thresh_in >= 0; a.halfpel_thresh = r->thresh_in;
        intptr_t off = lo + (intptr_t)r->by*stride + r->bx;
        xref_chroma_t ch;
        ch.fenc_uv = fenc_uv + (intptr_t)(r->by>>1)*stride_uv + (r->bx&~1); ch.fenc_uv_stride = stride_uv;
        ch.fref_uv = f->uv + co + (intptr_t)(r->by>>1)*stride_uv + (r->bx&~1); ch.fref_uv_stride = stride_uv;
        memcpy( ch.wt, f->weight[1], sizeof(ch.wt) );
        me_search_common( h, &a, fenc + (intptr_t)r->by*stride + r->bx, stride, f->planes[0] + off, f->planes[1] + off,
                          f->planes[2] + off, f->planes[3] + off, ( f->wplane ? f->wplane : f->planes[0] ) + off, stride, NULL,
                          chroma_me ? &ch : NULL );
        out[5*i] = a.mv[0]; out[5*i+1] = a.mv[1]; out[5*i+2] = a.cost; out[5*i+3] = a.cost_mv; out[5*i+4] = a.use_thresh ? a.thresh_out : -1;
    }
    g_fpel_border = 0;
}


/* ------------------------------------------------------------------ the lookahead stage alone, stock control flow ---- */
/* Steps 1-4 of x264_encoder_encode (encoder.c:3360-3445) without the slice encoding behind them: x264_frame_copy_picture,
 * x264_adaptive_quant_frame, x264_frame_init_lowres, x264_lookahead_put_frame, and -- once the delay is filled --
 * x264_lookahead_get_frames, i.e. the reference's own x264_slicetype_decide / x264_slicetype_analyse / macroblock_tree with its own
 * request order, threads (lookahead-threads) and memoisation.  Every frame the encoder would have picked up next is reported
 * (display index, decided type) and handed back to the frame pool.  luma: n pictures of width*height, chroma = 128.
 * Returns the number of frames decided.  This is bench.py --impl reference for the lookahead workload. */
XREF_API int xref_lookahead_types( void *hv, const uint8_t *luma, int n, int *out_idx, int *out_type )
{
    x264_t *h = hv;
    int w = h->param.i_width, ht = h->param.i_height;
    int cw = ( w + 1 ) / 2, ch = ( ht + 1 ) / 2;
    uint8_t *chroma = malloc( (size_t)cw * ch );
    if( !chroma ) return -1;
    memset( chroma, 128, (size_t)cw * ch );
    int n_out = 0;
    for( int i = 0; i <= n; i++ )
    {
        int flushing = i == n;
        if( !flushing )
        {
            x264_picture_t pic;
            x264_picture_init( &pic );
            pic.img.i_csp = X264_CSP_I420;
            pic.img.i_plane = 3;
            pic.img.plane[0] = (uint8_t*)luma + (size_t)i*w*ht; pic.img.i_stride[0] = w;
            pic.img.plane[1] = chroma; pic.img.i_stride[1] = cw;
            pic.img.plane[2] = chroma; pic.img.i_stride[2] = cw;
            pic.i_pts = i;
            if( xref_forced_types ) pic.i_type = xref_forced_types[i];
            x264_frame_t *fenc = x264_frame_pop_unused( h, 0 );
            if( !fenc || x264_frame_copy_picture( h, fenc, &pic ) < 0 ) { free( chroma ); return -1; }
            if( h->param.i_width != 16 * h->mb.i_mb_width || h->param.i_height != 16 * h->mb.i_mb_height )
                x264_frame_expand_border_mod16( h, fenc );
            fenc->i_frame = h->frames.i_input++;
            fenc->i_pic_struct = PIC_STRUCT_PROGRESSIVE;
            x264_adaptive_quant_frame( h, fenc, NULL );
            if( h->frames.b_have_lowres )
                x264_frame_init_lowres( h, fenc );
            x264_lookahead_put_frame( h, fenc );
            if( h->frames.i_input <= h->frames.i_delay + 1 - h->i_thread_frames )
                continue;
        }
        do
        {
            h->i_frame++;
            if( !h->frames.current[0] )
                x264_lookahead_get_frames( h );
            if( !h->frames.current[0] )
                break;
            x264_frame_t *f = x264_frame_shift( h->frames.current );
            out_idx[n_out] = f->i_frame; out_type[n_out] = f->i_type; n_out++;
            x264_frame_push_unused( h, f );
        } while( flushing );
    }
    free( chroma );
    return n_out;
}


/* The same as xref_encode_i420, also reporting what came out: an FNV-1a hash over every byte of the bitstream, its size, and the
 * display index / type of each coded frame (SEI units excepted: they spell out the options).  Used to show that the encoder's OUTPUT is unchanged when its lookahead runs behind
 * the offload hooks (libx264ref_b200.so, integration/x264_b200_hooks.c). */
XREF_API int xref_encode_i420_hash( void *hv, const uint8_t *yuv, int n, uint64_t *hash, int64_t *bytes, int *out_idx, int *out_type )
{
    x264_t *h = hv;
    int w = h->param.i_width, ht = h->param.i_height;
    int cw = ( w + 1 ) / 2, ch = ( ht + 1 ) / 2;
    size_t fsz = (size_t)w*ht + 2*(size_t)cw*ch;
    int n_out = 0;
    uint64_t hv64 = 1469598103934665603ull;
    int64_t total = 0;
    x264_nal_t *nal; int i_nal;
    x264_picture_t pic_in, pic_out;
    for( int i = 0; i < n || x264_encoder_delayed_frames( h ); i++ )
    {
        int r;
        if( i < n )
        {
            x264_picture_init( &pic_in );
            pic_in.img.i_csp = X264_CSP_I420;
            pic_in.img.i_plane = 3;
            pic_in.img.plane[0] = (uint8_t*)yuv + i*fsz;             pic_in.img.i_stride[0] = w;
            pic_in.img.plane[1] = pic_in.img.plane[0] + (size_t)w*ht; pic_in.img.i_stride[1] = cw;
            pic_in.img.plane[2] = pic_in.img.plane[1] + (size_t)cw*ch; pic_in.img.i_stride[2] = cw;
            pic_in.i_pts = i;
            r = x264_encoder_encode( h, &nal, &i_nal, &pic_in, &pic_out );
        }
        else
            r = x264_encoder_encode( h, &nal, &i_nal, NULL, &pic_out );
        if( r < 0 ) return -1;
        if( r > 0 )
        {
            for( int k = 0; k < i_nal; k++ )
            {   /* the SEI carries the option string ("... opencl=1", x264_param2string, common/base.c:1446): everything but it */
                if( nal[k].i_type == NAL_SEI )
                    continue;
                for( int j = 0; j < nal[k].i_payload; j++ )
                    hv64 = ( hv64 ^ nal[k].p_payload[j] ) * 1099511628211ull;
                total += nal[k].i_payload;
            }
            if( out_idx ) { out_idx[n_out] = (int)pic_out.i_pts; out_type[n_out] = pic_out.i_type; }
            n_out++;
        }
    }
    *hash = hv64; *bytes = total;
    return n_out;
}

/* 1 if the encoder behind hv still has its offload hooks switched on (h->param.b_opencl survives only if the backend came up) */
XREF_API int xref_offload_active( void *hv ) { return ((x264_t*)hv)->param.b_opencl; }


/* x264_frame_copy_picture (common/frame.c:363-480) on a frame of the encoder's own pool: the picture as the reference lays it out
 * internally -- out_luma = plane[0] (width x height, packed), out_uv = plane[1] (interleaved, width x height/2, packed) */
XREF_API int xref_frame_copy_picture( void *hv, int i_csp, uint8_t *const plane[3], const int stride[3], uint8_t *out_luma, uint8_t *out_uv )
{
    x264_t *h = hv;
    tables_init();
    x264_picture_t pic;
    x264_picture_init( &pic );
    pic.img.i_csp = i_csp;
    pic.img.i_plane = 3;
    for( int i = 0; i < 3; i++ ) { pic.img.plane[i] = plane[i]; pic.img.i_stride[i] = stride[i]; }
    x264_frame_t *f = x264_frame_pop_unused( h, 0 );
    if( !f ) return -1;
    int r = x264_frame_copy_picture( h, f, &pic );
    if( !r )
    {
        int w = h->param.i_width, ht = h->param.i_height;
        for( int y = 0; y < ht; y++ ) memcpy( out_luma + (size_t)y*w, f->plane[0] + (intptr_t)y*f->i_stride[0], w );
        for( int y = 0; y < ht/2; y++ ) memcpy( out_uv + (size_t)y*w, f->plane[1] + (intptr_t)y*f->i_stride[1], w );
    }
    x264_frame_push_unused( h, f );
    return r;
}


/* pixf.ads[i_pixel] (common/pixel.c:759-803) */
XREF_API int xref_pixel_ads( int i_pixel, int *enc_dc, uint16_t *sums, int delta, uint16_t *cost_mvx, int16_t *mvs, int width, int thresh )
{
    tables_init();
    return g_pf.ads[i_pixel]( enc_dc, sums, delta, cost_mvx, mvs, width, thresh );
}

/* The integral planes x264_frame_filter leaves for a reference frame (common/mc.c:748-783; the encoder must have been opened with
 * me=esa or tesa): out8 / out4 (out4 only with sub-8x8 partitions, else untouched) receive stride * (lines + 2*PADV) elements each,
 * starting PADV rows and PADH columns before position (0,0).  *stride_out = the frame's stride.  Returns 1 if the 4x4 plane exists. */
XREF_API int xref_frame_integral( void *hv, const uint8_t *luma, intptr_t luma_stride, uint16_t *out8, uint16_t *out4, int *stride_out )
{
    x264_t *h = hv;
    x264_frame_t *f = x264_frame_pop_unused( h, 1 );
    if( !f || !f->integral ) return -1;
    int W = h->mb.i_mb_width*16, H = h->mb.i_mb_height*16;
    for( int y = 0; y < H; y++ )
        memcpy( f->plane[0] + y*f->i_stride[0], luma + y*luma_stride, W );
    f->b_kept_as_ref = 1;
    h->i_threadslice_start = 0;
    h->i_threadslice_end = h->mb.i_mb_height;
    for( int mb_y = 0; mb_y < h->mb.i_mb_height; mb_y++ )
    {
        int end = mb_y == h->mb.i_mb_height - 1;
        x264_frame_expand_border( h, f, mb_y );
        x264_frame_filter( h, f, mb_y, end );
        x264_frame_expand_border_filtered( h, f, mb_y, end );
    }
    intptr_t st = f->i_stride[0];
    size_t n = (size_t)st * ( f->i_lines[0] + 2*PADV );
    memcpy( out8, f->integral - PADV*st - PADH, ( n - 64 ) * sizeof(uint16_t) );
    int sub = h->frames.b_have_sub8x8_esa;
    if( sub ) memcpy( out4, f->integral + n - PADV*st - PADH, ( n - 64 ) * sizeof(uint16_t) );
    *stride_out = st;
    x264_frame_push_unused( h, f );
    return sub;
}
