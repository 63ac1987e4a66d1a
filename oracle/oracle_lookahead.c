/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Scalar restatement of the reference's lowres lookahead
 * cost: slicetype_mb_cost / slicetype_slice_cost / slicetype_frame_cost (encoder/slicetype.c:514-995) with one
 * lookahead thread, the lowres intra predictors it uses (common/predict.c:221-320, :632-880) and the frame
 * set-up of x264_frame_init_lowres (common/mc.c:458-482).  Weighted prediction is passed in explicitly
 * (the reference's x264_weights_analyse, slicetype.c:284-501, stays a host-side float routine).
 *
 * Parity status: PINNED against the compiled reference's slicetype_frame_cost (tests/test_oracle_lookahead.py):
 * every lowres_mvs / lowres_mv_costs / lowres_costs / i_intra_cost / row_satds / cost_est value.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

#define FDEC 32                      /* FDEC_STRIDE, common/common.h:571 */
#define LOWRES_COST_MASK 0x3fff      /* common/frame.h:107-112 */
#define LOWRES_COST_SHIFT 14

static inline int clip3( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }
static inline int imin( int a, int b ) { return a < b ? a : b; }
static inline int imax( int a, int b ) { return a > b ? a : b; }
static inline int clip_u8( int v ) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int median3( int a, int b, int c ) { return imax( imin( a, b ), imin( imax( a, b ), c ) ); }   /* base.h:232 */

/* ---------------------------------------------------------------- frames ---------------------------- */
static int stride_lowres( int width_lowres )          /* common/frame.c:30-36, :125 */
{
    int s = ( width_lowres + 96 + 63 ) & ~63;
    if( !( s & 2047 ) ) s += 64;
    return s;
}

orc_la_frame_t *orc_la_frame_new( const orc_la_params_t *p, const uint8_t *luma, intptr_t luma_stride )
{
    orc_la_frame_t *f = calloc( 1, sizeof( *f ) );
    int W16 = p->mb_width * 16, H16 = p->mb_height * 16;
    f->width_lowres = W16 / 2;
    f->lines_lowres = H16 / 2;
    f->stride_lowres = stride_lowres( f->width_lowres );
    f->mb_count = p->mb_width * p->mb_height;
    f->bframes = p->bframes;
    /* x264_frame_expand_border_mod16, common/frame.c:640-665 */
    uint8_t *src = malloc( (size_t)W16 * H16 );
    for( int y = 0; y < H16; y++ )
    {
        const uint8_t *row = luma + (intptr_t)imin( y, p->height-1 ) * luma_stride;
        memcpy( src + (size_t)y*W16, row, p->width );
        memset( src + (size_t)y*W16 + p->width, row[p->width-1], W16 - p->width );
    }
    size_t plane = (size_t)f->stride_lowres * ( f->lines_lowres + 2*ORC_PAD ) + 64;
    for( int i = 0; i < 4; i++ )
    {
        f->lowres_buf[i] = calloc( 1, plane );
        f->lowres[i] = f->lowres_buf[i] + ORC_PAD * f->stride_lowres + ORC_PAD;
    }
    orc_frame_init_lowres( src, W16, W16, H16, f->lowres, f->stride_lowres, f->width_lowres, f->lines_lowres );
    {   /* frame statistics of x264_adaptive_quant_frame (ratecontrol.c:225-233, :405-414): sum and sum of squares over the
         * mod-16 picture, then ssd = sqr - (sum^2 + N/2)/N */
        uint64_t sum = 0, sqr = 0, N = (uint64_t)W16 * H16;
        for( size_t i = 0; i < (size_t)W16 * H16; i++ ) { sum += src[i]; sqr += (uint64_t)src[i] * src[i]; }
        f->pixel_sum = sum;
        f->pixel_ssd = sqr - ( sum * sum + N / 2 ) / N;
    }
    free( src );
    for( int l = 0; l < 2; l++ )
        for( int d = 0; d <= p->bframes; d++ )
        {
            f->lowres_mvs[l][d] = calloc( f->mb_count, 4 );              /* zeroed like frame.c:287-293 */
            f->lowres_mv_costs[l][d] = calloc( f->mb_count, sizeof(int) );
            f->lowres_mvs[l][d][0][0] = 0x7FFF;                           /* mc.c:478-480 */
        }
    for( int i = 0; i < p->bframes+2; i++ )
        for( int j = 0; j < p->bframes+2; j++ )
        {
            f->lowres_costs[i][j] = calloc( f->mb_count, sizeof(uint16_t) );
            f->row_satds[i][j] = calloc( p->mb_height, sizeof(int) );
            f->row_satds[i][j][0] = -1;
        }
    memset( f->cost_est, -1, sizeof( f->cost_est ) );
    memset( f->cost_est_aq, -1, sizeof( f->cost_est_aq ) );
    f->intra_cost = calloc( f->mb_count, sizeof(int) );
    for( int i = 0; i < f->mb_count; i++ ) f->intra_cost[i] = 0xFFFF;     /* memset( i_intra_cost, -1 ), frame.c:288 */
    f->inv_qscale_factor = malloc( f->mb_count * sizeof(uint16_t) );
    f->propagate_cost = calloc( f->mb_count, sizeof(uint16_t) );
    f->qp_offset = calloc( f->mb_count, sizeof(float) );
    f->qp_offset_aq = calloc( f->mb_count, sizeof(float) );
    f->mb_width = p->mb_width;
    for( int i = 0; i < f->mb_count; i++ ) f->inv_qscale_factor[i] = 256;
    return f;
}

void orc_la_frame_delete( orc_la_frame_t *f )
{
    if( !f ) return;
    for( int i = 0; i < 4; i++ ) free( f->lowres_buf[i] );
    for( int l = 0; l < 2; l++ )
        for( int d = 0; d <= f->bframes; d++ ) { free( f->lowres_mvs[l][d] ); free( f->lowres_mv_costs[l][d] ); }
    for( int i = 0; i < f->bframes+2; i++ )
        for( int j = 0; j < f->bframes+2; j++ ) { free( f->lowres_costs[i][j] ); free( f->row_satds[i][j] ); }
    free( f->intra_cost );
    free( f->inv_qscale_factor );
    free( f->weighted_buf );
    free( f->propagate_cost ); free( f->qp_offset ); free( f->qp_offset_aq );
    free( f );
}

/* accessors for the ctypes tests */
void orc_la_frame_get( orc_la_frame_t *f, int what, int i, int j, void *out )
{
    switch( what )
    {
        case 0: memcpy( out, f->lowres_mvs[i][j], f->mb_count * 4 ); break;
        case 1: memcpy( out, f->lowres_mv_costs[i][j], f->mb_count * sizeof(int) ); break;
        case 2: memcpy( out, f->lowres_costs[i][j], f->mb_count * 2 ); break;
        case 3: memcpy( out, f->intra_cost, f->mb_count * sizeof(int) ); break;
        case 4: ((int*)out)[0] = f->cost_est[i][j]; ((int*)out)[1] = f->cost_est_aq[i][j]; ((int*)out)[2] = f->intra_mbs[i]; break;
        case 5: memcpy( out, f->row_satds[i][j], ( f->lines_lowres / 8 ) * sizeof(int) ); break;
        case 6: ((int*)out)[0] = f->weight.enabled; ((int*)out)[1] = f->weight.scale; ((int*)out)[2] = f->weight.denom; ((int*)out)[3] = f->weight.offset; break;
    }
}
void orc_la_frame_set_qscale( orc_la_frame_t *f, const uint16_t *inv_qscale ) { memcpy( f->inv_qscale_factor, inv_qscale, f->mb_count * 2 ); }
uint8_t *orc_la_frame_plane( orc_la_frame_t *f, int i ) { return f->lowres_buf[i]; }
int orc_la_frame_stride( orc_la_frame_t *f ) { return (int)f->stride_lowres; }

/* ---------------------------------------------------------------- lowres intra ---------------------- */
/* pix: 8x8 block at stride FDEC with the top row (16 px), left column and corner in place (slicetype.c:722-727) */
static void pred8x8c_dc( uint8_t *s )                                   /* predict.c:221-257 */
{
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for( int i = 0; i < 4; i++ )
    {
        s0 += s[i - FDEC]; s1 += s[i + 4 - FDEC];
        s2 += s[-1 + i*FDEC]; s3 += s[-1 + (i+4)*FDEC];
    }
    int dc[4] = { ( s0 + s2 + 4 ) >> 3, ( s1 + 2 ) >> 2, ( s3 + 2 ) >> 2, ( s1 + s3 + 4 ) >> 3 };
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
            s[y*FDEC+x] = dc[( y >> 2 )*2 + ( x >> 2 )];
}
static void pred8x8c_h( uint8_t *s ) { for( int y = 0; y < 8; y++ ) memset( s + y*FDEC, s[y*FDEC-1], 8 ); }
static void pred8x8c_v( uint8_t *s ) { for( int y = 0; y < 8; y++ ) memcpy( s + y*FDEC, s - FDEC, 8 ); }
static void pred8x8c_p( uint8_t *s )                                    /* predict.c:282-310 */
{
    int H = 0, V = 0;
    for( int i = 0; i < 4; i++ )
    {
        H += ( i + 1 ) * ( s[4+i - FDEC] - s[2-i - FDEC] );
        V += ( i + 1 ) * ( s[-1 + (i+4)*FDEC] - s[-1 + (2-i)*FDEC] );
    }
    int a = 16 * ( s[-1 + 7*FDEC] + s[7 - FDEC] );
    int b = ( 17*H + 16 ) >> 5, c = ( 17*V + 16 ) >> 5;
    int i00 = a - 3*b - 3*c + 16;
    for( int y = 0; y < 8; y++, i00 += c )
        for( int x = 0, pix = i00; x < 8; x++, pix += b )
            s[y*FDEC+x] = clip_u8( pix >> 5 );
}

#define F1(a,b)   (((a)+(b)+1)>>1)
#define F2(a,b,c) (((a)+2*(b)+(c)+2)>>2)
/* low-pass filtered neighbours, all neighbours available: predict.c:632-680.  l[-1] == t[-1] == corner. */
typedef struct { int lbuf[9], tbuf[17]; } edge_t;
#define EL(e,i) ((e)->lbuf[(i)+1])
#define ET(e,i) ((e)->tbuf[(i)+1])
static void edge_filter( const uint8_t *s, edge_t *e )
{
#define S(x,y) ((int)s[(y)*FDEC+(x)])
    EL(e,-1) = ET(e,-1) = ( S(0,-1) + 2*S(-1,-1) + S(-1,0) + 2 ) >> 2;
    EL(e,0) = ( S(-1,-1) + 2*S(-1,0) + S(-1,1) + 2 ) >> 2;
    for( int y = 1; y < 7; y++ ) EL(e,y) = F2( S(-1,y-1), S(-1,y), S(-1,y+1) );
    EL(e,7) = ( S(-1,6) + 3*S(-1,7) + 2 ) >> 2;
    ET(e,0) = ( S(-1,-1) + 2*S(0,-1) + S(1,-1) + 2 ) >> 2;
    for( int x = 1; x < 15; x++ ) ET(e,x) = F2( S(x-1,-1), S(x,-1), S(x+1,-1) );
    ET(e,15) = ( S(14,-1) + 3*S(15,-1) + 2 ) >> 2;
#undef S
}
/* modes 3..8 of predict_8x8 (I_PRED_8x8_DDL, DDR, VR, HD, VL, HU), predict.c:741-880, as closed forms */
static void pred8x8_dir( uint8_t *s, const edge_t *e, int mode )
{
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int v;
            switch( mode )
            {
            case 3: /* DDL */
                v = ( x == 7 && y == 7 ) ? F2( ET(e,14), ET(e,15), ET(e,15) ) : F2( ET(e,x+y), ET(e,x+y+1), ET(e,x+y+2) );
                break;
            case 4: /* DDR */
            {
                int d = x - y;
                v = d > 0 ? F2( ET(e,d-2), ET(e,d-1), ET(e,d) )
                  : d == 0 ? F2( EL(e,0), ET(e,-1), ET(e,0) )
                  : F2( EL(e,-d), EL(e,-d-1), EL(e,-d-2) );
                break;
            }
            case 5: /* VR */
            {
                int z = 2*x - y;
                if( z >= 0 )
                    v = ( z & 1 ) ? F2( ET(e,x-(y>>1)-2), ET(e,x-(y>>1)-1), ET(e,x-(y>>1)) )
                                  : F1( ET(e,x-(y>>1)-1), ET(e,x-(y>>1)) );
                else if( z == -1 )
                    v = F2( EL(e,0), ET(e,-1), ET(e,0) );
                else
                    v = F2( EL(e,y-2*x-1), EL(e,y-2*x-2), EL(e,y-2*x-3) );
                break;
            }
            case 6: /* HD: value = q[2*(7-y)+x] over the sequence of (F1,F2) pairs walking up the left edge and
                     * then along the top (predict.c:824-851) */
            {
                int i = 2*( 7 - y ) + x;
                if( i == 15 )
                    v = F2( EL(e,0), ET(e,-1), ET(e,0) );
                else if( i < 16 )
                {
                    int k = 6 - ( i >> 1 );                  /* pair k uses l[k], l[k+1] (k = -1 is the corner) */
                    v = ( i & 1 ) ? F2( EL(e,k-1), EL(e,k), EL(e,k+1) ) : F1( EL(e,k), EL(e,k+1) );
                }
                else
                {
                    int k = i - 16;                          /* F2(t[k+1], t[k], t[k-1]) */
                    v = F2( ET(e,k+1), ET(e,k), ET(e,k-1) );
                }
                break;
            }
            case 7: /* VL */
                v = ( y & 1 ) ? F2( ET(e,x+(y>>1)), ET(e,x+(y>>1)+1), ET(e,x+(y>>1)+2) )
                              : F1( ET(e,x+(y>>1)), ET(e,x+(y>>1)+1) );
                break;
            default: /* 8: HU, value = q[2*y+x] (predict.c:862-880) */
            {
                int i = 2*y + x, k = i >> 1;
                if( i >= 14 ) v = EL(e,7);
                else if( i == 13 ) v = F2( EL(e,6), EL(e,7), EL(e,7) );
                else v = ( i & 1 ) ? F2( EL(e,k), EL(e,k+1), EL(e,k+2) ) : F1( EL(e,k), EL(e,k+1) );
                break;
            }
            }
            s[y*FDEC+x] = v;
        }
}

/* ---------------------------------------------------------------- per-MB cost ----------------------- */
typedef struct
{
    const orc_la_params_t *p;
    const uint16_t *cost_mv;
    orc_la_frame_t **frames;
    int p0, p1, b, dist_scale_factor;
    int do_search[2];
    const orc_weight_t *w;
    const uint8_t *weighted_plane;          /* fenc->weighted[0] origin when w->enabled */
    int lambda;
    /* accumulators (one lookahead thread): cost_est, cost_est_aq, intra_mbs + row sums */
    int inter[3], intra[3];
    int *row_inter, *row_intra;
} la_t;

static int mbcmp8x8( const la_t *L, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    return L->p->subpel_refine > 1 ? orc_satd( a, sa, b, sb, 8, 8 ) : orc_sad( a, sa, b, sb, 8, 8 );   /* encoder.c:1409-1427 */
}

static void la_mb_cost( la_t *L, int mb_x, int mb_y )
{
    const orc_la_params_t *p = L->p;
    orc_la_frame_t *fref0 = L->frames[L->p0], *fref1 = L->frames[L->p1], *fenc = L->frames[L->b];
    const int b = L->b, p0 = L->p0, p1 = L->p1;
    const int b_bidir = b < p1;
    const int mb_stride = p->mb_width, mb_xy = mb_x + mb_y * mb_stride;
    const intptr_t stride = fenc->stride_lowres;
    const intptr_t pel = 8 * ( mb_x + mb_y * stride );
    const int bipred_weight = p->weighted_bipred ? 64 - ( L->dist_scale_factor >> 2 ) : 32;
    const int b_frame_score_mb = ( mb_x > 0 && mb_x < p->mb_width-1 && mb_y > 0 && mb_y < p->mb_height-1 )
                                 || p->mb_width <= 2 || p->mb_height <= 2;
    const uint8_t *fenc_blk = fenc->lowres[0] + pel;
    int16_t (*fenc_mvs[2])[2] = { b != p0 ? &fenc->lowres_mvs[0][b-p0-1][mb_xy] : NULL,
                                  b != p1 ? &fenc->lowres_mvs[1][p1-b-1][mb_xy] : NULL };
    int *fenc_costs[2] = { b != p0 ? &fenc->lowres_mv_costs[0][b-p0-1][mb_xy] : NULL,
                           b != p1 ? &fenc->lowres_mv_costs[1][p1-b-1][mb_xy] : NULL };
    uint8_t pix1[16*16], pix2[16*16];
    int bcost = ORC_COST_MAX, list_used = 0;
    const int lowres_penalty = 4;
    orc_me_t m[2];
    orc_me_ctx_t c;
    memset( &c, 0, sizeof( c ) );          /* no chroma ME in the lookahead (slicetype.c:60) */
    memset( m, 0, sizeof( m ) );

    if( p0 != p1 )
    {
        int mv_range = 2 * p->mv_range;                                          /* slicetype.c:550-562 */
        c.mv_min_spel[0] = imax( 4*( -8*mb_x - 12 ), -mv_range );
        c.mv_max_spel[0] = imin( 4*( 8*( p->mb_width - mb_x - 1 ) + 12 ), mv_range-1 );
        c.mv_min_spel[1] = imax( 4*( -8*mb_y - 12 ), -mv_range );
        c.mv_max_spel[1] = imin( 4*( 8*( p->mb_height - mb_y - 1 ) + 12 ), mv_range-1 );
        for( int i = 0; i < 2; i++ )
        {
            c.mv_limit_fpel[0][i] = c.mv_min_spel[i] >> 2;
            c.mv_limit_fpel[1][i] = c.mv_max_spel[i] >> 2;
        }
        c.me_range = p->me_range;
        c.mbcmp_is_satd = p->subpel_refine > 1;
        if( p->subpel_refine > 1 ) { c.me_method = imin( ORC_ME_HEX, p->me_method ); c.subpel_refine = 4; }   /* slicetype.c:45-61 */
        else                       { c.me_method = ORC_ME_DIA; c.subpel_refine = 2; }

        for( int l = 0; l < 2; l++ )
        {
            orc_la_frame_t *fr = l ? fref1 : fref0;
            m[l].i_pixel = ORC_PIXEL_8x8;
            m[l].p_cost_mv = L->cost_mv;
            m[l].stride = stride;
            m[l].p_fenc = fenc_blk;
            m[l].fenc_stride = stride;
            for( int i = 0; i < 4; i++ ) m[l].p_fref[i] = fr->lowres[i] + pel;
            m[l].p_fref_w = m[l].p_fref[0];
        }
        if( L->w && L->w->enabled )
        {
            m[0].weight = *L->w;
            m[0].p_fref_w = L->weighted_plane + pel;
        }

#define TRY_BIDIR( mv0, mv1, penalty ) \
        { \
            if( p->subpel_refine <= 1 ) \
            {   /* half-pel planes addressed directly, slicetype.c:589-596 */ \
                int h1 = ( ( (mv0)[0] & 2 ) >> 1 ) + ( (mv0)[1] & 2 ), h2 = ( ( (mv1)[0] & 2 ) >> 1 ) + ( (mv1)[1] & 2 ); \
                const uint8_t *s1 = m[0].p_fref[h1] + ( (mv0)[0] >> 2 ) + ( (mv0)[1] >> 2 ) * stride; \
                const uint8_t *s2 = m[1].p_fref[h2] + ( (mv1)[0] >> 2 ) + ( (mv1)[1] >> 2 ) * stride; \
                orc_pixel_avg( pix1, 16, s1, stride, s2, stride, 8, 8, bipred_weight ); \
            } \
            else \
            { \
                orc_mc_luma( pix1, 16, m[0].p_fref, stride, (mv0)[0], (mv0)[1], 8, 8, L->w ); \
                orc_mc_luma( pix2, 16, m[1].p_fref, stride, (mv1)[0], (mv1)[1], 8, 8, L->w ); \
                orc_pixel_avg( pix1, 16, pix1, 16, pix2, 16, 8, 8, bipred_weight ); \
            } \
            int i_cost = (penalty) * L->lambda + mbcmp8x8( L, fenc_blk, stride, pix1, 16 ); \
            if( i_cost < bcost ) { bcost = i_cost; list_used = 3; } \
        }

        if( b_bidir )
        {
            int16_t dmv[2][2] = { {0,0}, {0,0} };
            if( fref1->lowres_mvs[0][p1-p0-1][0][0] != 0x7FFF )
            {   /* temporal direct from the co-located L0 vector of the later reference, slicetype.c:629-642 */
                int16_t *mvr = fref1->lowres_mvs[0][p1-p0-1][mb_xy];
                dmv[0][0] = ( mvr[0] * L->dist_scale_factor + 128 ) >> 8;
                dmv[0][1] = ( mvr[1] * L->dist_scale_factor + 128 ) >> 8;
                dmv[1][0] = dmv[0][0] - mvr[0];
                dmv[1][1] = dmv[0][1] - mvr[1];
                for( int l = 0; l < 2; l++ )
                    for( int i = 0; i < 2; i++ )
                        dmv[l][i] = clip3( dmv[l][i], c.mv_min_spel[i], c.mv_max_spel[i] );
                if( p->subpel_refine <= 1 )
                    for( int l = 0; l < 2; l++ )
                        for( int i = 0; i < 2; i++ )
                            dmv[l][i] &= ~1;
            }
            TRY_BIDIR( dmv[0], dmv[1], 0 );
            if( dmv[0][0] | dmv[0][1] | dmv[1][0] | dmv[1][1] )
            {
                orc_pixel_avg( pix1, 16, m[0].p_fref[0], stride, m[1].p_fref[0], stride, 8, 8, bipred_weight );
                int i_cost = mbcmp8x8( L, fenc_blk, stride, pix1, 16 );
                if( i_cost < bcost ) { bcost = i_cost; list_used = 3; }
            }
        }

        for( int l = 0; l < 1 + b_bidir; l++ )
        {
            if( L->do_search[l] )
            {
                int i_mvc = 0;
                int16_t (*fenc_mv)[2] = fenc_mvs[l];
                int16_t mvc[4][2] = { {0,0}, {0,0}, {0,0}, {0,0} };
                /* reverse-order predictors: right, below, below-left, below-right (slicetype.c:662-675) */
#define MVC(mv) { mvc[i_mvc][0] = (mv)[0]; mvc[i_mvc][1] = (mv)[1]; i_mvc++; }
                if( mb_x < p->mb_width - 1 )
                    MVC( fenc_mv[1] );
                if( mb_y < p->mb_height - 1 )
                {
                    MVC( fenc_mv[mb_stride] );
                    if( mb_x > 0 )
                        MVC( fenc_mv[mb_stride-1] );
                    if( mb_x < p->mb_width - 1 )
                        MVC( fenc_mv[mb_stride+1] );
                }
#undef MVC
                if( i_mvc <= 1 ) { m[l].mvp[0] = mvc[0][0]; m[l].mvp[1] = mvc[0][1]; }
                else
                {
                    m[l].mvp[0] = median3( mvc[0][0], mvc[1][0], mvc[2][0] );
                    m[l].mvp[1] = median3( mvc[0][1], mvc[1][1], mvc[2][1] );
                }
                int skip = 0;
                if( !( m[l].mvp[0] | m[l].mvp[1] ) )
                {   /* zero-predictor fast skip, slicetype.c:684-692 */
                    m[l].cost = mbcmp8x8( L, fenc_blk, stride, m[l].p_fref[0], stride );
                    if( m[l].cost < 64 )
                    {
                        m[l].mv[0] = m[l].mv[1] = 0;
                        skip = 1;
                    }
                }
                if( !skip )
                {
                    orc_me_search_ref( &c, &m[l], (const int16_t (*)[2])mvc, i_mvc, NULL );
                    m[l].cost -= L->cost_mv[0];
                    if( m[l].mv[0] | m[l].mv[1] )
                        m[l].cost += 5 * L->lambda;
                }
                (*fenc_mvs[l])[0] = m[l].mv[0]; (*fenc_mvs[l])[1] = m[l].mv[1];
                *fenc_costs[l] = m[l].cost;
            }
            else
            {
                m[l].mv[0] = (*fenc_mvs[l])[0]; m[l].mv[1] = (*fenc_mvs[l])[1];
                m[l].cost = *fenc_costs[l];
            }
            if( m[l].cost < bcost ) { bcost = m[l].cost; list_used = l+1; }
        }

        if( b_bidir && ( m[0].mv[0] | m[0].mv[1] | m[1].mv[0] | m[1].mv[1] ) )
            TRY_BIDIR( m[0].mv, m[1].mv, 5 );
#undef TRY_BIDIR
    }

    /* lowres intra, slicetype.c:714-757 */
    if( !fenc->b_intra_calculated )
    {
        uint8_t buf[10*FDEC];
        uint8_t *pix = buf + FDEC + 8;
        const uint8_t *src = fenc_blk;
        memcpy( pix - FDEC, src - stride, 16 );
        for( int i = -1; i < 8; i++ )
            memcpy( pix + i*FDEC - 4, src + i*stride - 4, 4 );
        int icost = ORC_COST_MAX, satd;
        pred8x8c_dc( pix ); satd = mbcmp8x8( L, pix, FDEC, fenc_blk, stride ); icost = imin( icost, satd );
        pred8x8c_h( pix );  satd = mbcmp8x8( L, pix, FDEC, fenc_blk, stride ); icost = imin( icost, satd );
        pred8x8c_v( pix );  satd = mbcmp8x8( L, pix, FDEC, fenc_blk, stride ); icost = imin( icost, satd );
        if( p->subpel_refine > 1 )
        {
            pred8x8c_p( pix ); satd = mbcmp8x8( L, fenc_blk, stride, pix, FDEC ); icost = imin( icost, satd );
            edge_t e;
            edge_filter( pix, &e );
            for( int mode = 3; mode < 9; mode++ )
            {
                pred8x8_dir( pix, &e, mode );
                satd = mbcmp8x8( L, fenc_blk, stride, pix, FDEC );
                icost = imin( icost, satd );
            }
        }
        icost = icost + 5 * L->lambda + lowres_penalty;
        fenc->intra_cost[mb_xy] = icost;
        int icost_aq = icost;
        if( p->aq_mode )
            icost_aq = ( icost_aq * fenc->inv_qscale_factor[mb_xy] + 128 ) >> 8;
        L->row_intra[mb_y] += icost_aq;
        if( b_frame_score_mb )
        {
            L->intra[0] += icost;
            L->intra[1] += icost_aq;
        }
    }
    bcost += lowres_penalty;

    if( !b_bidir )
    {   /* intra MBs are only allowed in P frames, slicetype.c:762-773 */
        int icost = fenc->intra_cost[mb_xy];
        int b_intra = icost < bcost;
        if( b_intra ) { bcost = icost; list_used = 0; }
        if( b_frame_score_mb )
            L->inter[2] += b_intra;
    }
    if( p0 != p1 )
    {
        int bcost_aq = bcost;
        if( p->aq_mode )
            bcost_aq = ( bcost_aq * fenc->inv_qscale_factor[mb_xy] + 128 ) >> 8;
        L->row_inter[mb_y] += bcost_aq;
        if( b_frame_score_mb )
        {
            L->inter[0] += bcost;
            L->inter[1] += bcost_aq;
        }
    }
    fenc->lowres_costs[b-p0][p1-b][mb_xy] = imin( bcost, LOWRES_COST_MASK ) + ( list_used << LOWRES_COST_SHIFT );
    if( p0 == p1 )   /* i_intra_cost IS lowres_costs[0][0] (frame.c:287): an I request leaves the clipped value behind */
        fenc->intra_cost[mb_xy] = fenc->lowres_costs[0][0][mb_xy];
}

/* ---------------------------------------------------------------- lookahead weightp ---------------- */
static int ue_bits( unsigned v ) { int n = 0; v++; while( v >> ( n + 1 ) ) n++; return 2*n + 1; }      /* bs_size_ue, bitstream.h:278 */
static int se_bits( int v ) { int t = 1 - 2*v; if( t < 0 ) t = 2*v; int n = 0; while( t >> ( n + 1 ) ) n++; return 2*n + 1; }   /* bs_size_se */

/* weight_cost_luma, slicetype.c:191-222: per 8x8 block min( mbcmp( weighted ref, fenc ), intra cost ) + header bits */
static unsigned weight_cost_luma( const orc_la_params_t *p, orc_la_frame_t *fenc, const uint8_t *src, const orc_weight_t *w )
{
    unsigned cost = 0;
    intptr_t stride = fenc->stride_lowres;
    int i_mb = 0;
    uint8_t buf[8*8];
    for( int y = 0; y < fenc->lines_lowres; y += 8 )
        for( int x = 0; x < fenc->width_lowres; x += 8, i_mb++ )
        {
            const uint8_t *s = src + y*stride + x, *f = fenc->lowres[0] + y*stride + x;
            int cmp;
            if( w )
            {
                orc_mc_weight( buf, 8, s, stride, w, 8, 8 );
                cmp = p->subpel_refine > 1 ? orc_satd( buf, 8, f, stride, 8, 8 ) : orc_sad( buf, 8, f, stride, 8, 8 );
            }
            else
                cmp = p->subpel_refine > 1 ? orc_satd( s, stride, f, stride, 8, 8 ) : orc_sad( s, stride, f, stride, 8, 8 );
            /* i_intra_cost aliases the u16 lowres_costs[0][0] array (frame.c:287) */
            int icost = (uint16_t)fenc->intra_cost[i_mb];
            cost += cmp < icost ? cmp : icost;
        }
    if( w )   /* weight_slice_header_cost, slicetype.c:170-189, one slice, lambda 1 */
        cost += 10 + ue_bits( w->denom ) * 2 + 2 * ( se_bits( w->scale ) + se_bits( w->offset ) );
    return cost;
}

int orc_la_frame_cost( const orc_la_params_t *p, const uint16_t *cost_mv_centre, orc_la_frame_t **frames, int p0, int p1, int b );

/* x264_weights_analyse with b_lookahead = 1, slicetype.c:284-501 (luma only).  Leaves fenc->weight / weighted_buf. */
static void la_weights_analyse( const orc_la_params_t *p, const uint16_t *cost_mv, orc_la_frame_t *fenc, orc_la_frame_t *ref, int delta_index )
{
    const float epsilon = 1.f / 128.f;
    orc_weight_t none = { 0, 1, 0, 0 };
    fenc->weight = none;
    int zero_bias = !ref->pixel_ssd;
    float fenc_var = fenc->pixel_ssd + zero_bias, ref_var = ref->pixel_ssd + zero_bias;
    float guess_scale = sqrtf( fenc_var / ref_var );
    float npix = (float)( ( fenc->lines_lowres * 2 ) * ( fenc->width_lowres * 2 ) );
    float fenc_mean = (float)( fenc->pixel_sum + zero_bias ) / ( ( fenc->lines_lowres * 2 ) * ( fenc->width_lowres * 2 ) );
    float ref_mean  = (float)( ref->pixel_sum + zero_bias ) / ( ( fenc->lines_lowres * 2 ) * ( fenc->width_lowres * 2 ) );
    (void)npix;
    if( fabsf( ref_mean - fenc_mean ) < 0.5f && fabsf( 1.f - guess_scale ) < epsilon )
        return;
    /* weight_get_h264( round( guess_scale * 128 ), 0 ), slicetype.c:64-75 */
    int denom = 7, scale = (int)round( guess_scale * 128 );
    while( denom > 0 && scale > 127 ) { denom--; scale >>= 1; }
    if( scale > 127 ) scale = 127;
    int mindenom = denom, minscale = scale, minoff = 0, found = 0;
    if( !fenc->b_intra_calculated )
    {
        orc_la_frame_t *one[1] = { fenc };
        orc_la_frame_cost( p, cost_mv, one, 0, 0, 0 );
    }
    const uint8_t *mcbuf = ref->lowres[0];
    unsigned origscore, minscore;
    origscore = minscore = weight_cost_luma( p, fenc, mcbuf, NULL );
    if( !minscore )
        return;
    {   /* scale_dist = offset_dist = 0 in the lookahead: exactly one (scale, offset) pair is tried */
        int cur_scale = minscale;
        int cur_offset = fenc_mean - ref_mean * cur_scale / ( 1 << mindenom ) + 0.5f;
        if( cur_offset < -128 || cur_offset > 127 )
        {
            cur_offset = clip3( cur_offset, -128, 127 );
            double v = ( 1 << mindenom ) * ( fenc_mean - cur_offset ) / ref_mean + 0.5f;
            cur_scale = (int)( v < 0 ? 0 : v > 127 ? 127 : v );
        }
        orc_weight_t w = { 1, cur_scale, mindenom, cur_offset };
        unsigned s = weight_cost_luma( p, fenc, mcbuf, &w );
        if( s < minscore ) { minscore = s; minscale = cur_scale; minoff = cur_offset; found = 1; }
    }
    while( mindenom > 0 && !( minscale & 1 ) ) { mindenom--; minscale >>= 1; }
    if( !found || ( minscale == 1 << mindenom && minoff == 0 ) || (float)minscore / origscore > 0.998f )
        return;
    fenc->weight.enabled = 1; fenc->weight.scale = minscale; fenc->weight.denom = mindenom; fenc->weight.offset = minoff;
    if( p->weighted_pred < 0 )                                  /* X264_WEIGHTP_FAKE, slicetype.c:462-463 */
        fenc->weighted_cost_delta[delta_index] = (float)minscore / origscore;
    /* x264_weight_scale_plane over the whole padded reference plane, slicetype.c:489-500 */
    size_t plane = (size_t)ref->stride_lowres * ( ref->lines_lowres + 2*ORC_PAD );
    free( fenc->weighted_buf );
    fenc->weighted_buf = malloc( plane + 64 );
    orc_mc_weight( fenc->weighted_buf, ref->stride_lowres, ref->lowres_buf[0], ref->stride_lowres, &fenc->weight,
                   (int)ref->stride_lowres, ref->lines_lowres + 2*ORC_PAD );
}

/* slicetype_frame_cost with one lookahead thread, encoder/slicetype.c:836-995.  `w` may be NULL. */
int orc_la_frame_cost_w( const orc_la_params_t *p, const uint16_t *cost_mv_centre, orc_la_frame_t **frames,
                         int p0, int p1, int b, const orc_weight_t *w, const uint8_t *weighted_plane )
{
    orc_la_frame_t *fenc = frames[b];
    if( fenc->cost_est[b-p0][p1-b] >= 0 && ( !p->vbv || fenc->row_satds[b-p0][p1-b][0] != -1 ) )
        return fenc->cost_est[b-p0][p1-b];

    la_t L;
    memset( &L, 0, sizeof( L ) );
    L.p = p; L.cost_mv = cost_mv_centre; L.frames = frames; L.p0 = p0; L.p1 = p1; L.b = b;
    L.lambda = 1;                                                    /* x264_lambda_tab[X264_LOOKAHEAD_QP=12] */
    L.dist_scale_factor = 128;
    L.do_search[0] = b != p0 && fenc->lowres_mvs[0][b-p0-1][0][0] == 0x7FFF;
    L.do_search[1] = b != p1 && fenc->lowres_mvs[1][p1-b-1][0][0] == 0x7FFF;
    if( L.do_search[0] )
    {
        if( w && w->enabled && b == p1 ) { L.w = w; L.weighted_plane = weighted_plane; }
        else if( p->weighted_pred && b == p1 )
        {   /* slicetype.c:857-864 */
            la_weights_analyse( p, cost_mv_centre, fenc, frames[p0], b - p0 - 1 );
            if( fenc->weight.enabled )
            {
                L.w = &fenc->weight;
                L.weighted_plane = fenc->weighted_buf + ORC_PAD * fenc->stride_lowres + ORC_PAD;
            }
        }
        fenc->lowres_mvs[0][b-p0-1][0][0] = 0;
    }
    if( L.do_search[1] ) fenc->lowres_mvs[1][p1-b-1][0][0] = 0;
    if( p1 != p0 )
        L.dist_scale_factor = ( ( ( b-p0 ) << 8 ) + ( ( p1-p0 ) >> 1 ) ) / ( p1-p0 );

    L.row_inter = calloc( p->mb_height, sizeof(int) );
    L.row_intra = calloc( p->mb_height, sizeof(int) );

    /* slicetype_slice_cost, slicetype.c:814-834: reverse raster, edges only when someone needs them */
    int do_edges = p->do_edges || p->mb_width <= 2 || p->mb_height <= 2;
    int start_y = imin( p->mb_height - 1, p->mb_height - 2 + do_edges ), end_y = imax( 0, 1 - do_edges );
    int start_x = p->mb_width - 2 + do_edges, end_x = 1 - do_edges;
    for( int y = start_y; y >= end_y; y-- )
        for( int x = start_x; x >= end_x; x-- )
            la_mb_cost( &L, x, y );

    /* accumulator hand-over in the reference's order (slicetype.c:946-966): for an I frame [b-p0][p1-b] IS [0][0] */
    if( b == p1 )
        fenc->intra_mbs[b-p0] = L.inter[2];
    if( !fenc->b_intra_calculated )
    {
        fenc->cost_est[0][0] = 0;
        fenc->cost_est_aq[0][0] = 0;
    }
    fenc->cost_est[b-p0][p1-b] = 0;
    fenc->cost_est_aq[b-p0][p1-b] = 0;
    if( !fenc->b_intra_calculated )
    {
        fenc->cost_est[0][0] += L.intra[0];
        fenc->cost_est_aq[0][0] += L.intra[1];
    }
    fenc->cost_est[b-p0][p1-b] += L.inter[0];
    fenc->cost_est_aq[b-p0][p1-b] += L.inter[1];
    if( p->vbv )
    {
        memcpy( fenc->row_satds[b-p0][p1-b], L.row_inter, p->mb_height * sizeof(int) );
        if( !fenc->b_intra_calculated )
            memcpy( fenc->row_satds[0][0], L.row_intra, p->mb_height * sizeof(int) );
    }
    int score = fenc->cost_est[b-p0][p1-b];
    if( b != p1 )
        score = (int)( (uint64_t)score * 100 / ( 120 + p->bframe_bias ) );
    else
        fenc->b_intra_calculated = 1;
    fenc->cost_est[b-p0][p1-b] = score;
    free( L.row_inter );
    free( L.row_intra );
    return score;
}

int orc_la_frame_cost( const orc_la_params_t *p, const uint16_t *cost_mv_centre, orc_la_frame_t **frames, int p0, int p1, int b )
{
    return orc_la_frame_cost_w( p, cost_mv_centre, frames, p0, p1, b, NULL, NULL );
}

/* ---------------------------------------------------------------- MB-tree ---------------------------- */
/* x264_log2 (common/base.h:226-230): table look-up, 7 mantissa bits.  The table is the reference's x264_log2_lut
 * (common/tables.c:66-85), whose entries are log2(1 + i/128) printed with five decimals */
static const float *log2_lut( void )
{
    static float lut[128];
    static int init = 0;
    if( !init )
    {
        for( int i = 0; i < 128; i++ )
        {
            char buf[32];
            snprintf( buf, sizeof( buf ), "%.5f", log2( 1.0 + i / 128.0 ) );
            lut[i] = strtof( buf, NULL );
        }
        init = 1;
    }
    return lut;
}
/* the two halves of x264_log2: mantissa table entry and integer part */
float orc_log2_frac( uint32_t x ) { int lz = __builtin_clz( x ); return log2_lut()[( x << lz >> 24 ) & 0x7f]; }
#define log2_frac orc_log2_frac
float orc_log2( uint32_t x ) { return log2_frac( x ) + (float)( 31 - __builtin_clz( x ) ); }
static float log2_int( uint32_t x ) { return (float)( 31 - __builtin_clz( x ) ); }

void orc_la_mbtree_reset( orc_la_frame_t *f ) { memset( f->propagate_cost, 0, f->mb_count * sizeof(uint16_t) ); }
void orc_la_frame_set_qp_offset_aq( orc_la_frame_t *f, const float *aq )
{
    memcpy( f->qp_offset_aq, aq, f->mb_count * sizeof(float) );
    memcpy( f->qp_offset, aq, f->mb_count * sizeof(float) );
}
void orc_la_frame_get_mbtree( orc_la_frame_t *f, int what, int i, void *out )
{
    switch( what )
    {
        case 0: memcpy( out, f->qp_offset, f->mb_count * sizeof(float) ); break;
        case 1: memcpy( out, f->qp_offset_aq, f->mb_count * sizeof(float) ); break;
        case 2: memcpy( out, f->propagate_cost, f->mb_count * sizeof(uint16_t) ); break;
        case 3: *(float*)out = f->weighted_cost_delta[i]; break;
    }
}

static void clip_add( uint16_t *s, int x ) { int v = *s + x; *s = v > 32767 ? 32767 : v; }       /* MC_CLIP_ADD, common/mc.h:29: (1<<15)-1 */

void orc_la_mbtree_propagate( const orc_la_params_t *p, orc_la_frame_t **frames, int p0, int p1, int b, int referenced, float fps_factor )
{
    orc_la_frame_t *fb = frames[b];
    uint16_t *ref_costs[2] = { frames[p0]->propagate_cost, frames[p1]->propagate_cost };
    int dist_scale_factor = ( ( ( b - p0 ) << 8 ) + ( ( p1 - p0 ) >> 1 ) ) / ( p1 - p0 );
    int bipred_weight = p->weighted_bipred ? 64 - ( dist_scale_factor >> 2 ) : 32;
    int bipred_weights[2] = { bipred_weight, 64 - bipred_weight };
    int16_t (*mvs[2])[2] = { b != p0 ? fb->lowres_mvs[0][b-p0-1] : NULL, b != p1 ? fb->lowres_mvs[1][p1-b-1] : NULL };
    const uint16_t *lowres_costs = fb->lowres_costs[b-p0][p1-b];
    const unsigned width = p->mb_width, height = p->mb_height;
    if( !referenced )                                   /* slicetype.c:1066-1067: one zeroed row is re-used as the input */
        memset( fb->propagate_cost, 0, width * sizeof(uint16_t) );
    for( unsigned mb_y = 0; mb_y < height; mb_y++ )
    {
        const uint16_t *propagate_in = fb->propagate_cost + ( referenced ? mb_y * width : 0 );
        int16_t amount[width];
        for( unsigned i = 0; i < width; i++ )
        {   /* mbtree_propagate_cost, mc.c:511-527 */
            unsigned mb = mb_y * width + i;
            int intra_cost = (uint16_t)fb->intra_cost[mb];
            int inter_cost = lowres_costs[mb] & 0x3fff;
            if( inter_cost > intra_cost ) inter_cost = intra_cost;
            float propagate_intra = intra_cost * fb->inv_qscale_factor[mb];
            float propagate_amount = propagate_in[i] + propagate_intra * fps_factor;
            float propagate_num = intra_cost - inter_cost;
            float propagate_denom = intra_cost;
            int v = (int)( propagate_amount * propagate_num / propagate_denom + 0.5f );
            amount[i] = v > 32767 ? 32767 : v;
        }
        for( int list = 0; list < ( b != p1 ? 2 : 1 ); list++ )
            for( unsigned i = 0; i < width; i++ )
            {   /* mbtree_propagate_list, mc.c:529-598 */
                unsigned mb = mb_y * width + i;
                int lists_used = lowres_costs[mb] >> 14;
                if( !( lists_used & ( 1 << list ) ) )
                    continue;
                int listamount = amount[i];
                if( lists_used == 3 )
                    listamount = ( listamount * bipred_weights[list] + 32 ) >> 6;
                int x = mvs[list][mb][0], y = mvs[list][mb][1];
                if( !( x | y ) )
                {
                    clip_add( &ref_costs[list][mb], listamount );
                    continue;
                }
                unsigned mbx = (unsigned)( ( x >> 5 ) + (int)i ), mby = (unsigned)( ( y >> 5 ) + (int)mb_y );
                unsigned idx0 = mbx + mby * width, idx2 = idx0 + width;
                x &= 31; y &= 31;
                int w0 = ( ( 32 - y ) * ( 32 - x ) * listamount + 512 ) >> 10;
                int w1 = ( ( 32 - y ) * x * listamount + 512 ) >> 10;
                int w2 = ( y * ( 32 - x ) * listamount + 512 ) >> 10;
                int w3 = ( y * x * listamount + 512 ) >> 10;
                if( mby < height )
                {
                    if( mbx < width ) clip_add( &ref_costs[list][idx0], w0 );
                    if( mbx + 1 < width ) clip_add( &ref_costs[list][idx0 + 1], w1 );
                }
                if( mby + 1 < height )
                {
                    if( mbx < width ) clip_add( &ref_costs[list][idx2], w2 );
                    if( mbx + 1 < width ) clip_add( &ref_costs[list][idx2 + 1], w3 );
                }
            }
    }
}

void orc_la_mbtree_finish( orc_la_frame_t *f, int fps_factor, int ref0_distance, float strength )
{
    float weightdelta = 0.0f;
    if( ref0_distance && f->weighted_cost_delta[ref0_distance-1] > 0 )
        weightdelta = ( 1.0 - f->weighted_cost_delta[ref0_distance-1] );
    for( int mb = 0; mb < f->mb_count; mb++ )
    {
        int intra_cost = ( (uint16_t)f->intra_cost[mb] * f->inv_qscale_factor[mb] + 128 ) >> 8;
        if( intra_cost )
        {
            int propagate_cost = ( f->propagate_cost[mb] * fps_factor + 128 ) >> 8;
            /* x264_log2(a) - x264_log2(b) + weightdelta with a = intra + propagate.  The reference is built with -ffast-math
             * (configure:1413), which lets the compiler re-associate the five float terms; the association below is the one
             * gcc 13 -O3 emits for slicetype.c:1044 (checked in the disassembly of oracle/_ref) and makes f_qp_offset
             * bit-identical to that build; any other association differs by at most a few ulp of 16.0 */
            uint32_t a = intra_cost + propagate_cost, b = intra_cost;
            float log2_ratio = ( ( log2_frac( a ) - log2_int( b ) ) + ( log2_int( a ) + weightdelta ) ) - log2_frac( b );
            f->qp_offset[mb] = f->qp_offset_aq[mb] - strength * log2_ratio;
        }
    }
}

/* slicetype_frame_cost_recalculate, slicetype.c:999-1024: the frame cost with every macroblock's lowres cost scaled by
 * x264_exp2fix8 of its quantiser offset (MB-tree's f_qp_offset; f_qp_offset_aq for B pictures); rewrites
 * row_satds[b-p0][p1-b] and returns the sum over the interior macroblocks (all of them for frames <= 2 macroblocks wide / high). */
static int la_exp2fix8( float x )                                /* x264_exp2fix8, common/base.h:218-224 */
{
    static uint8_t lut[64];
    static int init = 0;
    if( !init )
    {   /* x264_exp2_lut (common/tables.c:58-64): round( 256 * (2^(i/64) - 1) ) */
        for( int i = 0; i < 64; i++ ) lut[i] = (uint8_t)( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) + 0.5 );
        init = 1;
    }
    int i = x * ( -64.f / 6.f ) + 512.5f;
    if( i < 0 ) return 0;
    if( i > 1023 ) return 0xffff;
    return ( lut[i & 63] + 256 ) << ( i >> 6 ) >> 8;
}

int orc_la_frame_cost_recalculate( const orc_la_params_t *p, orc_la_frame_t **frames, int p0, int p1, int b, int b_is_b_type )
{
    orc_la_frame_t *f = frames[b];
    const float *qp_offset = b_is_b_type ? f->qp_offset_aq : f->qp_offset;
    const uint16_t *costs = f->lowres_costs[b-p0][p1-b];
    int *row_satd = f->row_satds[b-p0][p1-b];
    const int w = p->mb_width, h = p->mb_height;
    int score = 0;
    for( int y = h - 1; y >= 0; y-- )
    {
        row_satd[y] = 0;
        for( int x = w - 1; x >= 0; x-- )
        {
            /* LOWRES_COST_MASK, frame.h:110; lowres_costs[0][0] IS i_intra_cost in the reference (frame.c:287) */
            int cost = ( b == p0 && b == p1 ? (uint16_t)f->intra_cost[x + y*w] : costs[x + y*w] ) & 0x3fff;
            cost = ( cost * la_exp2fix8( qp_offset[x + y*w] ) + 128 ) >> 8;
            row_satd[y] += cost;
            if( ( y > 0 && y < h - 1 && x > 0 && x < w - 1 ) || w <= 2 || h <= 2 )
                score += cost;
        }
    }
    return score;
}
