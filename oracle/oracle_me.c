/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Scalar restatement of the reference's block-match
 * search: predictor stage, DIA / HEX / UMH integer search and the sub-pel refinement
 * (encoder/me.c:182-798 x264_me_search_ref, :865-992 refine_subpel) with the chroma branch of the quarter-pel
 * refinement (me.c:826-857, 4:2:0; the lookahead disables it: slicetype.c:60), the exhaustive searches ESA / TESA (me.c:618-771) with their successive-elimination
 * prefilter, and x264_me_refine_bidir_satd (me.c:1027-1183).  fpelcmp is SAD, except under TESA where it is the same
 * metric as mbcmp (encoder.c:1409-1427); mbcmp is SATD iff the encoder's subme > 1.  Written from the algorithm as a small state machine over one `search_t`;
 * tie-breaking follows the reference's packed (cost<<k)+index comparisons exactly (me.c:325-341, :369-418).
 *
 * Parity status: PINNED against the compiled reference's x264_me_search_ref (tests/test_oracle_me.py).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

typedef struct
{
    const orc_me_ctx_t *c;
    orc_me_t *m;
    int bw, bh;
    const uint16_t *cmx, *cmy;      /* cost tables re-centred on the predictor: cmx[mv_x], cmy[mv_y] (qpel units) */
    int x_min, x_max, y_min, y_max; /* full-pel window */
    int bmx, bmy, bcost;
} search_t;

static inline int clip3( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }
static inline int imin( int a, int b ) { return a < b ? a : b; }
static inline int imax( int a, int b ) { return a > b ? a : b; }
static inline uint32_t pack_mv( int x, int y ) { return ( (uint32_t)x & 0xFFFF ) + ( (uint32_t)y << 16 ); }  /* macroblock.h:395 */
#define FPEL(v) (((v)+2)>>2)                                                                                  /* me.c:178 */

static inline int in_range( const search_t *s, int mx, int my )     /* CHECK_MVRANGE, me.c:209 */
{
    return mx >= s->x_min && mx <= s->x_max && my >= s->y_min && my <= s->y_max;
}

/* fpelcmp is SATD only when the encoder runs TESA with subme > 1 (encoder.c:1409-1427) */
static inline int fpel_satd( const search_t *s ) { return s->c->me_method == ORC_ME_TESA && s->c->mbcmp_is_satd; }

/* integer-pel fpelcmp against the (weighted) full-pel plane: COST_MV's value without the mv bits, me.c:63-70 */
static int sad_fpel( const search_t *s, int mx, int my )
{
    const orc_me_t *m = s->m;
    const uint8_t *r = m->p_fref_w + (intptr_t)my*m->stride + mx;
    return fpel_satd( s ) ? orc_satd( m->p_fenc, m->fenc_stride, r, m->stride, s->bw, s->bh )
                          : orc_sad ( m->p_fenc, m->fenc_stride, r, m->stride, s->bw, s->bh );
}
static inline int bits_fpel( const search_t *s, int mx, int my ) { return s->cmx[mx*4] + s->cmy[my*4]; }   /* BITS_MVD */
static int cost_fpel( const search_t *s, int mx, int my ) { return sad_fpel( s, mx, my ) + bits_fpel( s, mx, my ); }

static void try_fpel( search_t *s, int mx, int my )                 /* COST_MV + COPY3_IF_LT */
{
    int cost = cost_fpel( s, mx, my );
    if( cost < s->bcost ) { s->bcost = cost; s->bmx = mx; s->bmy = my; }
}

/* quarter-pel candidate: interpolate (mc.c:198-249) then compare.  satd selects mbcmp vs fpelcmp. */
static int cost_qpel( const search_t *s, int mx, int my, int use_mbcmp )
{
    const orc_me_t *m = s->m;
    uint8_t blk[16*16];
    orc_mc_luma( blk, 16, m->p_fref, m->stride, mx, my, s->bw, s->bh, &m->weight );
    int d = ( use_mbcmp ? s->c->mbcmp_is_satd : fpel_satd( s ) ) ? orc_satd( m->p_fenc, m->fenc_stride, blk, 16, s->bw, s->bh )
                                                                 : orc_sad ( m->p_fenc, m->fenc_stride, blk, 16, s->bw, s->bh );
    d += s->cmx[mx] + s->cmy[my];
    if( use_mbcmp && s->c->chroma_me && m->i_pixel <= ORC_PIXEL_8x8 )
    {   /* COST_MV_SATD's chroma branch, me.c:826-857 (4:2:0): both planes at the luma vector read in eighth-pels, their explicit
         * weights, mbcmp of the half-size block.  The reference stops adding once the sum reaches the best cost, which cannot
         * change what is accepted (partial sums only grow) -- so the whole sum is formed here. */
        const int cw = s->bw >> 1, ch = s->bh >> 1;
        for( int comp = 0; comp < 2; comp++ )
        {
            uint8_t pred[8*8], src[8*8];
            orc_mc_chroma( pred, 8, m->p_fref_uv, m->stride_uv, mx, my, cw, ch, comp );
            if( m->weight_uv[comp].enabled )
                orc_mc_weight( pred, 8, pred, 8, &m->weight_uv[comp], cw, ch );
            for( int y = 0; y < ch; y++ )
                for( int x = 0; x < cw; x++ )
                    src[y*8 + x] = m->p_fenc_uv[y*m->fenc_uv_stride + 2*x + comp];
            d += s->c->mbcmp_is_satd ? orc_satd( src, 8, pred, 8, cw, ch ) : orc_sad( src, 8, pred, 8, cw, ch );
        }
    }
    return d;
}

/* four full-pel candidates around (ox,oy), each a plain strict-less update in order: COST_MV_X4, me.c:100-118 */
static void try4( search_t *s, int ox, int oy, const int8_t d[4][2] )
{
    int costs[4];
    for( int i = 0; i < 4; i++ ) costs[i] = cost_fpel( s, ox + d[i][0], oy + d[i][1] );
    for( int i = 0; i < 4; i++ )
        if( costs[i] < s->bcost ) { s->bcost = costs[i]; s->bmx = ox + d[i][0]; s->bmy = oy + d[i][1]; }
}
static void dia1( search_t *s, int ox, int oy )                     /* DIA1_ITER, me.c:143-150 */
{
    static const int8_t d[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
    try4( s, ox, oy, d );
}

/* CROSS, me.c:152-176: centred on (ox,oy), does not move the centre */
static void cross( search_t *s, int ox, int oy, int start, int x_max, int y_max )
{
    int i = start;
    if( x_max <= imin( s->x_max - ox, ox - s->x_min ) )
        for( ; i < x_max-2; i += 4 )
        {
            const int8_t d[4][2] = { {i,0}, {-i,0}, {i+2,0}, {-i-2,0} };
            try4( s, ox, oy, d );
        }
    for( ; i < x_max; i += 2 )
    {
        if( ox+i <= s->x_max ) try_fpel( s, ox+i, oy );
        if( ox-i >= s->x_min ) try_fpel( s, ox-i, oy );
    }
    i = start;
    if( y_max <= imin( s->y_max - oy, oy - s->y_min ) )
        for( ; i < y_max-2; i += 4 )
        {
            const int8_t d[4][2] = { {0,i}, {0,-i}, {0,i+2}, {0,-i-2} };
            try4( s, ox, oy, d );
        }
    for( ; i < y_max; i += 2 )
    {
        if( oy+i <= s->y_max ) try_fpel( s, ox, oy+i );
        if( oy-i >= s->y_min ) try_fpel( s, ox, oy-i );
    }
}

/* hexagon + square refinement, me.c:344-420 (the de-duplicated form) */
static void hex_search( search_t *s, int me_range )
{
    static const int8_t hex2[8][2] = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
    static const uint8_t mod6m1[8] = { 5,0,1,2,3,4,5,0 };
    static const int8_t square1[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
    static const int8_t first[6][2] = { {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2} };   /* indices 2..7 */
    int packed = s->bcost << 3;
    for( int k = 0; k < 6; k++ )
    {
        int cst = ( cost_fpel( s, s->bmx + first[k][0], s->bmy + first[k][1] ) << 3 ) + k + 2;
        if( cst < packed ) packed = cst;
    }
    if( packed & 7 )
    {
        int dir = ( packed & 7 ) - 2;
        s->bmx += hex2[dir+1][0];
        s->bmy += hex2[dir+1][1];
        for( int i = ( me_range >> 1 ) - 1; i > 0 && in_range( s, s->bmx, s->bmy ); i-- )
        {
            packed &= ~7;
            for( int k = 0; k < 3; k++ )
            {
                int cst = ( cost_fpel( s, s->bmx + hex2[dir+k][0], s->bmy + hex2[dir+k][1] ) << 3 ) + k + 1;
                if( cst < packed ) packed = cst;
            }
            if( !( packed & 7 ) )
                break;
            dir += ( packed & 7 ) - 2;
            dir = mod6m1[dir+1];
            s->bmx += hex2[dir+1][0];
            s->bmy += hex2[dir+1][1];
        }
    }
    packed = ( packed >> 3 ) << 4;
    for( int k = 1; k < 9; k++ )
    {
        int cst = ( cost_fpel( s, s->bmx + square1[k][0], s->bmy + square1[k][1] ) << 4 ) + k;
        if( cst < packed ) packed = cst;
    }
    s->bmx += square1[packed & 15][0];
    s->bmy += square1[packed & 15][1];
    s->bcost = packed >> 4;
}

/* ---- exhaustive search with successive elimination, me.c:618-771 -------------------------------------------------------
 * The reference keeps, per reference frame, an "integral" plane holding the pixel sum of the 8x8 (and 4x4) block at every
 * position of the UNWEIGHTED padded plane (mc.c:424-456, :748-783; uint16 arithmetic, exact because 64*255 < 65536) and
 * drops a position before measuring it when ads = sum over the block's 8x8 / 4x4 sub-blocks |dc(fenc) - dc(ref)| + the x mv
 * cost reaches a threshold (pixel.c:759-803).  The sub-block layout per partition (me.c:637-654 with pixf.ads, pixel.c:860,
 * :1660): 16x16 four 8x8s, 16x8 / 8x16 two 8x8s, 8x8 one, 8x4 / 4x8 two 4x4s, 4x4 one.  Here the sums are taken from the
 * pixels directly.  ESA: threshold = the best cost, survivors go through COST_MV in raster order.  TESA: threshold
 * 17/16 of the best SAD, survivors within sad_thresh/8 of the running best SAD are listed, the list is thinned to
 * me_range/2 entries and those are measured with fpelcmp (SATD). */
static int block_sum( const uint8_t *p, intptr_t stride, int w, int h )
{
    int v = 0;
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
            v += p[y*stride + x];
    return v;
}

typedef struct { int sb, nx, ny; int enc_dc[4]; } ads_t;

static void ads_init( const search_t *s, ads_t *a )
{
    a->sb = ( s->bw >= 8 && s->bh >= 8 ) ? 8 : 4;
    a->nx = s->bw / a->sb; a->ny = s->bh / a->sb;
    for( int j = 0; j < a->ny; j++ )
        for( int i = 0; i < a->nx; i++ )
            a->enc_dc[j*a->nx + i] = block_sum( s->m->p_fenc + j*a->sb*s->m->fenc_stride + i*a->sb, s->m->fenc_stride, a->sb, a->sb );
}

static int ads_at( const search_t *s, const ads_t *a, int mx, int my )   /* pixf.ads' value without the mv cost */
{
    const orc_me_t *m = s->m;
    const uint8_t *r = m->p_fref[0] + (intptr_t)my*m->stride + mx;      /* the integral image is built from the unweighted plane */
    int v = 0;
    for( int j = 0; j < a->ny; j++ )
        for( int i = 0; i < a->nx; i++ )
            v += abs( a->enc_dc[j*a->nx + i] - block_sum( r + j*a->sb*m->stride + i*a->sb, m->stride, a->sb, a->sb ) );
    return v;
}

typedef struct { int sad; int16_t mv[2]; } mvsad_t;

static void exhaustive_search( search_t *s, int me_range )
{
    const orc_me_t *m = s->m;
    const int min_x = imax( s->bmx - me_range, s->x_min ), min_y = imax( s->bmy - me_range, s->y_min );
    const int max_x = imin( s->bmx + me_range, s->x_max ), max_y = imin( s->bmy + me_range, s->y_max );
    const int width = ( max_x - min_x + 3 ) & ~3;       /* rounded up: reaches up to 3 positions past mv_x_max, as the reference */
    ads_t a;
    ads_init( s, &a );
    if( s->c->me_method == ORC_ME_ESA )
    {   /* me.c:751-768: "just ADS and SAD" */
        for( int my = min_y; my <= max_y; my++ )
        {
            int ycost = s->cmy[my*4];
            if( s->bcost <= ycost )
                continue;
            const int thresh = s->bcost - ycost;        /* fixed for the whole row (me.c:760) */
            for( int mx = min_x; mx < min_x + width; mx++ )
                if( ads_at( s, &a, mx, my ) + s->cmx[mx*4] < thresh )
                    try_fpel( s, mx, my );
        }
        return;
    }
    /* TESA, me.c:656-747 */
    int rows = max_y - min_y + 1;
    mvsad_t *list = malloc( sizeof(mvsad_t) * ( rows > 0 && width > 0 ? (size_t)rows * width : 1 ) );
    int n = 0;
    int sad_thresh = me_range <= 16 ? 10 : me_range <= 24 ? 11 : 12;
    int bsad = orc_sad( m->p_fenc, m->fenc_stride, m->p_fref_w + (intptr_t)s->bmy*m->stride + s->bmx, m->stride, s->bw, s->bh )
             + bits_fpel( s, s->bmx, s->bmy );
    for( int my = min_y; my <= max_y; my++ )
    {
        int ycost = s->cmy[my*4];
        if( bsad <= ycost )
            continue;
        bsad -= ycost;
        const int athresh = bsad * 17 >> 4;
        for( int mx = min_x; mx < min_x + width; mx++ )
        {
            if( ads_at( s, &a, mx, my ) + s->cmx[mx*4] >= athresh )
                continue;
            /* the x mv cost of the LISTING step is read at the position's index within the row, cost_fpel_mvx[xs[i]] (me.c:683,
             * :699), not at min_x + xs[i] as the ADS step does (me.c:675): reproduced as is */
            int sad = orc_sad( m->p_fenc, m->fenc_stride, m->p_fref_w + (intptr_t)my*m->stride + mx, m->stride, s->bw, s->bh )
                    + s->cmx[( mx - min_x )*4];
            if( sad < bsad * sad_thresh >> 3 )
            {
                if( sad < bsad ) bsad = sad;
                list[n].sad = sad + ycost; list[n].mv[0] = mx; list[n].mv[1] = my;
                n++;
            }
        }
        bsad += ycost;
    }
    int limit = me_range >> 1;
    sad_thresh = bsad * sad_thresh >> 3;
    while( n > limit*2 && sad_thresh > bsad )
    {   /* halve the admitted range and keep, in order, what is still inside it */
        sad_thresh = ( sad_thresh + bsad ) >> 1;
        int k = 0;
        for( int j = 0; j < n; j++ )
            if( list[j].sad <= sad_thresh )
                list[k++] = list[j];
        n = k;
    }
    while( n > limit )
    {   /* drop the (first) worst, the last entry takes its place */
        int bi = 0;
        for( int i = 1; i < n; i++ )
            if( list[i].sad > list[bi].sad )
                bi = i;
        n--;
        list[bi] = list[n];
    }
    for( int i = 0; i < n; i++ )
        try_fpel( s, list[i].mv[0], list[i].mv[1] );
    free( list );
}

static void refine_subpel( search_t *s, int hpel_iters, int qpel_iters, int *p_halfpel_thresh, int b_refine_qpel );

void orc_me_search_ref( const orc_me_ctx_t *c, orc_me_t *m, const int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh )
{
    static const uint8_t subpel_iterations[12][4] =                  /* me.c:38-50 */
        { {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2}, {0,0,2,2}, {0,0,2,2},
          {0,0,4,10}, {0,0,4,10}, {0,0,4,10}, {0,0,4,10} };
    search_t S, *s = &S;
    s->c = c; s->m = m;
    s->bw = orc_pixel_w[m->i_pixel]; s->bh = orc_pixel_h[m->i_pixel];
    s->cmx = m->p_cost_mv - m->mvp[0];
    s->cmy = m->p_cost_mv - m->mvp[1];
    s->x_min = c->mv_limit_fpel[0][0]; s->y_min = c->mv_limit_fpel[0][1];
    s->x_max = c->mv_limit_fpel[1][0]; s->y_max = c->mv_limit_fpel[1][1];
    s->bcost = ORC_COST_MAX;
    int me_range = c->me_range;
    int bpred_cost = ORC_COST_MAX;
    int pmx, pmy;
    uint32_t pmv, bpred_mv = 0;
    int16_t cand[16][2];

    /* ---- predictor stage, me.c:216-318 ---- */
    if( c->subpel_refine >= 3 )
    {
        int bpx = clip3( m->mvp[0], s->x_min*4, s->x_max*4 );
        int bpy = clip3( m->mvp[1], s->y_min*4, s->y_max*4 );
        pmv = pack_mv( bpx, bpy );
        pmx = FPEL( bpx ); pmy = FPEL( bpy );
        bpred_cost = cost_qpel( s, bpx, bpy, 0 );
        int pmv_cost = bpred_cost;
        if( i_mvc > 0 )
        {
            /* x264_predictor_clip, common.h:791-806: drop zero and == pmv (compared before clipping) */
            int n = 0;
            for( int i = 0; i < i_mvc; i++ )
            {
                uint32_t mv = pack_mv( mvc[i][0], mvc[i][1] );
                if( !mv || mv == pmv ) continue;
                cand[n][0] = clip3( mvc[i][0], s->x_min*4, s->x_max*4 );
                cand[n][1] = clip3( mvc[i][1], s->y_min*4, s->y_max*4 );
                n++;
            }
            if( n > 0 )
            {
                int packed = bpred_cost << 4;          /* index 0 = the predictor itself */
                for( int i = 1; i <= n; i++ )
                {
                    int cst = ( cost_qpel( s, cand[i-1][0], cand[i-1][1], 0 ) << 4 ) + i;
                    if( cst < packed ) packed = cst;
                }
                if( packed & 15 ) { bpx = cand[(packed&15)-1][0]; bpy = cand[(packed&15)-1][1]; }
                bpred_cost = packed >> 4;
            }
        }
        s->bmx = FPEL( bpx ); s->bmy = FPEL( bpy );
        bpred_mv = pack_mv( bpx, bpy );
        if( bpred_mv & 0x00030003 )
            try_fpel( s, s->bmx, s->bmy );
        else
            s->bcost = bpred_cost;
        if( pmv )
        {
            if( s->bmx | s->bmy ) try_fpel( s, 0, 0 );
        }
        else if( pmv_cost < s->bcost ) { s->bcost = pmv_cost; s->bmx = 0; s->bmy = 0; }
    }
    else
    {
        s->bmx = pmx = clip3( FPEL( m->mvp[0] ), s->x_min, s->x_max );
        s->bmy = pmy = clip3( FPEL( m->mvp[1] ), s->y_min, s->y_max );
        pmv = pack_mv( s->bmx, s->bmy );
        /* the rounded predictor is measured WITHOUT its mv cost (me.c:283-291) */
        s->bcost = sad_fpel( s, s->bmx, s->bmy );
        if( i_mvc > 0 )
        {
            /* x264_predictor_roundclip, common.h:774-789 */
            int n = 0;
            for( int i = 0; i < i_mvc; i++ )
            {
                int mx = ( mvc[i][0] + 2 ) >> 2, my = ( mvc[i][1] + 2 ) >> 2;
                uint32_t mv = pack_mv( mx, my );
                if( !mv || mv == pmv ) continue;
                cand[n][0] = clip3( mx, s->x_min, s->x_max );
                cand[n][1] = clip3( my, s->y_min, s->y_max );
                n++;
            }
            if( n > 0 )
            {
                int packed = s->bcost << 4;
                for( int i = 1; i <= n; i++ )
                {
                    int cst = ( cost_fpel( s, cand[i-1][0], cand[i-1][1] ) << 4 ) + i;
                    if( cst < packed ) packed = cst;
                }
                if( packed & 15 ) { s->bmx = cand[(packed&15)-1][0]; s->bmy = cand[(packed&15)-1][1]; }
                s->bcost = packed >> 4;
            }
        }
        if( pmv )
            try_fpel( s, 0, 0 );
    }

    /* ---- integer search ---- */
    switch( c->me_method )
    {
    case ORC_ME_DIA:                                                    /* me.c:322-342 */
    {
        static const int8_t d[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
        int i = me_range;
        do
        {
            /* the reference tags the four directions 1,3,4,12 inside (cost<<4)+tag; only strict cost order
             * and, at equal cost, the smaller tag matter -- the tags grow in candidate order and the centre
             * has tag 0, so "first strictly smaller cost wins, ties keep the earlier" reproduces it */
            int packed = s->bcost, best = -1;
            for( int k = 0; k < 4; k++ )
            {
                int cst = cost_fpel( s, s->bmx + d[k][0], s->bmy + d[k][1] );
                if( cst < packed ) { packed = cst; best = k; }
            }
            s->bcost = packed;
            if( best < 0 )
                break;
            s->bmx += d[best][0];
            s->bmy += d[best][1];
        } while( --i && in_range( s, s->bmx, s->bmy ) );
        break;
    }
    case ORC_ME_HEX:
        hex_search( s, me_range );
        break;
    case ORC_ME_ESA:                                                    /* me.c:618-771 */
    case ORC_ME_TESA:
        exhaustive_search( s, me_range );
        break;
    case ORC_ME_UMH:                                                    /* me.c:422-616 */
    {
        static const uint8_t pixel_size_shift[7] = { 0, 1, 1, 2, 3, 3, 4 };
        int ucost1, ucost2, cross_start = 1, done = 0;
        ucost1 = s->bcost;
        dia1( s, pmx, pmy );
        if( pmx | pmy )
            dia1( s, 0, 0 );
        if( m->i_pixel == ORC_PIXEL_4x4 )
        {
            hex_search( s, me_range );
            break;
        }
        ucost2 = s->bcost;
        if( ( s->bmx | s->bmy ) && ( ( s->bmx - pmx ) | ( s->bmy - pmy ) ) )
            dia1( s, s->bmx, s->bmy );
        if( s->bcost == ucost2 )
            cross_start = 3;
        int omx = s->bmx, omy = s->bmy;
#define SAD_THRESH(v) ( s->bcost < ( (v) >> pixel_size_shift[m->i_pixel] ) )
        if( s->bcost == ucost2 && SAD_THRESH(2000) )
        {
            static const int8_t o1[4][2] = { {0,-2}, {-1,-1}, {1,-1}, {-2,0} };
            static const int8_t o2[4][2] = { {2,0}, {-1,1}, {1,1}, {0,2} };
            try4( s, omx, omy, o1 );
            try4( s, omx, omy, o2 );
            if( s->bcost == ucost1 && SAD_THRESH(500) )
                done = 1;
            else if( s->bcost == ucost2 )
            {
                static const int8_t o3[4][2] = { {-1,-2}, {1,-2}, {-2,-1}, {2,-1} };
                static const int8_t o4[4][2] = { {-2,1}, {2,1}, {-1,2}, {1,2} };
                int range = ( me_range >> 1 ) | 1;
                cross( s, omx, omy, 3, range, range );
                try4( s, omx, omy, o3 );
                try4( s, omx, omy, o4 );
                if( s->bcost == ucost2 )
                    done = 1;
                cross_start = range + 2;
            }
        }
        if( done )
            break;
        /* adaptive search range, me.c:469-519 */
        if( i_mvc )
        {
            static const uint8_t range_mul[4][4] = { {3,3,4,4}, {3,4,4,4}, {4,4,4,5}, {4,4,5,6} };
            int mvd, denom = 1;
            if( i_mvc == 1 )
            {
                if( m->i_pixel == ORC_PIXEL_16x16 )
                    mvd = 25;
                else
                    mvd = abs( m->mvp[0] - mvc[0][0] ) + abs( m->mvp[1] - mvc[0][1] );
            }
            else
            {
                denom = i_mvc - 1;
                mvd = 0;
                if( m->i_pixel != ORC_PIXEL_16x16 )
                {
                    mvd = abs( m->mvp[0] - mvc[0][0] ) + abs( m->mvp[1] - mvc[0][1] );
                    denom++;
                }
                for( int i = 0; i < i_mvc-1; i++ )                   /* x264_predictor_difference, base.h:248 */
                    mvd += abs( mvc[i][0] - mvc[i+1][0] ) + abs( mvc[i][1] - mvc[i+1][1] );
            }
            int sad_ctx = SAD_THRESH(1000) ? 0 : SAD_THRESH(2000) ? 1 : SAD_THRESH(4000) ? 2 : 3;
            int mvd_ctx = mvd < 10*denom ? 0 : mvd < 20*denom ? 1 : mvd < 40*denom ? 2 : 3;
            me_range = me_range * range_mul[mvd_ctx][sad_ctx] >> 2;
        }
#undef SAD_THRESH
        /* uneven cross and 5x5 corners stay centred on omx/omy even if the best moved (me.c:521-525) */
        cross( s, omx, omy, cross_start, me_range, me_range >> 1 );
        {
            static const int8_t o5[4][2] = { {-2,-2}, {-2,2}, {2,-2}, {2,2} };
            try4( s, omx, omy, o5 );
        }
        /* hexagon grid, me.c:527-612: 16 points scaled by i = 1 .. range/4, centre fixed at the best so far */
        omx = s->bmx; omy = s->bmy;
        {
            static const int8_t hex4[16][2] = {
                { 0,-4}, { 0, 4}, {-2,-3}, { 2,-3}, {-4,-2}, { 4,-2}, {-4,-1}, { 4,-1},
                {-4, 0}, { 4, 0}, {-4, 1}, { 4, 1}, {-4, 2}, { 4, 2}, {-2, 3}, { 2, 3} };
            int i = 1;
            do
            {
                int room = imin( imin( s->x_max - omx, omx - s->x_min ), imin( s->y_max - omy, omy - s->y_min ) );
                if( 4*i > room )
                {
                    for( int j = 0; j < 16; j++ )
                    {
                        int mx = omx + hex4[j][0]*i, my = omy + hex4[j][1]*i;
                        if( in_range( s, mx, my ) )
                            try_fpel( s, mx, my );
                    }
                }
                else
                {
                    /* all 16 costs first, then one pass of strict-less minimum in table order */
                    int costs[16], dir = -1;
                    for( int j = 0; j < 16; j++ )
                        costs[j] = cost_fpel( s, omx + hex4[j][0]*i, omy + hex4[j][1]*i );
                    for( int j = 0; j < 16; j++ )
                        if( costs[j] < s->bcost ) { s->bcost = costs[j]; dir = j; }
                    if( dir >= 0 )
                    {
                        s->bmx = omx + i*hex4[dir][0];
                        s->bmy = omy + i*hex4[dir][1];
                    }
                }
            } while( ++i <= me_range >> 2 );
        }
        if( s->bmy <= s->y_max && s->bmy >= s->y_min && s->bmx <= s->x_max && s->bmx >= s->x_min )
            hex_search( s, me_range );
        break;
    }
    }

    /* ---- back to quarter-pel units, me.c:774-789 ---- */
    uint32_t bmv = pack_mv( s->bmx, s->bmy );
    if( c->subpel_refine < 3 )
    {
        m->cost_mv = s->cmx[s->bmx*4] + s->cmy[s->bmy*4];
        m->cost = s->bcost;
        if( bmv == pmv ) m->cost += m->cost_mv;                       /* re-add the omitted cost if the MVP won */
        m->mv[0] = s->bmx*4; m->mv[1] = s->bmy*4;
    }
    else
    {
        if( bpred_cost < s->bcost )
        {
            m->mv[0] = (int16_t)( bpred_mv & 0xFFFF ); m->mv[1] = (int16_t)( bpred_mv >> 16 );
            m->cost = bpred_cost;
        }
        else
        {
            m->mv[0] = s->bmx*4; m->mv[1] = s->bmy*4;
            m->cost = s->bcost;
        }
    }
    if( c->subpel_refine >= 2 )
        refine_subpel( s, subpel_iterations[c->subpel_refine][2], subpel_iterations[c->subpel_refine][3], p_halfpel_thresh, 0 );
}

/* encoder/me.c:865-992 */
static void refine_subpel( search_t *s, int hpel_iters, int qpel_iters, int *p_halfpel_thresh, int b_refine_qpel )
{
    const orc_me_ctx_t *c = s->c;
    orc_me_t *m = s->m;
    int bmx = m->mv[0], bmy = m->mv[1], bcost = m->cost;
    int odir = -1, bdir = -1;

    if( hpel_iters )
    {
        if( c->subpel_refine < 3 )
        {   /* the sub-pel part of the predictor, me.c:889-895 */
            int mx = clip3( m->mvp[0], c->mv_min_spel[0]+2, c->mv_max_spel[0]-2 );
            int my = clip3( m->mvp[1], c->mv_min_spel[1]+2, c->mv_max_spel[1]-2 );
            if( ( mx - bmx ) | ( my - bmy ) )
            {
                int cost = cost_qpel( s, mx, my, 0 );
                if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; }
            }
        }
        for( int i = hpel_iters; i > 0; i-- )
        {   /* half-pel diamond, candidates in the order (0,-2) (0,+2) (-2,0) (+2,0); ties keep the earlier */
            static const int8_t d[4][2] = { {0,-2}, {0,2}, {-2,0}, {2,0} };
            int best = -1, bc = bcost;
            for( int k = 0; k < 4; k++ )
            {
                int cost = cost_qpel( s, bmx + d[k][0], bmy + d[k][1], 0 );
                if( cost < bc ) { bc = cost; best = k; }
            }
            bcost = bc;
            if( best < 0 )
                break;
            bmx += d[best][0];
            bmy += d[best][1];
        }
    }

    if( !b_refine_qpel && ( ( c->mbcmp_is_satd && !fpel_satd( s ) ) || ( c->chroma_me && m->i_pixel <= ORC_PIXEL_8x8 ) ) )
    {   /* re-measure the winner with mbcmp (SATD) when the half-pel steps used another metric or left the chroma out, me.c:925-929 */
        bcost = cost_qpel( s, bmx, bmy, 1 );
        bdir = -1;
    }

    if( p_halfpel_thresh )
    {
        if( ( bcost*7 ) >> 3 > *p_halfpel_thresh )
        {
            m->cost = bcost; m->mv[0] = bmx; m->mv[1] = bmy;
            return;
        }
        else if( bcost < *p_halfpel_thresh )
            *p_halfpel_thresh = bcost;
    }

    if( c->subpel_refine != 1 )
    {
        static const int8_t d[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
        bdir = -1;
        for( int i = qpel_iters; i > 0; i-- )
        {
            if( bmy <= c->mv_min_spel[1] || bmy >= c->mv_max_spel[1] || bmx <= c->mv_min_spel[0] || bmx >= c->mv_max_spel[0] )
                break;
            odir = bdir;
            int omx = bmx, omy = bmy;
            for( int dir = 0; dir < 4; dir++ )
            {
                if( !b_refine_qpel && ( dir ^ 1 ) == odir )      /* never step straight back, me.c:828 */
                    continue;
                int cost = cost_qpel( s, omx + d[dir][0], omy + d[dir][1], 1 );
                if( cost < bcost ) { bcost = cost; bmx = omx + d[dir][0]; bmy = omy + d[dir][1]; bdir = dir; }
            }
            if( bmx == omx && bmy == omy )
                break;
        }
    }
    else if( bmy > c->mv_min_spel[1] && bmy < c->mv_max_spel[1] && bmx > c->mv_min_spel[0] && bmx < c->mv_max_spel[0] )
    {   /* subme 1: one SAD quarter-pel diamond, me.c:965-985; candidate order (0,-1) (0,+1) (-1,0) (+1,0) */
        static const int8_t d[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
        int best = -1, bc = bcost;
        for( int k = 0; k < 4; k++ )
        {
            int cost = cost_qpel( s, bmx + d[k][0], bmy + d[k][1], 0 );
            if( cost < bc ) { bc = cost; best = k; }
        }
        bcost = bc;
        if( best >= 0 ) { bmx += d[best][0]; bmy += d[best][1]; }
    }

    m->cost = bcost;
    m->mv[0] = bmx;
    m->mv[1] = bmy;
    m->cost_mv = s->cmx[bmx] + s->cmy[bmy];
}

/* ---- x264_me_refine_qpel / x264_me_refine_qpel_refdupe, encoder/me.c:800-814 ---------------------------------------------
 * Both continue from m->mv / m->cost.  mode 0 (refine_qpel): the iteration counts of the FINAL refinement
 * (subpel_iterations[subme][0..1]), every direction tried in each quarter-pel round and no re-measurement of the start;
 * partitions of 8x8 and larger first give back i_ref_cost.  mode 1 (refdupe): quarter-pel rounds only, at most two. */
void orc_me_refine_qpel( const orc_me_ctx_t *c, orc_me_t *m, int mode, int i_ref_cost, int *p_halfpel_thresh )
{
    static const uint8_t subpel_iterations[12][4] =                  /* me.c:38-50 */
        { {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2}, {0,0,2,2}, {0,0,2,2},
          {0,0,4,10}, {0,0,4,10}, {0,0,4,10}, {0,0,4,10} };
    search_t S, *s = &S;
    s->c = c; s->m = m;
    s->bw = orc_pixel_w[m->i_pixel]; s->bh = orc_pixel_h[m->i_pixel];
    s->cmx = m->p_cost_mv - m->mvp[0];
    s->cmy = m->p_cost_mv - m->mvp[1];
    if( mode == 0 )
    {
        if( m->i_pixel <= ORC_PIXEL_8x8 )
            m->cost -= i_ref_cost;
        refine_subpel( s, subpel_iterations[c->subpel_refine][0], subpel_iterations[c->subpel_refine][1], NULL, 1 );
    }
    else
        refine_subpel( s, 0, imin( 2, subpel_iterations[c->subpel_refine][3] ), p_halfpel_thresh, 0 );
}

/* ---- x264_me_refine_bidir_satd, encoder/me.c:1027-1183 with rd = 0 --------------------------------------------------------
 * Joint refinement of the two vectors of a bi-predicted block: up to 8 passes; each pass measures the (list 0, list 1)
 * vector pairs that differ from the current pair by +-1 quarter-pel in at most two of the four components (33 pairs, the
 * current one only in the first pass), skipping pairs already measured -- remembered in a 4096-bit map indexed by the low 3
 * bits of every component (so it aliases after a drift of 8, as the reference's) --, cost = mbcmp( fenc, avg( ref0, ref1,
 * weight ) ) + the four mv costs; strictly smaller wins, first in table order on ties; stops when the centre stays. */
static const uint8_t bidir_pairs[33] =        /* offset = base-3 digits - 1, in the order (m0x, m0y, m1x, m1y); me.c:1063-1074 */
    { 40, 67, 13, 49, 31, 43, 37, 41, 39, 76, 4, 52, 28, 44, 36, 68, 12, 70, 10, 50, 30, 58, 22, 46, 34, 42, 38, 14, 66, 64, 16, 48, 32 };

void orc_me_refine_bidir_satd( const orc_me_ctx_t *c, orc_me_t *m0, orc_me_t *m1, int i_weight )
{
    const int bw = orc_pixel_w[m0->i_pixel], bh = orc_pixel_h[m0->i_pixel];
    int bm[4] = { m0->mv[0], m0->mv[1], m1->mv[0], m1->mv[1] };
    m0->cost = m1->cost = ORC_COST_MAX;
    for( int k = 0; k < 4; k++ )                                        /* me.c:1076-1080: too close to the window edge */
        if( bm[k] < c->mv_min_spel[k&1] + 8 || bm[k] > c->mv_max_spel[k&1] - 8 )
            return;
    const uint16_t *tab[4] = { m0->p_cost_mv - m0->mvp[0], m0->p_cost_mv - m0->mvp[1], m1->p_cost_mv - m1->mvp[0], m1->p_cost_mv - m1->mvp[1] };
    static const orc_weight_t none;
    uint8_t visited[8][8][8];
    memset( visited, 0, sizeof(visited) );
    int bcost = ORC_COST_MAX;
    for( int pass = 0; pass < 8; pass++ )
    {
        int bestj = 0;
        for( int j = !!pass; j < 33; j++ )
        {
            int v[4], code = bidir_pairs[j];
            for( int k = 0; k < 4; k++, code /= 3 )
                v[k] = bm[k] + code % 3 - 1;
            uint8_t *seen = &visited[v[0]&7][v[1]&7][v[2]&7];
            if( pass && ( *seen & ( 1 << ( v[3]&7 ) ) ) )
                continue;
            *seen |= 1 << ( v[3]&7 );
            uint8_t p0[16*16], p1[16*16], avg[16*16];
            orc_mc_luma( p0, 16, m0->p_fref, m0->stride, v[0], v[1], bw, bh, &none );
            orc_mc_luma( p1, 16, m1->p_fref, m1->stride, v[2], v[3], bw, bh, &none );
            orc_pixel_avg( avg, 16, p0, 16, p1, 16, bw, bh, i_weight );
            int cost = ( c->mbcmp_is_satd ? orc_satd( m0->p_fenc, m0->fenc_stride, avg, 16, bw, bh )
                                          : orc_sad ( m0->p_fenc, m0->fenc_stride, avg, 16, bw, bh ) )
                     + tab[0][v[0]] + tab[1][v[1]] + tab[2][v[2]] + tab[3][v[3]];
            if( cost < bcost ) { bcost = cost; bestj = j; }
        }
        if( !bestj )
            break;
        int code = bidir_pairs[bestj];
        for( int k = 0; k < 4; k++, code /= 3 )
            bm[k] += code % 3 - 1;
    }
    m0->mv[0] = bm[0]; m0->mv[1] = bm[1];
    m1->mv[0] = bm[2]; m1->mv[1] = bm[3];
    m0->cost = m1->cost = bcost;                                        /* not an output of the reference: kept for the tests */
}
