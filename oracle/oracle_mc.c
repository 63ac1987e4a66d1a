/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain-C restatement of the reference's motion
 * compensation helpers and frame preparation (common/mc.c, common/frame.c, encoder/analyse.c cost table).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int clip_u8( int v ) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int clampi( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }

/* common/frame.c:512-560 plane_expand_border (luma form): replicate the edge pixels padh to the left/right
 * and the (already widened) edge rows padv up/down. */
void orc_plane_expand_border( uint8_t *pix, intptr_t stride, int width, int height, int padh, int padv )
{
    for( int y = 0; y < height; y++ )
    {
        uint8_t *row = pix + y*stride;
        memset( row - padh, row[0], padh );
        memset( row + width, row[width-1], padh );
    }
    for( int y = 1; y <= padv; y++ )
    {
        memcpy( pix - padh - y*stride, pix - padh, width + 2*padh );
        memcpy( pix - padh + (height-1+y)*stride, pix - padh + (height-1)*stride, width + 2*padh );
    }
}

/* common/mc.c:458-507 x264_frame_init_lowres + frame_init_lowres_core, common/frame.c:627-631.
 * `src` is the luma plane of size width x height where width/height are already the mod-16 padded
 * size (the reference pads with x264_frame_expand_border_mod16 before this is called).  The reference
 * duplicates the last column and row of the source in place so that 2x+2 / 2y+2 never leave the plane;
 * here the same thing is done by clamping the read coordinates. */
void orc_frame_init_lowres( const uint8_t *src, intptr_t src_stride, int width, int height,
                            uint8_t *lowres[4], intptr_t dst_stride, int width_lowres, int lines_lowres )
{
#define S(xx,yy) ((int)src[ (intptr_t)clampi( yy, 0, height-1 )*src_stride + clampi( xx, 0, width-1 ) ])
#define FILT(a,b,c,d) ((((a+b+1)>>1)+((c+d+1)>>1)+1)>>1)
    for( int y = 0; y < lines_lowres; y++ )
        for( int x = 0; x < width_lowres; x++ )
        {
            int X = 2*x, Y = 2*y;
            lowres[0][y*dst_stride+x] = FILT( S(X,Y),     S(X,Y+1),   S(X+1,Y),   S(X+1,Y+1) );
            lowres[1][y*dst_stride+x] = FILT( S(X+1,Y),   S(X+1,Y+1), S(X+2,Y),   S(X+2,Y+1) );
            lowres[2][y*dst_stride+x] = FILT( S(X,Y+1),   S(X,Y+2),   S(X+1,Y+1), S(X+1,Y+2) );
            lowres[3][y*dst_stride+x] = FILT( S(X+1,Y+1), S(X+1,Y+2), S(X+2,Y+1), S(X+2,Y+2) );
        }
#undef FILT
#undef S
    for( int i = 0; i < 4; i++ )
        orc_plane_expand_border( lowres[i], dst_stride, width_lowres, lines_lowres, ORC_PAD, ORC_PAD );
}

/* common/mc.c:172-196 hpel_filter as driven by x264_frame_filter (mc.c:704-746) and followed by
 * x264_frame_expand_border_filtered (frame.c:596-625).  The net effect over a whole frame is: each of the
 * H, V, C planes equals the 6-tap (1,-5,20,20,-5,1) filter evaluated on the edge-replicated source at
 * every position of the padded domain [-pad,width+pad) x [-pad,height+pad) (a constant run filters to
 * itself, so the reference's "filter 8 extra pixels, then replicate" produces exactly this).
 * C is filtered horizontally from the UNROUNDED vertical sums, (x+512)>>10.
 * dsth/dstv/dstc point at the plane origin (pixel 0,0) and must have `pad` pixels of room around. */
void orc_hpel_filter_plane( const uint8_t *src, intptr_t stride, int width, int height,
                            uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, intptr_t dst_stride, int pad )
{
    static const int tap[6] = { 1, -5, 20, 20, -5, 1 };
#define S(xx,yy) ((int)src[ (intptr_t)clampi( yy, 0, height-1 )*stride + clampi( xx, 0, width-1 ) ])
    for( int y = -pad; y < height+pad; y++ )
        for( int x = -pad; x < width+pad; x++ )
        {
            int hsum = 0, vsum = 0, csum = 0;
            for( int k = 0; k < 6; k++ )
            {
                hsum += tap[k] * S( x+k-2, y );
                vsum += tap[k] * S( x, y+k-2 );
                int vk = 0;
                for( int j = 0; j < 6; j++ )
                    vk += tap[j] * S( x+k-2, y+j-2 );
                csum += tap[k] * vk;
            }
            dsth[y*dst_stride+x] = clip_u8( (hsum + 16) >> 5 );
            dstv[y*dst_stride+x] = clip_u8( (vsum + 16) >> 5 );
            dstc[y*dst_stride+x] = clip_u8( (csum + 512) >> 10 );
        }
#undef S
}

/* common/mc.c:117-137 mc_weight */
void orc_mc_weight( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, const orc_weight_t *wt, int w, int h )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            int p = src[y*ss+x];
            int v = wt->denom >= 1 ? ((p * wt->scale + (1 << (wt->denom-1))) >> wt->denom) + wt->offset
                                   : p * wt->scale + wt->offset;
            dst[y*sd+x] = clip_u8( v );
        }
}

/* common/frame.c:825-841: whole-plane weighting is just mc_weight over the plane */
void orc_weight_scale_plane( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, int w, int h, const orc_weight_t *wt )
{
    orc_mc_weight( dst, sd, src, ss, wt, w, h );
}

/* common/mc.c:198-249 mc_luma / get_ref (value semantics: always materialises the block) with the
 * quarter-pel plane-pair tables of common/tables.c:183-184 */
void orc_mc_luma( uint8_t *dst, intptr_t dst_stride, const uint8_t *const src[4], intptr_t src_stride,
                  int mvx, int mvy, int w, int h, const orc_weight_t *wt )
{
    static const uint8_t ref0[16] = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1};
    static const uint8_t ref1[16] = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2};
    int qidx = ((mvy&3)<<2) + (mvx&3);
    intptr_t off = (intptr_t)(mvy>>2)*src_stride + (mvx>>2);
    const uint8_t *s1 = src[ref0[qidx]] + off + ((mvy&3) == 3)*src_stride;
    const uint8_t *s2 = src[ref1[qidx]] + off + ((mvx&3) == 3);
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            int v = s1[y*src_stride+x];
            if( qidx & 5 )
                v = ( v + s2[y*src_stride+x] + 1 ) >> 1;
            dst[y*dst_stride+x] = v;
        }
    if( wt && wt->enabled )
        orc_mc_weight( dst, dst_stride, dst, dst_stride, wt, w, h );
}

/* common/mc.c:251-283 mc_chroma: one plane (comp 0 = U, 1 = V) of an interleaved NV12 reference at an eighth-pel
 * chroma vector, bilinear weights (8-dx)(8-dy), dx(8-dy), (8-dx)dy, dx*dy, rounded (+32) >> 6 */
void orc_mc_chroma( uint8_t *dst, intptr_t dst_stride, const uint8_t *src_uv, intptr_t src_stride,
                    int mvx, int mvy, int w, int h, int comp )
{
    const int dx = mvx & 7, dy = mvy & 7;
    const int w00 = ( 8 - dx ) * ( 8 - dy ), w01 = dx * ( 8 - dy ), w10 = ( 8 - dx ) * dy, w11 = dx * dy;
    const uint8_t *row = src_uv + (intptr_t)( mvy >> 3 ) * src_stride + ( mvx >> 3 ) * 2 + comp;
    for( int y = 0; y < h; y++, row += src_stride )
        for( int x = 0; x < w; x++ )
        {
            const uint8_t *t = row + 2*x, *b = t + src_stride;
            dst[y*dst_stride + x] = ( w00*t[0] + w01*t[2] + w10*b[0] + w11*b[2] + 32 ) >> 6;
        }
}

/* common/mc.c:49-111 pixel_avg_WxH: weight==32 -> rounded mean, else implicit bipred weights */
void orc_pixel_avg( uint8_t *dst, intptr_t sd, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb,
                    int w, int h, int weight )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            int p = a[y*sa+x], q = b[y*sb+x];
            dst[y*sd+x] = weight == 32 ? ( p + q + 1 ) >> 1
                                       : clip_u8( ( p*weight + q*(64-weight) + 32 ) >> 6 );
        }
}

/* encoder/analyse.c:143-157, :179-188: cost_mv[i] = cost_mv[-i] = min( (int)(lambda*logs[i] + .5f), 65535 ),
 * logs[0] = 0.718f, logs[i] = log2f(i+1)*2 + 1.718f.  table has 2*len+1 entries, centre at table[len]. */
void orc_cost_mv_table( uint16_t *table, int len, int lambda )
{
    for( int i = 0; i <= len; i++ )
    {
        float l = i ? log2f( (float)(i+1) ) * 2.0f + 1.718f : 0.718f;
        int c = (int)( lambda * l + .5f );
        if( c > 65535 ) c = 65535;
        table[len+i] = table[len-i] = (uint16_t)c;
    }
}


/* ---- input staging: common/mc.c:294-339 (plane_copy, _swap, _interleave, _deinterleave) and x264_frame_copy_picture
 * (common/frame.c:363-480) for the 8-bit 4:2:0 colour spaces ---------------------------------------------------------------- */
void orc_plane_copy_interleave( uint8_t *dst, intptr_t sd, const uint8_t *u, intptr_t su, const uint8_t *v, intptr_t sv, int w, int h )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            dst[y*sd + 2*x]     = u[y*su + x];
            dst[y*sd + 2*x + 1] = v[y*sv + x];
        }
}

void orc_plane_copy_deinterleave( uint8_t *a, intptr_t sa, uint8_t *b, intptr_t sb, const uint8_t *src, intptr_t ss, int w, int h )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            a[y*sa + x] = src[y*ss + 2*x];
            b[y*sb + x] = src[y*ss + 2*x + 1];
        }
}

void orc_plane_copy_swap( uint8_t *dst, intptr_t sd, const uint8_t *src, intptr_t ss, int w, int h )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            dst[y*sd + 2*x]     = src[y*ss + 2*x + 1];
            dst[y*sd + 2*x + 1] = src[y*ss + 2*x];
        }
}

/* i_csp: X264_CSP_I420 2, YV12 3, NV12 4, NV21 5, | X264_CSP_VFLIP 0x1000 (x264.h:251-271); luma w x h, chroma (w/2 pairs) x (h/2) */
int orc_frame_copy_picture( int i_csp, const uint8_t *const plane[3], const int stride[3], int w, int h,
                            uint8_t *luma, intptr_t sl, uint8_t *chroma, intptr_t sc )
{
    const int csp = i_csp & 0xff, flip = ( i_csp & 0x1000 ) != 0, cw = w >> 1, ch = h >> 1;
    if( csp < 2 || csp > 5 ) return -1;
    const uint8_t *p[3]; intptr_t st[3];
    for( int i = 0; i < 3; i++ )
    {   /* get_plane_ptr, frame.c:342-358: a flipped plane starts at its last row and walks up */
        const int rows = i ? ch : h;
        p[i] = plane[i]; st[i] = stride[i];
        if( flip && p[i] ) { p[i] += (intptr_t)( rows - 1 ) * stride[i]; st[i] = -st[i]; }
    }
    for( int y = 0; y < h; y++ )
        memcpy( luma + y*sl, p[0] + y*st[0], w );
    if( csp == 4 )
        for( int y = 0; y < ch; y++ )
            memcpy( chroma + y*sc, p[1] + y*st[1], 2*cw );
    else if( csp == 5 )
        orc_plane_copy_swap( chroma, sc, p[1], st[1], cw, ch );
    else
    {
        const int iu = csp == 3 ? 2 : 1, iv = csp == 3 ? 1 : 2;
        orc_plane_copy_interleave( chroma, sc, p[iu], st[iu], p[iv], st[iv], cw, ch );
    }
    return 0;
}
