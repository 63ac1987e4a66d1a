/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product
 * (x264_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Plain-C restatement of the reference's pixel-metric table (jpsdr/x264 common/pixel.c), 8-bit depth.
 * Written from the algorithm, not from the source text: plain int32 arithmetic instead of the
 * reference's packed 2x16-bit "sum2_t" trick (the two are equal for 8-bit input -- every 4x4
 * coefficient magnitude is <= 4080 and every per-4x4 abs-sum <= 16320, so the packed lanes never
 * overflow; pinned against the compiled reference in tests/test_oracle_vs_ref.py).
 *
 * Parity status: PINNED against oracle/_ref/libx264ref.so (the unmodified reference compiled by
 * oracle/Makefile.ref) on checkasm-style random + worst-case buffers, and against tests/golden/.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

const int orc_pixel_w[ORC_PIXEL_NB] = { 16, 16, 8, 8, 8, 4, 4, 4 };
const int orc_pixel_h[ORC_PIXEL_NB] = { 16, 8, 16, 8, 4, 8, 4, 16 };

/* common/pixel.c:55-80 (PIXEL_SAD_C) */
int orc_sad( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h )
{
    int s = 0;
    for( int y = 0; y < h; y++, a += sa, b += sb )
        for( int x = 0; x < w; x++ )
            s += abs( (int)a[x] - (int)b[x] );
    return s;
}

/* common/pixel.c:85-110 (PIXEL_SSD_C) */
int orc_ssd( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h )
{
    int s = 0;
    for( int y = 0; y < h; y++, a += sa, b += sb )
        for( int x = 0; x < w; x++ )
        {
            int d = (int)a[x] - (int)b[x];
            s += d * d;
        }
    return s;
}

/* 4-point Hadamard butterfly, in place, on a strided vector */
static void had4( int *v, int st )
{
    int s01 = v[0] + v[st], d01 = v[0] - v[st];
    int s23 = v[2*st] + v[3*st], d23 = v[2*st] - v[3*st];
    v[0]    = s01 + s23;
    v[st]   = d01 + d23;
    v[2*st] = s01 - s23;
    v[3*st] = d01 - d23;
}

/* Sum |H4 * (a-b) * H4^T| over one 4x4 block (no normalisation).
 * common/pixel.c:262-286 (x264_pixel_satd_4x4 before the final >>1) */
static int satd4x4_raw( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int d[16];
    for( int y = 0; y < 4; y++ )
        for( int x = 0; x < 4; x++ )
            d[4*y+x] = (int)a[y*sa+x] - (int)b[y*sb+x];
    for( int y = 0; y < 4; y++ ) had4( d + 4*y, 1 );
    for( int x = 0; x < 4; x++ ) had4( d + x, 4 );
    int s = 0;
    for( int i = 0; i < 16; i++ ) s += abs( d[i] );
    return s;
}

/* common/pixel.c:262-332: W x H SATD = sum over 4x4 tiles of (raw>>1).  The reference sums 8x4
 * tiles as ((raw_left + raw_right) >> 1); each raw 4x4 sum is even, so the two agree. */
int orc_satd( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h )
{
    int s = 0;
    for( int y = 0; y < h; y += 4 )
        for( int x = 0; x < w; x += 4 )
            s += satd4x4_raw( a + y*sa + x, sa, b + y*sb + x, sb ) >> 1;
    return s;
}

/* common/pixel.c:334-367 (sa8d_8x8): 8x8 Hadamard abs-sum, un-normalised */
static int sa8d8x8_raw( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int d[64];
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
            d[8*y+x] = (int)a[y*sa+x] - (int)b[y*sb+x];
    /* 8-point Hadamard = three butterfly stages; any stage order gives the same multiset of |coef| */
    for( int pass = 0; pass < 2; pass++ )
    {
        int st = pass ? 8 : 1, ot = pass ? 1 : 8;
        for( int i = 0; i < 8; i++ )
        {
            int *v = d + i*ot;
            for( int span = 1; span < 8; span <<= 1 )
                for( int j = 0; j < 8; j++ )
                    if( !(j & span) )
                    {
                        int p = v[j*st], q = v[(j+span)*st];
                        v[j*st] = p + q;
                        v[(j+span)*st] = p - q;
                    }
        }
    }
    int s = 0;
    for( int i = 0; i < 64; i++ ) s += abs( d[i] );
    return s;
}

/* common/pixel.c:369-381: (sum+2)>>2 applied once, after summing the 8x8 quadrants */
int orc_sa8d( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb, int w, int h )
{
    int s = 0;
    for( int y = 0; y < h; y += 8 )
        for( int x = 0; x < w; x += 8 )
            s += sa8d8x8_raw( a + y*sa + x, sa, b + y*sb + x, sb );
    return ( s + 2 ) >> 2;
}

int orc_pixel_cmp( int metric, int i_pixel, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int w = orc_pixel_w[i_pixel], h = orc_pixel_h[i_pixel];
    switch( metric )
    {
        case ORC_SAD:  return orc_sad( a, sa, b, sb, w, h );
        case ORC_SSD:  return orc_ssd( a, sa, b, sb, w, h );
        case ORC_SATD: return orc_satd( a, sa, b, sb, w, h );
        case ORC_SA8D: return orc_sa8d( a, sa, b, sb, w, h );
    }
    return -1;
}

/* Batched twin used by the parity tests: candidate i compares the block at fenc+cand[i].fenc_off
 * with the block at ref+cand[i].ref_off (byte offsets; both planes share their own stride).
 * Mirrors the call shape of x264_pixel_cmp_t (common/pixel.h:33). */
void orc_pixel_cmp_batch( int metric, int i_pixel,
                          const uint8_t *fenc, intptr_t fenc_stride,
                          const uint8_t *ref, intptr_t ref_stride,
                          const orc_cand_t *cand, int n, int32_t *out )
{
    for( int i = 0; i < n; i++ )
        out[i] = orc_pixel_cmp( metric, i_pixel, fenc + cand[i].fenc_off, fenc_stride,
                                ref + cand[i].ref_off, ref_stride );
}

/* MV-field twin: block (bx,by) of the frame's WxH tiling compared with the reference block displaced
 * by the full-pel vector mv[k][by*bw+bx]. */
void orc_pixel_cmp_mvfield( int metric, int i_pixel,
                            const uint8_t *fenc, intptr_t fenc_stride,
                            const uint8_t *ref, intptr_t ref_stride,
                            int blocks_x, int blocks_y, int k_cands, const int16_t *mv, int32_t *out )
{
    int w = orc_pixel_w[i_pixel], h = orc_pixel_h[i_pixel];
    int nb = blocks_x * blocks_y;
    for( int k = 0; k < k_cands; k++ )
        for( int by = 0; by < blocks_y; by++ )
            for( int bx = 0; bx < blocks_x; bx++ )
            {
                int i = k*nb + by*blocks_x + bx;
                int mx = mv[2*i], my = mv[2*i+1];
                out[i] = orc_pixel_cmp( metric, i_pixel, fenc + (intptr_t)by*h*fenc_stride + bx*w, fenc_stride,
                                        ref + ((intptr_t)by*h + my)*ref_stride + bx*w + mx, ref_stride );
            }
}


/* ---- successive elimination: pixf.ads[] (common/pixel.c:759-803) and the integral image (common/mc.c:424-456, :748-783) ------ */
/* k = 1, 2 or 4 terms (x264_pixel_ads1 / ads2 / ads4); returns the number of passing positions, their indices in mvs[] */
int orc_pixel_ads( int k, const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx, int16_t *mvs, int width, int thresh )
{
    int nmv = 0;
    for( int i = 0; i < width; i++ )
    {
        int ads = abs( enc_dc[0] - sums[i] ) + cost_mvx[i];
        if( k == 2 ) ads += abs( enc_dc[1] - sums[i + delta] );
        if( k == 4 ) ads += abs( enc_dc[1] - sums[i + 8] ) + abs( enc_dc[2] - sums[i + delta] ) + abs( enc_dc[3] - sums[i + delta + 8] );
        if( ads < thresh )
            mvs[nmv++] = i;
    }
    return nmv;
}

/* The integral planes as x264_frame_filter builds them row by row: a horizontal running box (integral_init4h / 8h) added to the
 * row above (a vertical prefix, u16 wrap-around and all), then differences 4 / 8 rows apart (integral_init4v / 8v).  plane /
 * sum8 / sum4 point at position (0,0); the padded plane is `pad` wide on every side; stride in elements.  Rows and columns are
 * walked from -pad.  sum4 == NULL: the 8x8-only form (no sub-8x8 partitions). */
void orc_integral_init( const uint8_t *plane, intptr_t stride, int width, int height, int pad, uint16_t *sum8, uint16_t *sum4 )
{
    const int n = sum4 ? 4 : 8, cols = width + 2*pad - n;             /* positions per row: x = -pad .. -pad + cols */
    uint16_t *top = sum8 - (intptr_t)pad*stride - pad;
    for( int x = 0; x <= cols + n; x++ ) top[x] = 0;                  /* the row above the first one (mc.c:757-760) */
    for( int y = -pad; y < height + pad; y++ )
    {
        const uint8_t *pix = plane + (intptr_t)y*stride - pad;
        uint16_t *row = sum8 + (intptr_t)( y + 1 )*stride - pad;      /* prefix row y+1 = box row y + prefix row y */
        int v = 0;
        for( int i = 0; i < n; i++ ) v += pix[i];
        for( int x = 0; x <= cols; x++ )
        {
            row[x] = (uint16_t)( v + row[x - stride] );
            if( x < cols ) v += pix[x + n] - pix[x];
        }
        if( y < 8 - pad )
            continue;
        uint16_t *s8 = row - 8*stride;                                /* final row y-7 */
        if( sum4 )
        {
            uint16_t *s4 = sum4 + ( s8 - sum8 );
            for( int x = 0; x <= cols; x++ ) s4[x] = (uint16_t)( s8[x + 4*stride] - s8[x] );
            for( int x = 0; x <= cols - 4; x++ ) s8[x] = (uint16_t)( s8[x + 8*stride] + s8[x + 8*stride + 4] - s8[x] - s8[x + 4] );
        }
        else
            for( int x = 0; x <= cols; x++ ) s8[x] = (uint16_t)( s8[x + 8*stride] - s8[x] );
    }
}
