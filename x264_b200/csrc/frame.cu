// Frame preparation kernels: the streaming (HBM-bound) entries of x264_mc_functions_t.
//   lowres_fused_kernel : x264_frame_init_lowres + frame_init_lowres_core + expand_border_lowres
//                    (common/mc.c:458-507, common/frame.c:627-631)            2*W*H algorithmic bytes / frame
//   hpel_kernel    : hpel_filter over a frame + border expansion of the three planes
//                    (common/mc.c:172-196, :704-746, common/frame.c:596-625)   4*W*H algorithmic bytes / frame
// Both write the full padded domain directly (every output pixel, border included, is a pure function of the
// edge-clamped source), so no separate border pass is needed.
#include "ctx.h"

namespace {

__device__ __forceinline__ int clampi( int v, int lo, int hi ) { return min( max( v, lo ), hi ); }

// ------------------------------------------------------------------------------------------------
// lowres: thread = 4 horizontally adjacent output pixels of all four planes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void
lowres_border_body( int t, int pic, const uint8_t *__restrict__ src, intptr_t src_stride, int width, int height,
                    uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc, intptr_t dst_stride, int wl, int ll, int fast_ok,
                    int skip_x1, intptr_t src_pitch, intptr_t dst_pitch )
{
    // output domain incl. border: x in [-PAD, wl+PAD) in groups of 4, y in [-PAD, ll+PAD).  Columns [0, skip_x1) of the picture's
    // own rows are lowres_wide_body's, so the threads are numbered over what is left: the two bands of PAD rows above and below
    // the picture at full width, then per picture row the PAD/4 groups left of it and the groups from skip_x1 on.
    const int groups_x = ( wl + 2*X264CU_PAD ) / 4;
    const int band = 2 * X264CU_PAD * groups_x, left = X264CU_PAD / 4, per_row = groups_x - skip_x1 / 4;
    int gx, oy;
    if( t < band )
    {
        const int row = t / groups_x;
        gx = t - row * groups_x;
        oy = row < X264CU_PAD ? row - X264CU_PAD : ll + row - X264CU_PAD;
    }
    else
    {
        const int u = t - band;
        oy = u / per_row;
        if( oy >= ll ) return;
        const int k = u - oy * per_row;
        gx = k < left ? k : k + skip_x1 / 4;
    }
    const int ox = gx * 4 - X264CU_PAD;
    src += (intptr_t)pic * src_pitch;
    d0 += (intptr_t)pic * dst_pitch; dh += (intptr_t)pic * dst_pitch;
    dv += (intptr_t)pic * dst_pitch; dc += (intptr_t)pic * dst_pitch;
    const int y = clampi( oy, 0, ll-1 );
    if( fast_ok && ox >= 0 && ox + 4 <= wl && 2*( ox + 4 ) + 1 <= width )
    {   // columns 2ox .. 2ox+8 inside the picture (rows are clamped: the rows of the top / bottom border repeat the edge rows'
        // results, the last source row repeats below the picture).  8-byte aligned vector loads; the whole filter runs on packed
        // bytes: FILTER(a,b,c,d) = avg( avg(a,b), avg(c,d) ) with avg(x,y) = (x+y+1)>>1 = __vavgu4 (mc.c:494-500)
        uint32_t lo[3], hi[3], nx[3];
#pragma unroll
        for( int r = 0; r < 3; r++ )
        {
            const uint8_t *p = src + (intptr_t)min( 2*y + r, height-1 ) * src_stride + 2*ox;
            const uint2 v = __ldg( (const uint2 *)p );
            lo[r] = v.x; hi[r] = v.y; nx[r] = __ldg( p + 8 );
        }
        // vertical means of rows (0,1) and (1,2), columns 0..8
        const uint32_t a_lo = __vavgu4( lo[0], lo[1] ), a_hi = __vavgu4( hi[0], hi[1] ), a_nx = __vavgu4( nx[0], nx[1] );
        const uint32_t b_lo = __vavgu4( lo[1], lo[2] ), b_hi = __vavgu4( hi[1], hi[2] ), b_nx = __vavgu4( nx[1], nx[2] );
        // columns {0,2,4,6}, {1,3,5,7}, {2,4,6,8}
        const uint32_t a_ev = __byte_perm( a_lo, a_hi, 0x6420 ), a_od = __byte_perm( a_lo, a_hi, 0x7531 );
        const uint32_t b_ev = __byte_perm( b_lo, b_hi, 0x6420 ), b_od = __byte_perm( b_lo, b_hi, 0x7531 );
        const uint32_t a_e2 = __byte_perm( a_ev, a_nx, 0x4321 ), b_e2 = __byte_perm( b_ev, b_nx, 0x4321 );
        *(uint32_t *)( d0 + (intptr_t)oy*dst_stride + ox ) = __vavgu4( a_ev, a_od );
        *(uint32_t *)( dh + (intptr_t)oy*dst_stride + ox ) = __vavgu4( a_od, a_e2 );
        *(uint32_t *)( dv + (intptr_t)oy*dst_stride + ox ) = __vavgu4( b_ev, b_od );
        *(uint32_t *)( dc + (intptr_t)oy*dst_stride + ox ) = __vavgu4( b_od, b_e2 );
        return;
    }
    // edges and border: per-pixel clamped coordinates (the picture is edge-replicated to the mod-16 size and one
    // column / row beyond: mc.c:466-469, frame.c:640-665; the lowres border replicates the computed edge: frame.c:627)
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    for( int i = 0; i < 4; i++ )
    {
        const int x = clampi( ox + i, 0, wl-1 );
        int s[3][3];
#pragma unroll
        for( int r = 0; r < 3; r++ )
#pragma unroll
            for( int c = 0; c < 3; c++ )
                s[r][c] = src[(intptr_t)min( 2*y + r, height-1 ) * src_stride + min( 2*x + c, width-1 )];
#define FILT( a, b, c, d ) ( ( ( ( (a) + (b) + 1 ) >> 1 ) + ( ( (c) + (d) + 1 ) >> 1 ) + 1 ) >> 1 )
        o0 |= (uint32_t)FILT( s[0][0], s[1][0], s[0][1], s[1][1] ) << ( 8*i );
        o1 |= (uint32_t)FILT( s[0][1], s[1][1], s[0][2], s[1][2] ) << ( 8*i );
        o2 |= (uint32_t)FILT( s[1][0], s[2][0], s[1][1], s[2][1] ) << ( 8*i );
        o3 |= (uint32_t)FILT( s[1][1], s[2][1], s[1][2], s[2][2] ) << ( 8*i );
#undef FILT
    }
    *(uint32_t *)( d0 + (intptr_t)oy*dst_stride + ox ) = o0;
    *(uint32_t *)( dh + (intptr_t)oy*dst_stride + ox ) = o1;
    *(uint32_t *)( dv + (intptr_t)oy*dst_stride + ox ) = o2;
    *(uint32_t *)( dc + (intptr_t)oy*dst_stride + ox ) = o3;
}


// ------------------------------------------------------------------------------------------------
// lowres, wide path (16-byte aligned source rows): thread = 8 output pixels x 2 output rows of all four planes -- five source
// rows of 16 + 1 bytes (128-bit loads), eight 64-bit stores; the filter itself as in lowres_border_body, on packed bytes.
// Interior of the picture only; the border and the picture's last columns stay with lowres_border_body (the grid's last block rows).
// blockIdx.z = picture of a stack.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lowres_row8( const uint32_t a[5], const uint32_t b[5], uint8_t *p0, uint8_t *ph, uint8_t *pv, uint8_t *pc )
{   // a = vertical mean of source rows (2y, 2y+1), b = of (2y+1, 2y+2): words 0..3 = columns 0..15, word 4 byte 0 = column 16
    const uint32_t a_ev0 = __byte_perm( a[0], a[1], 0x6420 ), a_od0 = __byte_perm( a[0], a[1], 0x7531 );
    const uint32_t a_ev1 = __byte_perm( a[2], a[3], 0x6420 ), a_od1 = __byte_perm( a[2], a[3], 0x7531 );
    const uint32_t b_ev0 = __byte_perm( b[0], b[1], 0x6420 ), b_od0 = __byte_perm( b[0], b[1], 0x7531 );
    const uint32_t b_ev1 = __byte_perm( b[2], b[3], 0x6420 ), b_od1 = __byte_perm( b[2], b[3], 0x7531 );
    const uint32_t a_e20 = __byte_perm( a_ev0, a_ev1, 0x4321 ), a_e21 = __byte_perm( a_ev1, a[4], 0x4321 );
    const uint32_t b_e20 = __byte_perm( b_ev0, b_ev1, 0x4321 ), b_e21 = __byte_perm( b_ev1, b[4], 0x4321 );
    *(uint2 *)p0 = make_uint2( __vavgu4( a_ev0, a_od0 ), __vavgu4( a_ev1, a_od1 ) );
    *(uint2 *)ph = make_uint2( __vavgu4( a_od0, a_e20 ), __vavgu4( a_od1, a_e21 ) );
    *(uint2 *)pv = make_uint2( __vavgu4( b_ev0, b_od0 ), __vavgu4( b_ev1, b_od1 ) );
    *(uint2 *)pc = make_uint2( __vavgu4( b_od0, b_e20 ), __vavgu4( b_od1, b_e21 ) );
}

__device__ __forceinline__ void
lowres_wide_body( int g, int y, int pic, const uint8_t *__restrict__ src, intptr_t src_stride, intptr_t src_pitch, int height,
                  uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc, intptr_t dst_stride, intptr_t dst_pitch, int groups, int ll )
{   // output rows y, y+1 (ll is even), columns 8g .. 8g+7
    if( g >= groups ) return;
    src += (intptr_t)pic * src_pitch;
    const intptr_t doff = (intptr_t)pic * dst_pitch + (intptr_t)y * dst_stride + 8 * g;
    uint32_t r[5][5];
#pragma unroll
    for( int k = 0; k < 5; k++ )
    {
        const uint8_t *p = src + (intptr_t)min( 2 * y + k, height - 1 ) * src_stride + 16 * g;
        const uint4 v = __ldg( (const uint4 *)p );
        r[k][0] = v.x; r[k][1] = v.y; r[k][2] = v.z; r[k][3] = v.w;
        r[k][4] = __ldg( (const uint32_t *)( p + 16 ) );
    }
    uint32_t m[4][5];                                              // vertical means of consecutive source rows
#pragma unroll
    for( int k = 0; k < 4; k++ )
#pragma unroll
        for( int i = 0; i < 5; i++ ) m[k][i] = __vavgu4( r[k][i], r[k + 1][i] );
    lowres_row8( m[0], m[1], d0 + doff, dh + doff, dv + doff, dc + doff );
    if( y + 1 < ll )
        lowres_row8( m[2], m[3], d0 + doff + dst_stride, dh + doff + dst_stride, dv + doff + dst_stride, dc + doff + dst_stride );
}

// ONE launch for a stack of pictures: block rows [0, wide_rows) of the grid run the wide body (one block row = two output rows),
// the block rows after them are numbered linearly over the border tasks.  The border's scattered, mostly scalar work (7 % of the
// output) then runs under the streaming blocks instead of after them.
__global__ void __launch_bounds__( 128 )
lowres_fused_kernel( const uint8_t *__restrict__ src, intptr_t src_stride, intptr_t src_pitch, int width, int height,
                     uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc, intptr_t dst_stride, intptr_t dst_pitch,
                     int groups, int wl, int ll, int fast_ok, int wide_rows )
{
    if( (int)blockIdx.y < wide_rows )
        lowres_wide_body( blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * 2, blockIdx.z, src, src_stride, src_pitch, height,
                          d0, dh, dv, dc, dst_stride, dst_pitch, groups, ll );
    else
        lowres_border_body( ( ( blockIdx.y - wide_rows ) * gridDim.x + blockIdx.x ) * blockDim.x + threadIdx.x, blockIdx.z,
                            src, src_stride, width, height, d0, dh, dv, dc, dst_stride, wl, ll, fast_ok, 8 * groups, src_pitch, dst_pitch );
}

// ------------------------------------------------------------------------------------------------
// hpel: CTA = 64x16 output tile; the source tile (+3/+2 halo) is staged in shared memory, the vertical
// 6-tap sums are kept as int16 in shared memory for the centre plane.
// ------------------------------------------------------------------------------------------------
constexpr int HT_W = 64, HT_H = 16;

__global__ void __launch_bounds__( 256 )
hpel_kernel( const uint8_t *__restrict__ src, intptr_t stride, int width, int height,
             uint8_t *dh, uint8_t *dv, uint8_t *dc, uint8_t *dsrc_border )
{
    __shared__ uint8_t s_src[HT_H + 5][HT_W + 8];       // rows y-2..y+3, cols x-2..x+3 (+pad)
    __shared__ int16_t s_v[HT_H][HT_W + 8];             // vertical tap sums for cols x-2..x+3
    const int x0 = blockIdx.x * HT_W - X264CU_PAD, y0 = blockIdx.y * HT_H - X264CU_PAD;
    for( int i = threadIdx.x; i < ( HT_H + 5 ) * ( HT_W + 5 ); i += blockDim.x )
    {
        int r = i / ( HT_W + 5 ), c = i - r * ( HT_W + 5 );
        int sy = clampi( y0 + r - 2, 0, height-1 ), sx = clampi( x0 + c - 2, 0, width-1 );
        s_src[r][c] = src[(intptr_t)sy * stride + sx];
    }
    __syncthreads();
    for( int i = threadIdx.x; i < HT_H * ( HT_W + 5 ); i += blockDim.x )
    {
        int r = i / ( HT_W + 5 ), c = i - r * ( HT_W + 5 );
        int v = s_src[r][c] + s_src[r+5][c] - 5 * ( s_src[r+1][c] + s_src[r+4][c] ) + 20 * ( s_src[r+2][c] + s_src[r+3][c] );
        s_v[r][c] = (int16_t)v;
    }
    __syncthreads();
    const int full_w = width + 2*X264CU_PAD, full_h = height + 2*X264CU_PAD;
    for( int i = threadIdx.x; i < HT_H * HT_W; i += blockDim.x )
    {
        int r = i / HT_W, c = i - r * HT_W;
        int ox = x0 + c, oy = y0 + r;
        if( ox + X264CU_PAD >= full_w || oy + X264CU_PAD >= full_h ) continue;
        const uint8_t *row = &s_src[r+2][c];            // row[k] = src(x-2+k, y)
        int hsum = row[0] + row[5] - 5 * ( row[1] + row[4] ) + 20 * ( row[2] + row[3] );
        const int16_t *vr = &s_v[r][c];                 // vr[k] = vsum(x-2+k, y)
        int vsum = vr[2];
        int csum = vr[0] + vr[5] - 5 * ( vr[1] + vr[4] ) + 20 * ( vr[2] + vr[3] );
        intptr_t o = (intptr_t)oy * stride + ox;
        dh[o] = (uint8_t)clampi( ( hsum + 16 ) >> 5, 0, 255 );
        dv[o] = (uint8_t)clampi( ( vsum + 16 ) >> 5, 0, 255 );
        dc[o] = (uint8_t)clampi( ( csum + 512 ) >> 10, 0, 255 );
        if( dsrc_border && ( ox < 0 || ox >= width || oy < 0 || oy >= height ) )
            dsrc_border[o] = row[2];
    }
}

// ------------------------------------------------------------------------------------------------
// hpel, word path (4-byte aligned planes, width % 4 == 0): CTA = 128x16 output tile of the padded domain.  The source tile
// (+halo) is staged as 32-bit words (interior tiles) or byte by byte with clamped coordinates (tiles touching the border: the
// padded domain of the three planes is the filter of the edge-replicated source, frame.c:596-625); a thread produces 4
// horizontally adjacent pixels per step: first the vertical 6-tap sums (V plane + 16-bit intermediates in shared memory,
// mc.c:176-183), then the horizontal taps over the source row (H) and over the intermediates (C, mc.c:184-193).
// ------------------------------------------------------------------------------------------------
constexpr int FT_W = 128, FT_H = 16, FT_SW = 144, FT_WORDS = 34;     // smem row: columns x0-4 .. x0+131 = 34 words (+pad)

__device__ __forceinline__ int tap6( int a, int b, int c, int d, int e, int f ) { return a + f - 5 * ( b + e ) + 20 * ( c + d ); }
__device__ __forceinline__ uint32_t clip4( int a, int b, int c, int d, int add, int sh )
{
    a = clampi( ( a + add ) >> sh, 0, 255 ); b = clampi( ( b + add ) >> sh, 0, 255 );
    c = clampi( ( c + add ) >> sh, 0, 255 ); d = clampi( ( d + add ) >> sh, 0, 255 );
    return (uint32_t)a | ( (uint32_t)b << 8 ) | ( (uint32_t)c << 16 ) | ( (uint32_t)d << 24 );
}

__global__ void __launch_bounds__( 256 )
hpel_words_kernel( const uint8_t *__restrict__ src, intptr_t stride, int width, int height,
                   uint8_t *dh, uint8_t *dv, uint8_t *dc, uint8_t *dsrc_border )
{
    __shared__ __align__( 16 ) uint8_t s_src[FT_H + 5][FT_SW];       // rows y0-2 .. y0+18, columns x0-4 .. x0+131
    __shared__ __align__( 16 ) int16_t s_v[FT_H][FT_SW];             // vertical tap sums, same columns
    const int x0 = blockIdx.x * FT_W - X264CU_PAD, y0 = blockIdx.y * FT_H - X264CU_PAD;
    const bool interior = x0 - 4 >= 0 && x0 + FT_W + 4 <= width && y0 - 2 >= 0 && y0 + FT_H + 3 <= height;
    if( interior )
    {
        for( int i = threadIdx.x; i < ( FT_H + 5 ) * FT_WORDS; i += blockDim.x )
        {
            const int r = i / FT_WORDS, c = i - r * FT_WORDS;
            *(uint32_t *)&s_src[r][4 * c] = __ldg( (const uint32_t *)( src + (intptr_t)( y0 + r - 2 ) * stride + x0 - 4 ) + c );
        }
    }
    else
    {
        for( int i = threadIdx.x; i < ( FT_H + 5 ) * FT_WORDS * 4; i += blockDim.x )
        {
            const int r = i / ( FT_WORDS * 4 ), c = i - r * ( FT_WORDS * 4 );
            const int sy = clampi( y0 + r - 2, 0, height - 1 ), sx = clampi( x0 + c - 4, 0, width - 1 );
            s_src[r][c] = src[(intptr_t)sy * stride + sx];
        }
    }
    __syncthreads();
    const int full_w = width + 2 * X264CU_PAD, full_h = height + 2 * X264CU_PAD;
    // vertical sums of every staged column (the C plane needs 2 / 3 columns beyond the tile), V plane of the tile's own columns
    for( int i = threadIdx.x; i < FT_H * FT_WORDS; i += blockDim.x )
    {
        const int r = i / FT_WORDS, wc = i - r * FT_WORDS;
        uint32_t w[6];
#pragma unroll
        for( int k = 0; k < 6; k++ ) w[k] = *(const uint32_t *)&s_src[r + k][4 * wc];
        int v[4];
#pragma unroll
        for( int b = 0; b < 4; b++ )
            v[b] = tap6( ( w[0] >> ( 8*b ) ) & 255, ( w[1] >> ( 8*b ) ) & 255, ( w[2] >> ( 8*b ) ) & 255,
                         ( w[3] >> ( 8*b ) ) & 255, ( w[4] >> ( 8*b ) ) & 255, ( w[5] >> ( 8*b ) ) & 255 );
        *(uint2 *)&s_v[r][4 * wc] = make_uint2( ( v[0] & 0xffff ) | ( (uint32_t)v[1] << 16 ), ( v[2] & 0xffff ) | ( (uint32_t)v[3] << 16 ) );
        const int ox = x0 + 4 * ( wc - 1 ), oy = y0 + r;
        if( wc >= 1 && wc <= FT_W / 4 && ox + X264CU_PAD < full_w && oy + X264CU_PAD < full_h )
            *(uint32_t *)( dv + (intptr_t)oy * stride + ox ) = clip4( v[0], v[1], v[2], v[3], 16, 5 );
    }
    __syncthreads();
    for( int i = threadIdx.x; i < FT_H * ( FT_W / 4 ); i += blockDim.x )
    {
        const int r = i / ( FT_W / 4 ), g = i - r * ( FT_W / 4 );
        const int ox = x0 + 4 * g, oy = y0 + r;
        if( ox + X264CU_PAD >= full_w || oy + X264CU_PAD >= full_h ) continue;
        // source row of the output row, columns ox-4 .. ox+7 (bytes 0..11; the pixel of output k is byte 4+k)
        const uint32_t *sw = (const uint32_t *)&s_src[r + 2][4 * g];
        const uint32_t s0 = sw[0], s1 = sw[1], s2 = sw[2];
        int b[12];
#pragma unroll
        for( int k = 0; k < 4; k++ ) { b[k] = ( s0 >> ( 8*k ) ) & 255; b[4 + k] = ( s1 >> ( 8*k ) ) & 255; b[8 + k] = ( s2 >> ( 8*k ) ) & 255; }
        int h[4], c[4];
#pragma unroll
        for( int k = 0; k < 4; k++ ) h[k] = tap6( b[2 + k], b[3 + k], b[4 + k], b[5 + k], b[6 + k], b[7 + k] );
        const uint32_t *vw = (const uint32_t *)&s_v[r][4 * g];        // 12 int16: columns ox-4 .. ox+7
        int v[12];
#pragma unroll
        for( int k = 0; k < 6; k++ ) { const uint32_t t = vw[k]; v[2*k] = (int16_t)( t & 0xffff ); v[2*k + 1] = (int16_t)( t >> 16 ); }
#pragma unroll
        for( int k = 0; k < 4; k++ ) c[k] = tap6( v[2 + k], v[3 + k], v[4 + k], v[5 + k], v[6 + k], v[7 + k] );
        const intptr_t o = (intptr_t)oy * stride + ox;
        *(uint32_t *)( dh + o ) = clip4( h[0], h[1], h[2], h[3], 16, 5 );
        *(uint32_t *)( dc + o ) = clip4( c[0], c[1], c[2], c[3], 512, 10 );
        if( dsrc_border && ( ox < 0 || ox >= width || oy < 0 || oy >= height ) )
            *(uint32_t *)( dsrc_border + o ) = s1;
    }
}

// ------------------------------------------------------------------------------------------------
// hpel, packed path (16-byte aligned planes, width % 4 == 0): CTA = 128 x 32 output tile of the padded domain, 128 threads.
// All arithmetic runs on packed data: the vertical taps on 16-bit fields, two pixels per 32-bit word (no carries between fields:
// every partial sum stays within 16 bits once biased by 4112), the horizontal taps with dp4a on the bytes (H) and dp2a on the
// biased 16-bit vertical sums (C), clamping with the packed min/max instructions (VIMNMX / VIADDMNMX .S16x2).
//   phase 1  stage rows y0-2 .. y0+34, columns x0-16 .. x0+143: one TMA load of the box (cp.async.bulk.tensor, completion on an
//            mbarrier) where it lies inside the picture; the tiles at the picture's edges stage 16-byte vectors with clamped rows
//   phase 2  a thread walks 16 output rows of one 4-pixel column group down a sliding window of 6 unpacked rows: V plane + the
//            biased vertical sums (v + 4112) into shared memory (mc.c:176-183)
//   phase 3  H from the source row, C from the vertical sums (mc.c:184-193); 8 pixels per thread and step
// ------------------------------------------------------------------------------------------------
constexpr int PT_W = 128, PT_H = 32, PT_ROWS = PT_H + 5, PT_PITCH = 160, PT_VW = 72, PT_THREADS = 128;
#ifndef HPEL_CHUNK
#define HPEL_CHUNK 16
#endif
constexpr int PT_CHUNK = HPEL_CHUNK, PT_MAIN = 32 * ( PT_H / PT_CHUNK );   // rows per phase-2 walk; the threads walking
constexpr uint32_t PT_BIAS = 4112u * 0x00010001u;              // 4096 + 16: the rounding of the V plane rides along

__device__ __forceinline__ int dp4a_u8s8( uint32_t a, int32_t b, int32_t c )
{
    int d;
    asm( "dp4a.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}

// bytes 0 and 2 from a, bytes 1 and 3 from b
__device__ __forceinline__ uint32_t byte_lanes( uint32_t a, uint32_t b )
{
    uint32_t d;
    asm( "lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( 0x00ff00ffu ) );
    return d;
}

// 32 registers, 15 CTAs per SM (shared memory allows no more): measured cold at 4K 7.87 us per picture against 8.03 us with the
// compiler's own choice (40 registers, 12 CTAs)
#ifndef HPEL_MIN_CTAS
#define HPEL_MIN_CTAS 15
#endif
__global__ void __launch_bounds__( PT_THREADS, HPEL_MIN_CTAS )
hpel_packed_kernel( const __grid_constant__ CUtensorMap tm_src, const uint8_t *__restrict__ src, intptr_t stride, intptr_t pitch, int width, int height,
                    uint8_t *dh, uint8_t *dv, uint8_t *dc, uint8_t *dsrc_border, uint32_t less4096 )
{   // less4096 = -4096 in both 16-bit fields: a kernel parameter, so that VIADDMNMX (one immediate only) reads it from the constant bank
    __shared__ __align__( 128 ) uint8_t s_src[PT_ROWS][PT_PITCH];
    __shared__ __align__( 16 ) uint32_t s_v[PT_H][PT_VW];          // two biased vertical sums per word
    // a stack of pictures: blockIdx.z selects the plane (all five planes share stride and pitch)
    const intptr_t poff = (intptr_t)blockIdx.z * pitch;
    src += poff; dh += poff; dv += poff; dc += poff;
    if( dsrc_border ) dsrc_border += poff;
    const int x0 = blockIdx.x * PT_W - X264CU_PAD, y0 = blockIdx.y * PT_H - X264CU_PAD;
    // the staged rectangle inside the picture: one TMA load of the whole box (the tensor map's out-of-bounds fill is zero, not
    // the edge pixel, so the tiles at the picture's edges take the other path)
    const bool inside = x0 - 16 >= 0 && x0 + PT_W + 16 <= width && y0 - 2 >= 0 && y0 + PT_ROWS - 2 <= height;
    if( inside )
    {
        __shared__ __align__( 8 ) uint64_t s_bar;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared( &s_bar );
        if( threadIdx.x == 0 )
        {
            asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"( bar ) );
            asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
            asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( bar ), "r"( PT_ROWS * PT_PITCH ) : "memory" );
            asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                          ::"r"( (uint32_t)__cvta_generic_to_shared( &s_src[0][0] ) ), "l"( &tm_src ), "r"( bar ),
                            "r"( x0 - 16 ), "r"( y0 - 2 ), "r"( (int)blockIdx.z ) : "memory" );
        }
        __syncthreads();                                           // the barrier's initialisation, for the waiting threads
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "HPEL_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra HPEL_DONE_%=;\n"
            "bra HPEL_WAIT_%=;\n"
            "HPEL_DONE_%=:\n"
            "}\n" ::"r"( bar ) : "memory" );
    }
    else
    {   // tiles at the picture's edges (rows clamped): 16-byte vectors where they lie inside the row, the edge pixel repeated where
        // they lie outside, byte by byte where they straddle the edge
        for( int i = threadIdx.x; i < PT_ROWS * ( PT_PITCH / 16 ); i += PT_THREADS )
        {
            const int r = i / ( PT_PITCH / 16 ), c = i - r * ( PT_PITCH / 16 );
            const int sy = clampi( y0 + r - 2, 0, height - 1 ), lo = x0 - 16 + 16 * c;
            const uint8_t *row = src + (intptr_t)sy * stride;
            uint4 v;
            if( lo >= 0 && lo + 16 <= width )
                v = __ldg( (const uint4 *)( row + lo ) );
            else if( lo + 16 <= 0 || lo >= width )
                v.x = v.y = v.z = v.w = 0x01010101u * row[lo < 0 ? 0 : width - 1];
            else
            {
                uint32_t w[4];
#pragma unroll
                for( int k = 0; k < 4; k++ )
                    w[k] = row[clampi( lo + 4*k, 0, width - 1 )] | ( row[clampi( lo + 4*k + 1, 0, width - 1 )] << 8 ) |
                           ( row[clampi( lo + 4*k + 2, 0, width - 1 )] << 16 ) | ( (uint32_t)row[clampi( lo + 4*k + 3, 0, width - 1 )] << 24 );
                v = make_uint4( w[0], w[1], w[2], w[3] );
            }
            *(uint4 *)&s_src[r][16 * c] = v;
        }
        __syncthreads();
    }
    const int full_w = width + 2 * X264CU_PAD, full_h = height + 2 * X264CU_PAD;
    // ---- phase 2: word column wc covers x0-4+4wc .. +3 (wc = 0 .. 33).  Warps 0-1: the tile's own 32 columns, warp = rows
    // 16*warp .. +15 down a sliding window (the longer the walk, the fewer window refills); warps 2-3: the two columns either side
    // that only the C plane's taps read, one row each
    if( threadIdx.x < PT_MAIN )
    {
        const int chunk = threadIdx.x >> 5, wc = 1 + ( threadIdx.x & 31 );
        const int r0 = chunk * PT_CHUNK;
        uint32_t lo[6], hi[6];                                       // the window: rows r .. r+5, pixels (0,1) and (2,3) as 16-bit fields
#pragma unroll
        for( int k = 0; k < 5; k++ )
        {
            const uint32_t w = *(const uint32_t *)&s_src[r0 + k][12 + 4 * wc];
            lo[k + 1] = __byte_perm( w, 0, 0x4140 ); hi[k + 1] = __byte_perm( w, 0, 0x4342 );
        }
        const int ox = x0 + 4 * ( wc - 1 );
        const int rows_v = ox + X264CU_PAD < full_w ? full_h - X264CU_PAD - ( y0 + r0 ) : 0;      // rows of this chunk inside the domain
        uint8_t *pv = dv + (intptr_t)( y0 + r0 ) * stride + ox;
        asm( "" : "+l"( pv ) );                                      // one 64-bit row pointer carried down the rows
#pragma unroll
        for( int j = 0; j < PT_CHUNK; j++ )
        {
#pragma unroll
            for( int k = 0; k < 5; k++ ) { lo[k] = lo[k + 1]; hi[k] = hi[k + 1]; }
            const uint32_t w = *(const uint32_t *)&s_src[r0 + j + 5][12 + 4 * wc];
            lo[5] = __byte_perm( w, 0, 0x4140 ); hi[5] = __byte_perm( w, 0, 0x4342 );
            // (a + f) + 20 (c + d) - 5 (b + e), both fields at once
            const uint32_t vl = ( lo[2] + lo[3] ) * 20u + ( lo[0] + lo[5] + PT_BIAS ) - ( lo[1] + lo[4] ) * 5u;
            const uint32_t vh = ( hi[2] + hi[3] ) * 20u + ( hi[0] + hi[5] + PT_BIAS ) - ( hi[1] + hi[4] ) * 5u;
            *(uint2 *)&s_v[r0 + j][2 * wc] = make_uint2( vl, vh );
            // clamp( v + 16, 0, 8191 ) per field (the bias less 4096); times 8, the result >> 5 is the field's high byte
            const uint32_t pl = __viaddmin_s16x2_relu( vl, less4096, 0x1fff1fffu ) << 3;
            const uint32_t ph = __viaddmin_s16x2_relu( vh, less4096, 0x1fff1fffu ) << 3;
            if( j < rows_v )
                *(uint32_t *)pv = __byte_perm( pl, ph, 0x7531 );
            pv += stride;
        }
    }
    else if( threadIdx.x < PT_MAIN + 64 )
    {
        const int e = threadIdx.x - PT_MAIN, r = e & 31, wc = e < 32 ? 0 : 33;
        uint32_t vl = PT_BIAS, vh = PT_BIAS;
#pragma unroll
        for( int k = 0; k < 6; k++ )
        {
            const uint32_t w = *(const uint32_t *)&s_src[r + k][12 + 4 * wc];
            const uint32_t tap = k == 0 || k == 5 ? 1u : k == 1 || k == 4 ? (uint32_t)-5 : 20u;
            vl += __byte_perm( w, 0, 0x4140 ) * tap; vh += __byte_perm( w, 0, 0x4342 ) * tap;
        }
        *(uint2 *)&s_v[r][2 * wc] = make_uint2( vl, vh );
    }
    __syncthreads();
    // ---- phase 3: 8 pixels per thread and step.  H from the source row, C from the vertical sums
    // (128 threads = 8 rows of 16 groups: a thread keeps its group of columns and steps 8 rows down)
    const int g = threadIdx.x & 15, ox = x0 + 8 * g;
    const bool second = ox + 4 + X264CU_PAD < full_w;               // the domain's width is a multiple of 4, not of 8
    if( ox + X264CU_PAD >= full_w ) return;
    const bool col_out = ox < 0 || ox >= width, col_out2 = ox < 0 || ox + 4 >= width;
    // only the tiles that reach outside the picture have source border to write
    const bool border_tile = dsrc_border && ( x0 < 0 || x0 + PT_W > width || y0 < 0 || y0 + PT_H > height );
    const int rb = threadIdx.x >> 4;
    const intptr_t o0 = (intptr_t)( y0 + rb ) * stride + ox, step = (intptr_t)( PT_THREADS / 16 ) * stride;
    uint8_t *ph = dh + o0, *pc = dc + o0;
#pragma unroll
    for( int r = rb; r < PT_H; r += PT_THREADS / 16, ph += step, pc += step )
    {
        const int oy = y0 + r;
        if( oy + X264CU_PAD >= full_h ) break;
        // H: taps (1,-5,20,20,-5,1); A[k] = the four bytes from column x+k-2 on, out of the 16 bytes x-4 .. x+11
        const uint32_t *sw = (const uint32_t *)&s_src[r + 2][12 + 8 * g];
        const uint32_t w0 = sw[0], w3 = sw[3];
        const uint2 w12 = *(const uint2 *)( sw + 1 );
        const uint32_t w1 = w12.x, w2 = w12.y;
        uint32_t A[12];
        A[0] = __funnelshift_r( w0, w1, 16 ); A[1] = __funnelshift_r( w0, w1, 24 ); A[2] = w1; A[3] = __funnelshift_r( w1, w2, 8 );
        A[4] = __funnelshift_r( w1, w2, 16 ); A[5] = __funnelshift_r( w1, w2, 24 ); A[6] = w2; A[7] = __funnelshift_r( w2, w3, 8 );
        A[8] = __funnelshift_r( w2, w3, 16 ); A[9] = __funnelshift_r( w2, w3, 24 ); A[10] = w3; A[11] = w3 >> 8;
        const int T0123 = 0x1414FB01, T45 = 0x000001FB;
        int hv[8];
#pragma unroll
        for( int k = 0; k < 8; k++ ) hv[k] = dp4a_u8s8( A[k], T0123, dp4a_u8s8( A[k + 4], T45, 16 ) );
        // C: the same taps over the biased vertical sums; q[j] = the fields of columns x-4+2j, x-3+2j
        const uint4 qa = *(const uint4 *)&s_v[r][4 * g], qb = *(const uint4 *)&s_v[r][4 * g + 4];
        const uint32_t q1 = qa.y, q2 = qa.z, q3 = qa.w, q4 = qb.x, q5 = qb.y, q6 = qb.z, q7 = qb.w;
        const uint32_t f12 = __funnelshift_r( q1, q2, 16 ), f23 = __funnelshift_r( q2, q3, 16 ), f34 = __funnelshift_r( q3, q4, 16 ),
                       f45 = __funnelshift_r( q4, q5, 16 ), f56 = __funnelshift_r( q5, q6, 16 ), f67 = __funnelshift_r( q6, q7, 16 );
        const int U01 = 0x0000FB01, U23 = 0x00001414, U45 = 0x000001FB, C0 = 512 - 32 * 4112;
        int cv[8];
        cv[0] = __dp2a_lo( (int)q1, U01, __dp2a_lo( (int)q2, U23, __dp2a_lo( (int)q3, U45, C0 ) ) );
        cv[1] = __dp2a_lo( (int)f12, U01, __dp2a_lo( (int)f23, U23, __dp2a_lo( (int)f34, U45, C0 ) ) );
        cv[2] = __dp2a_lo( (int)q2, U01, __dp2a_lo( (int)q3, U23, __dp2a_lo( (int)q4, U45, C0 ) ) );
        cv[3] = __dp2a_lo( (int)f23, U01, __dp2a_lo( (int)f34, U23, __dp2a_lo( (int)f45, U45, C0 ) ) );
        cv[4] = __dp2a_lo( (int)q3, U01, __dp2a_lo( (int)q4, U23, __dp2a_lo( (int)q5, U45, C0 ) ) );
        cv[5] = __dp2a_lo( (int)f34, U01, __dp2a_lo( (int)f45, U23, __dp2a_lo( (int)f56, U45, C0 ) ) );
        cv[6] = __dp2a_lo( (int)q4, U01, __dp2a_lo( (int)q5, U23, __dp2a_lo( (int)q6, U45, C0 ) ) );
        cv[7] = __dp2a_lo( (int)f45, U01, __dp2a_lo( (int)f56, U23, __dp2a_lo( (int)f67, U45, C0 ) ) );
        // four results to four bytes: pixels (0,2) and (1,3) share a word of 16-bit fields, so that the clamped fields shift straight
        // into their byte lanes.  H: clamp( h + 16, 0, 8191 ) >> 5.  C: bytes 1-2 of c + 512 are c >> 8; clamp( ., 0, 1023 ) >> 2
        uint32_t ho[2], co[2];
#pragma unroll
        for( int m = 0; m < 2; m++ )
        {
            const uint32_t h02 = __vimin_s16x2_relu( __byte_perm( hv[4*m], hv[4*m + 2], 0x5410 ), 0x1fff1fffu );
            const uint32_t h13 = __vimin_s16x2_relu( __byte_perm( hv[4*m + 1], hv[4*m + 3], 0x5410 ), 0x1fff1fffu );
            ho[m] = byte_lanes( h02 >> 5, h13 << 3 );
            const uint32_t c02 = __vimin_s16x2_relu( __byte_perm( cv[4*m], cv[4*m + 2], 0x6521 ), 0x03ff03ffu );
            const uint32_t c13 = __vimin_s16x2_relu( __byte_perm( cv[4*m + 1], cv[4*m + 3], 0x6521 ), 0x03ff03ffu );
            co[m] = byte_lanes( c02 >> 2, c13 << 6 );
        }
        if( second )
        {
            *(uint2 *)ph = make_uint2( ho[0], ho[1] );
            *(uint2 *)pc = make_uint2( co[0], co[1] );
        }
        else
        {
            *(uint32_t *)ph = ho[0];
            *(uint32_t *)pc = co[0];
        }
        if( border_tile )
        {
            const bool row_out = oy < 0 || oy >= height;
            uint8_t *ps = dsrc_border + (intptr_t)oy * stride + ox;
            if( row_out || col_out ) *(uint32_t *)ps = w1;
            if( second && ( row_out || col_out2 ) ) *(uint32_t *)( ps + 4 ) = w2;
        }
    }
}

} // namespace

// the same launch on a stream of the caller's choice (the lookahead's upload stream)
static int lowres_launch( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, intptr_t luma_pitch, int n_pictures,
                          int width, int height, uint8_t *const d_lowres[4], intptr_t lowres_stride, intptr_t lowres_pitch )
{
    if( width < 2 || height < 2 || n_pictures < 1 ) return x264cu_fail( ctx, "frame_init_lowres: bad size %dx%d", width, height );
    const int wl = ( ( width + 15 ) >> 4 ) * 8, ll = ( ( height + 15 ) >> 4 ) * 8;
    if( ( lowres_stride & 3 ) || lowres_stride < wl + 2*X264CU_PAD )
        return x264cu_fail( ctx, "frame_init_lowres: lowres stride %ld too small / unaligned", (long)lowres_stride );
    uintptr_t dst_bits = (uintptr_t)lowres_stride | (uintptr_t)lowres_pitch;
    for( int i = 0; i < 4; i++ )
    {
        if( (uintptr_t)d_lowres[i] & 3 ) return x264cu_fail( ctx, "frame_init_lowres: plane origins must be 4-byte aligned" );
        dst_bits |= (uintptr_t)d_lowres[i];
    }
    // the vector fast path needs 8-byte aligned source rows; otherwise every pixel takes the clamped path
    const int aligned = !( (uintptr_t)d_luma & 7 ) && !( luma_stride & 7 ) && !( luma_pitch & 7 );
    // interior columns in groups of 8 output pixels whose 17 source bytes lie inside the picture: the wide kernel
    int groups = 0;
    if( !( ( (uintptr_t)d_luma | (uintptr_t)luma_stride | (uintptr_t)luma_pitch ) & 15 ) && !( dst_bits & 7 ) )
        groups = min( wl / 8, ( width - 17 ) / 16 + 1 );
    if( groups > 0 && 16 * ( groups - 1 ) + 20 > luma_stride ) groups--;                 // the 4 bytes past column 16 must be readable
    groups = max( groups, 0 );
    const int groups_x = ( wl + 2*X264CU_PAD ) / 4;
    const long border_threads = 2L * X264CU_PAD * groups_x + (long)ll * ( groups_x - 2 * groups );
    const int blocks_x = groups ? ( groups + 127 ) / 128 : 8, wide_rows = groups ? ( ll + 1 ) / 2 : 0;
    const int border_rows = (int)( ( border_threads + 128L * blocks_x - 1 ) / ( 128L * blocks_x ) );
    dim3 grid( blocks_x, wide_rows + border_rows, n_pictures );
    lowres_fused_kernel<<<grid, 128, 0, stream>>>( d_luma, luma_stride, luma_pitch, width, height, d_lowres[0], d_lowres[1], d_lowres[2],
                                                   d_lowres[3], lowres_stride, lowres_pitch, groups, wl, ll, aligned, wide_rows );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

int x264cu_frame_init_lowres_on( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                                 uint8_t *const d_lowres[4], intptr_t lowres_stride )
{
    if( !ctx ) return -1;
    return lowres_launch( ctx, stream, d_luma, luma_stride, 0, 1, width, height, d_lowres, lowres_stride, 0 );
}

extern "C" {

int x264cu_frame_init_lowres( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                              uint8_t *const d_lowres[4], intptr_t lowres_stride )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    return x264cu_frame_init_lowres_on( ctx, ctx->stream, d_luma, luma_stride, width, height, d_lowres, lowres_stride );
}

static int hpel_launch( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, intptr_t pitch, int n_planes, int width, int height,
                        uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src )
{
    if( width < 1 || height < 1 || n_planes < 1 ) return x264cu_fail( ctx, "hpel_filter: bad size" );
    const uintptr_t all = (uintptr_t)d_src | (uintptr_t)d_h | (uintptr_t)d_v | (uintptr_t)d_c | (uintptr_t)stride | (uintptr_t)pitch;
    if( !( all & 15 ) && !( width & 3 ) && width >= 8 )
    {
        // the source pictures as a 3D tensor (column, row, picture) for the interior tiles' TMA loads
        CUtensorMap tm;
        cuuint64_t dims[3] = { (cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)n_planes };
        cuuint64_t strides[2] = { (cuuint64_t)stride, (cuuint64_t)( n_planes > 1 ? pitch : stride * height ) };
        cuuint32_t box[3] = { PT_PITCH, PT_ROWS, 1 }, estr[3] = { 1, 1, 1 };
        CUresult r = ctx->encode_tiled( &tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_src, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
        if( r != CUDA_SUCCESS ) return x264cu_fail( ctx, "hpel_filter: cuTensorMapEncodeTiled failed (%d)", (int)r );
        dim3 grid( ( width + 2*X264CU_PAD + PT_W - 1 ) / PT_W, ( height + 2*X264CU_PAD + PT_H - 1 ) / PT_H, n_planes );
        hpel_packed_kernel<<<grid, PT_THREADS, 0, ctx->stream>>>( tm, d_src, stride, pitch, width, height, d_h, d_v, d_c, expand_src ? d_src : nullptr, 0xf000f000u );
        CU_LAUNCH_CHECK( ctx );
        return 0;
    }
    for( int p = 0; p < n_planes; p++ )
    {
        const intptr_t o = (intptr_t)p * pitch;
        const bool words = !( all & 3 ) && !( width & 3 );
        if( words )
        {
            dim3 grid( ( width + 2*X264CU_PAD + FT_W - 1 ) / FT_W, ( height + 2*X264CU_PAD + FT_H - 1 ) / FT_H );
            hpel_words_kernel<<<grid, 256, 0, ctx->stream>>>( d_src + o, stride, width, height, d_h + o, d_v + o, d_c + o, expand_src ? d_src + o : nullptr );
        }
        else
        {
            dim3 grid( ( width + 2*X264CU_PAD + HT_W - 1 ) / HT_W, ( height + 2*X264CU_PAD + HT_H - 1 ) / HT_H );
            hpel_kernel<<<grid, 256, 0, ctx->stream>>>( d_src + o, stride, width, height, d_h + o, d_v + o, d_c + o, expand_src ? d_src + o : nullptr );
        }
        CU_LAUNCH_CHECK( ctx );
    }
    return 0;
}

int x264cu_frame_init_lowres_batch( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, intptr_t luma_pitch, int n_pictures,
                                    int width, int height, uint8_t *const d_lowres[4], intptr_t lowres_stride, intptr_t lowres_pitch )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    return lowres_launch( ctx, ctx->stream, d_luma, luma_stride, luma_pitch, n_pictures, width, height, d_lowres, lowres_stride, lowres_pitch );
}

int x264cu_hpel_filter( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, int width, int height,
                        uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    return hpel_launch( ctx, d_src, stride, 0, 1, width, height, d_h, d_v, d_c, expand_src );
}

int x264cu_hpel_filter_batch( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, intptr_t plane_pitch, int n_planes, int width, int height,
                              uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    return hpel_launch( ctx, d_src, stride, plane_pitch, n_planes, width, height, d_h, d_v, d_c, expand_src );
}

} // extern "C"
