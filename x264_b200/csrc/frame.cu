// Frame preparation kernels: the streaming (HBM-bound) entries of x264_mc_functions_t.
//   lowres_kernel  : x264_frame_init_lowres + frame_init_lowres_core + expand_border_lowres
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
__global__ void __launch_bounds__( 256 )
lowres_kernel( const uint8_t *__restrict__ src, intptr_t src_stride, int width, int height,
               uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc, intptr_t dst_stride, int wl, int ll, int fast_ok )
{
    // output domain incl. border: x in [-PAD, wl+PAD) in groups of 4, y in [-PAD, ll+PAD)
    const int groups_x = ( wl + 2*X264CU_PAD ) / 4;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y - X264CU_PAD;
    if( gx >= groups_x ) return;
    const int ox = gx * 4 - X264CU_PAD;
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

} // namespace

// the same launch on a stream of the caller's choice (the lookahead's upload stream)
int x264cu_frame_init_lowres_on( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                                 uint8_t *const d_lowres[4], intptr_t lowres_stride )
{
    if( !ctx ) return -1;
    if( width < 2 || height < 2 ) return x264cu_fail( ctx, "frame_init_lowres: bad size %dx%d", width, height );
    const int wl = ( ( width + 15 ) >> 4 ) * 8, ll = ( ( height + 15 ) >> 4 ) * 8;
    if( ( lowres_stride & 3 ) || lowres_stride < wl + 2*X264CU_PAD )
        return x264cu_fail( ctx, "frame_init_lowres: lowres stride %ld too small / unaligned", (long)lowres_stride );
    for( int i = 0; i < 4; i++ )
        if( (uintptr_t)d_lowres[i] & 3 ) return x264cu_fail( ctx, "frame_init_lowres: plane origins must be 4-byte aligned" );
    // the vector fast path needs 8-byte aligned source rows; otherwise every pixel takes the clamped path
    const int aligned = !( (uintptr_t)d_luma & 7 ) && !( luma_stride & 7 );
    dim3 block( 256 ), grid( ( ( wl + 2*X264CU_PAD ) / 4 + 255 ) / 256, ll + 2*X264CU_PAD );
    lowres_kernel<<<grid, block, 0, stream>>>( d_luma, luma_stride, width, height, d_lowres[0], d_lowres[1],
                                               d_lowres[2], d_lowres[3], lowres_stride, wl, ll, aligned );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

extern "C" {

int x264cu_frame_init_lowres( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                              uint8_t *const d_lowres[4], intptr_t lowres_stride )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    return x264cu_frame_init_lowres_on( ctx, ctx->stream, d_luma, luma_stride, width, height, d_lowres, lowres_stride );
}

int x264cu_hpel_filter( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, int width, int height,
                        uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( width < 1 || height < 1 ) return x264cu_fail( ctx, "hpel_filter: bad size" );
    const bool words = !( ( (uintptr_t)d_src | (uintptr_t)d_h | (uintptr_t)d_v | (uintptr_t)d_c | (uintptr_t)stride ) & 3 ) && !( width & 3 );
    if( words )
    {
        dim3 grid( ( width + 2*X264CU_PAD + FT_W - 1 ) / FT_W, ( height + 2*X264CU_PAD + FT_H - 1 ) / FT_H );
        hpel_words_kernel<<<grid, 256, 0, ctx->stream>>>( d_src, stride, width, height, d_h, d_v, d_c, expand_src ? d_src : nullptr );
        CU_LAUNCH_CHECK( ctx );
        return 0;
    }
    dim3 grid( ( width + 2*X264CU_PAD + HT_W - 1 ) / HT_W, ( height + 2*X264CU_PAD + HT_H - 1 ) / HT_H );
    hpel_kernel<<<grid, 256, 0, ctx->stream>>>( d_src, stride, width, height, d_h, d_v, d_c, expand_src ? d_src : nullptr );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

} // extern "C"
