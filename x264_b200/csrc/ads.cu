// Successive elimination: the integral image the reference builds per reference frame (integral_init4h / 8h / 4v / 8v
// driven by x264_frame_filter, common/mc.c:424-456, :748-783) and pixf.ads[] (x264_pixel_ads1 / 2 / 4, common/pixel.c:759-803), the
// prefilter of the exhaustive searches (me.c:618-771).
//   * the reference's four passes leave, at every position of the padded plane, the sum of the 8x8 (upper plane) and 4x4 (lower
//     plane, sub-8x8 partitions only) pixel box whose top-left corner is that position, as u16 -- here ONE kernel per plane size:
//     running column sums down a strip (sliding window, one subtraction and one addition per new row), horizontal box by prefix
//     difference through shuffles; HBM-bound: 1 byte in, 2 (4) bytes out per position
//   * ads: one warp walks one row of candidate positions (what one call of pixf.ads does), 32 at a time, and writes the passing
//     indices in ascending order (ballot compaction) -- the list the reference's loop leaves in mvs[]
#include "ctx.h"

namespace {

// box sums of size N x N at every (x, y) of [-PAD, w+PAD-N] x [-PAD, h+PAD-N]; plane = pixel (0,0), padded; out same geometry (elements)
template <int N>
__global__ void __launch_bounds__( 128 )
box_sum_kernel( const uint8_t *__restrict__ plane, intptr_t stride, int width, int height, uint16_t *__restrict__ out, int rows_per_block )
{
    const int lane = threadIdx.x & 31;
    // a warp owns 32 - (N-1) output columns: lanes carry N-1 extra columns to the right for the horizontal box
    constexpr int OUT = 32 - ( N - 1 );
    const int warp = ( blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 ) );
    const int x_first = -X264CU_PAD, x_last = width + X264CU_PAD - N;           // output columns
    const int y_first = -X264CU_PAD, y_last = height + X264CU_PAD - N;
    const int ox = x_first + warp * OUT + lane;                                 // this lane's column (output column if lane < OUT)
    if( x_first + warp * OUT > x_last ) return;
    const int cx = min( ox, width + X264CU_PAD - 1 );                           // clamp the halo lanes of the last strip
    const int y0 = y_first + blockIdx.y * rows_per_block;
    if( y0 > y_last ) return;
    const int y1 = min( y0 + rows_per_block - 1, y_last );
    const uint8_t *col = plane + cx;
    int v = 0;                                                                  // column sum of rows y .. y+N-1
#pragma unroll
    for( int k = 0; k < N; k++ ) v += col[(intptr_t)( y0 + k ) * stride];
    for( int y = y0; ; y++ )
    {
        // horizontal box over lanes lane .. lane+N-1
        int s = v;
#pragma unroll
        for( int k = 1; k < N; k++ ) s += __shfl_down_sync( 0xffffffffu, v, k );
        if( lane < OUT && ox <= x_last )
            out[(intptr_t)y * stride + ox] = (uint16_t)s;
        if( y == y1 ) break;
        v += col[(intptr_t)( y + N ) * stride] - col[(intptr_t)y * stride];
    }
}

struct AdsJob                                   // == x264cu_ads_job_t
{
    int32_t enc_dc[4];
    uint32_t sums_off;                          // element offset of the row's first position in the sums plane
    int32_t delta;                              // elements between the sub-blocks (8 / 4 or that times the stride)
    uint32_t cost_off;                          // offset of cost_mvx[0] in the cost table
    int32_t width, thresh;
    uint32_t out_off;                           // where this row's list starts in d_mvs
};
static_assert( sizeof( AdsJob ) == sizeof( x264cu_ads_job_t ), "ABI struct" );

template <int K>                                // ads1 / ads2 / ads4
__global__ void __launch_bounds__( 128 )
ads_kernel( const uint16_t *__restrict__ sums, const uint16_t *__restrict__ cost, const AdsJob *__restrict__ jobs, int n,
            int32_t *__restrict__ counts, int16_t *__restrict__ mvs )
{
    const int lane = threadIdx.x & 31;
    const int w = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    if( w >= n ) return;
    const AdsJob j = jobs[w];
    const uint16_t *s = sums + j.sums_off, *c = cost + j.cost_off;
    int16_t *out = mvs + j.out_off;
    int nmv = 0;
    for( int base = 0; base < j.width; base += 32 )
    {
        const int i = base + lane;
        bool pass = false;
        if( i < j.width )
        {
            int ads = abs( j.enc_dc[0] - (int)s[i] ) + (int)c[i];
            if( K == 2 ) ads += abs( j.enc_dc[1] - (int)s[i + j.delta] );
            if( K == 4 ) ads += abs( j.enc_dc[1] - (int)s[i + 8] ) + abs( j.enc_dc[2] - (int)s[i + j.delta] ) + abs( j.enc_dc[3] - (int)s[i + j.delta + 8] );
            pass = ads < j.thresh;
        }
        const uint32_t m = __ballot_sync( 0xffffffffu, pass );
        if( pass ) out[nmv + __popc( m & ( ( 1u << lane ) - 1 ) )] = (int16_t)i;
        nmv += __popc( m );
    }
    if( lane == 0 ) counts[w] = nmv;
}

}

extern "C" {

int x264cu_integral_init( x264cu_ctx_t *ctx, const uint8_t *d_plane, intptr_t stride, int width, int height,
                          uint16_t *d_sum8, uint16_t *d_sum4 )
{
    X264CU_ENTER( ctx );
    if( !ctx || !d_plane || !d_sum8 ) return -1;
    if( width < 8 || height < 8 || stride < width + 2 * X264CU_PAD ) return x264cu_fail( ctx, "integral_init: bad geometry" );
    const int rows = 64;
    {
        const int cols = width + 2 * X264CU_PAD - 7, strips = ( cols + 24 ) / 25, lines = height + 2 * X264CU_PAD - 7;
        box_sum_kernel<8><<<dim3( ( strips + 3 ) / 4, ( lines + rows - 1 ) / rows ), 128, 0, ctx->stream>>>( d_plane, stride, width, height, d_sum8, rows );
        CU_LAUNCH_CHECK( ctx );
    }
    if( d_sum4 )
    {
        const int cols = width + 2 * X264CU_PAD - 3, strips = ( cols + 28 ) / 29, lines = height + 2 * X264CU_PAD - 3;
        box_sum_kernel<4><<<dim3( ( strips + 3 ) / 4, ( lines + rows - 1 ) / rows ), 128, 0, ctx->stream>>>( d_plane, stride, width, height, d_sum4, rows );
        CU_LAUNCH_CHECK( ctx );
    }
    return 0;
}

int x264cu_pixel_ads_batch( x264cu_ctx_t *ctx, int i_pixel, const uint16_t *d_sums, const uint16_t *d_cost_mvx,
                            const x264cu_ads_job_t *d_jobs, int n, int32_t *d_counts, int16_t *d_mvs )
{
    X264CU_ENTER( ctx );
    if( !ctx || !d_sums || !d_cost_mvx || !d_jobs || !d_counts || !d_mvs ) return -1;
    if( (unsigned)i_pixel >= X264CU_PIXEL_4x16 ) return x264cu_fail( ctx, "pixel_ads_batch: bad block size index %d", i_pixel );
    if( n <= 0 ) return 0;
    const int blocks = ( n + 3 ) / 4;
    // pixf.ads[] (common/pixel.c:860-862, :1660-1663): 16x16 -> ads4; 16x8, 8x16, 8x4, 4x8 -> ads2; 8x8, 4x4 -> ads1
    if( i_pixel == X264CU_PIXEL_16x16 )
        ads_kernel<4><<<blocks, 128, 0, ctx->stream>>>( d_sums, d_cost_mvx, (const AdsJob *)d_jobs, n, d_counts, d_mvs );
    else if( i_pixel == X264CU_PIXEL_8x8 || i_pixel == X264CU_PIXEL_4x4 )
        ads_kernel<1><<<blocks, 128, 0, ctx->stream>>>( d_sums, d_cost_mvx, (const AdsJob *)d_jobs, n, d_counts, d_mvs );
    else
        ads_kernel<2><<<blocks, 128, 0, ctx->stream>>>( d_sums, d_cost_mvx, (const AdsJob *)d_jobs, n, d_counts, d_mvs );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

} // extern "C"
