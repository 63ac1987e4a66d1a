// Stand-alone batched twins of the x264_mc_functions_t block entries (common/mc.h:267-340) and of the whole-plane SSD:
//   mc_luma / get_ref  common/mc.c:198-249   quarter-pel sample = one half-pel plane, or the rounded mean of two (+ weight)
//   avg[]              common/mc.c:49-111    bi-prediction mean / weighted mean of two blocks
//   weight             common/mc.c:117-160   + x264_weight_scale_plane, common/frame.c:825-841
//   x264_pixel_ssd_wxh common/pixel.c:112-151
// Inside the lookahead and the motion search these are fused into the cost kernels (lookahead_dev.cuh, me_dev.cuh: the
// interpolated block never exists in memory); these entries materialise it, for callers that want the prediction itself.
// All of them are streaming byte kernels: a thread produces 4 horizontally adjacent pixels, a block's (w/4)*h groups are
// laid out so that the stores of a warp are contiguous.
#include "ctx.h"
#include "lookahead_dev.cuh"

using namespace x264cu;

namespace {

struct McParams
{
    const uint8_t *src[4];           // F,H,V,C plane bases (the offsets of the jobs are relative to these)
    intptr_t stride;
    int w4, h;                       // block width / 4, block height
    LaWeight wt;
};

__global__ void __launch_bounds__( 256 )
mc_luma_kernel( McParams p, const x264cu_mc_job_t *__restrict__ jobs, int n, uint32_t *__restrict__ dst )
{
    const int groups = p.w4 * p.h;                           // 4-pixel groups per block
    const long long total = (long long)n * groups;
    for( long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x )
    {
        const int job = (int)( t / groups ), g = (int)( t - (long long)job * groups );
        const int row = g / p.w4, col = ( g - row * p.w4 ) * 4;
        const x264cu_mc_job_t j = jobs[job];
        const int mvx = j.mvx, mvy = j.mvy;
        // hpel_ref0 / hpel_ref1 (common/tables.c:183-184), 2 bits per entry
        const uint32_t R0 = 0x54FE5454u, R1 = 0xBABABA10u;
        const int idx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
        const intptr_t off = (intptr_t)j.src_off + (intptr_t)( ( mvy >> 2 ) + row ) * p.stride + ( mvx >> 2 ) + col;
        const int k0 = ( R0 >> ( 2 * idx ) ) & 3, k1 = ( R1 >> ( 2 * idx ) ) & 3;
        const uint8_t *s1 = ( k0 == 0 ? p.src[0] : k0 == 1 ? p.src[1] : k0 == 2 ? p.src[2] : p.src[3] ) + off + ( ( mvy & 3 ) == 3 ? p.stride : 0 );
        uint32_t v = ldg4u( s1 );
        if( idx & 5 )
        {
            const uint8_t *s2 = ( k1 == 0 ? p.src[0] : k1 == 1 ? p.src[1] : k1 == 2 ? p.src[2] : p.src[3] ) + off + ( ( mvx & 3 ) == 3 ? 1 : 0 );
            v = __vavgu4( v, ldg4u( s2 ) );                  // pixel_avg: (a+b+1)>>1
        }
        if( p.wt.enabled ) v = weight4( v, p.wt );
        dst[t] = v;
    }
}

__device__ __forceinline__ uint32_t avg4_weighted( uint32_t a, uint32_t b, int weight )       // pixel_avg_weight_wxh, mc.c:63-75
{
    uint32_t out = 0;
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int p = ( a >> ( 8 * i ) ) & 255, q = ( b >> ( 8 * i ) ) & 255;
        int v = ( p * weight + q * ( 64 - weight ) + 32 ) >> 6;
        v = min( max( v, 0 ), 255 );
        out |= (uint32_t)v << ( 8 * i );
    }
    return out;
}

__global__ void __launch_bounds__( 256 )
pixel_avg_kernel( const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t words, int weight, uint32_t *__restrict__ dst )
{
    for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x )
        dst[i] = weight == 32 ? __vavgu4( a[i], b[i] ) : avg4_weighted( a[i], b[i], weight );
}

__global__ void __launch_bounds__( 256 )
weight_rows_kernel( const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, intptr_t stride, int width, int height, LaWeight w )
{
    const int w4 = ( width + 3 ) >> 2;
    const long long total = (long long)w4 * height;
    for( long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x )
    {
        const int y = (int)( t / w4 ), x = (int)( t - (long long)y * w4 ) * 4;
        const uint8_t *s = src + (intptr_t)y * stride + x;
        uint8_t *d = dst + (intptr_t)y * stride + x;
        if( x + 4 <= width && !( ( (uintptr_t)s | (uintptr_t)d ) & 3 ) )
            *(uint32_t *)d = weight4( *(const uint32_t *)s, w );
        else
            for( int k = 0; k < 4 && x + k < width; k++ )
                d[k] = (uint8_t)( weight4( s[k], w ) & 255 );
    }
}

__global__ void __launch_bounds__( 256 )
ssd_wxh_kernel( const uint8_t *__restrict__ a, intptr_t sa, const uint8_t *__restrict__ b, intptr_t sb, int width, int height,
                unsigned long long *out )
{
    const int w4 = ( width + 3 ) >> 2;
    const long long total = (long long)w4 * height;
    unsigned long long acc = 0;
    for( long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x )
    {
        const int y = (int)( t / w4 ), x = (int)( t - (long long)y * w4 ) * 4;
        const uint8_t *pa = a + (intptr_t)y * sa + x, *pb = b + (intptr_t)y * sb + x;
        if( x + 4 <= width )
        {
            // arbitrary caller planes (no readable border promised): aligned word or four byte loads, never past x+3
            auto ld4 = []( const uint8_t *q ) {
                return ( (uintptr_t)q & 3 ) ? (uint32_t)q[0] | ( (uint32_t)q[1] << 8 ) | ( (uint32_t)q[2] << 16 ) | ( (uint32_t)q[3] << 24 )
                                            : __ldg( (const uint32_t *)q );
            };
            const uint32_t va = ld4( pa ), vb = ld4( pb );
            // sum (a-b)^2 over 4 bytes = a.a - 2 a.b + b.b, three DP4A on the packed bytes
            acc += __dp4a( va, va, 0u ) + __dp4a( vb, vb, 0u ) - 2u * __dp4a( va, vb, 0u );
        }
        else
            for( int k = 0; x + k < width; k++ ) { int d = (int)pa[k] - (int)pb[k]; acc += (unsigned)( d * d ); }
    }
    for( int m = 16; m; m >>= 1 ) acc += __shfl_xor_sync( 0xffffffffu, acc, m );
    if( ( threadIdx.x & 31 ) == 0 && acc ) atomicAdd( out, acc );
}

int grid_for( x264cu_ctx *ctx, long long work_items )
{
    long long blocks = ( work_items + 255 ) / 256;
    const long long cap = (long long)ctx->sm_count * 8;      // a whole number of waves: 8 resident 256-thread blocks per SM
    return (int)( blocks < 1 ? 1 : blocks > cap ? cap : blocks );
}

} // namespace

extern "C" {

int x264cu_mc_luma_batch( x264cu_ctx_t *ctx, const uint8_t *const d_src[4], intptr_t src_stride, int i_pixel,
                          const x264cu_mc_job_t *d_jobs, int n, const int weight[4], uint8_t *d_dst )
{
    X264CU_ENTER( ctx );
    static const int W[8] = { 16, 16, 8, 8, 8, 4, 4, 4 }, H[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };
    if( !ctx ) return -1;
    if( i_pixel < 0 || i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "mc_luma_batch: bad i_pixel %d", i_pixel );
    if( n < 0 || !d_src || ( n && ( !d_jobs || !d_dst ) ) ) return x264cu_fail( ctx, "mc_luma_batch: bad arguments" );
    if( (uintptr_t)d_dst & 3 ) return x264cu_fail( ctx, "mc_luma_batch: d_dst must be 4-byte aligned" );
    if( !n ) return 0;
    McParams p;
    for( int i = 0; i < 4; i++ ) p.src[i] = d_src[i];
    p.stride = src_stride;
    p.w4 = W[i_pixel] / 4; p.h = H[i_pixel];
    p.wt.enabled = weight ? weight[0] : 0;
    p.wt.scale = weight ? weight[1] : 1; p.wt.denom = weight ? weight[2] : 0; p.wt.offset = weight ? weight[3] : 0;
    mc_luma_kernel<<<grid_for( ctx, (long long)n * p.w4 * p.h ), 256, 0, ctx->stream>>>( p, d_jobs, n, (uint32_t *)d_dst );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

int x264cu_pixel_avg_batch( x264cu_ctx_t *ctx, int i_pixel, const uint8_t *d_a, const uint8_t *d_b, int n, int weight, uint8_t *d_dst )
{
    X264CU_ENTER( ctx );
    static const int W[8] = { 16, 16, 8, 8, 8, 4, 4, 4 }, H[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };
    if( !ctx ) return -1;
    if( i_pixel < 0 || i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "pixel_avg_batch: bad i_pixel %d", i_pixel );
    if( n < 0 || ( n && ( !d_a || !d_b || !d_dst ) ) ) return x264cu_fail( ctx, "pixel_avg_batch: bad arguments" );
    if( ( (uintptr_t)d_a | (uintptr_t)d_b | (uintptr_t)d_dst ) & 3 ) return x264cu_fail( ctx, "pixel_avg_batch: blocks must be 4-byte aligned" );
    if( !n ) return 0;
    const size_t words = (size_t)n * W[i_pixel] * H[i_pixel] / 4;
    pixel_avg_kernel<<<grid_for( ctx, (long long)words ), 256, 0, ctx->stream>>>( (const uint32_t *)d_a, (const uint32_t *)d_b, words, weight,
                                                                                  (uint32_t *)d_dst );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

int x264cu_weight_scale_plane( x264cu_ctx_t *ctx, const uint8_t *d_src, uint8_t *d_dst, intptr_t stride, int width, int height,
                               const int weight[4] )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( !d_src || !d_dst || !weight || width < 1 || height < 1 ) return x264cu_fail( ctx, "weight_scale_plane: bad arguments" );
    LaWeight w = { 1, weight[1], weight[2], weight[3] };
    weight_rows_kernel<<<grid_for( ctx, (long long)( ( width + 3 ) / 4 ) * height ), 256, 0, ctx->stream>>>( d_src, d_dst, stride, width, height, w );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

int x264cu_pixel_ssd_wxh( x264cu_ctx_t *ctx, const uint8_t *d_pix1, intptr_t stride1, const uint8_t *d_pix2, intptr_t stride2,
                          int width, int height, uint64_t *h_ssd )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( !d_pix1 || !d_pix2 || !h_ssd || width < 0 || height < 0 ) return x264cu_fail( ctx, "pixel_ssd_wxh: bad arguments" );
    *h_ssd = 0;
    if( !width || !height ) return 0;
    unsigned long long *d_acc = (unsigned long long *)x264cu_scratch( ctx, 6, 8 );
    if( !d_acc ) return -1;
    CU_CHECK( ctx, cudaMemsetAsync( d_acc, 0, 8, ctx->stream ) );
    ssd_wxh_kernel<<<grid_for( ctx, (long long)( ( width + 3 ) / 4 ) * height ), 256, 0, ctx->stream>>>( d_pix1, stride1, d_pix2, stride2,
                                                                                                    width, height, d_acc );
    CU_LAUNCH_CHECK( ctx );
    unsigned long long v = 0;
    CU_CHECK( ctx, cudaMemcpyAsync( &v, d_acc, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    *h_ssd = v;
    return 0;
}

} // extern "C"
