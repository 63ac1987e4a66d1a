// x264_adaptive_quant_frame (encoder/ratecontrol.c:225-420), aq-mode 0 - 3: the producer of the lookahead's per-macroblock inputs
// fenc->i_inv_qscale_factor / f_qp_offset_aq and of the frame statistics the lookahead weight analysis reads.
//   ac_energy_mb (:261-303) = pixel_var of the 16x16 luma block and of the two 8x8 chroma blocks (common/pixel.c:183-203),
//   f_qp_offset_aq = strength * (x264_log2( max(energy,1) ) - 14.427) (:397), i_inv_qscale_factor = x264_exp2fix8 (base.h:218-224).
// One warp per macroblock: every lane sums 8 luma pixels (dp4a with ones / with itself), lanes 0-15 / 16-31 four pixels of Cb / Cr;
// xor-shuffle reductions; lane 0 does the float step operation by operation in the order the reference's -ffast-math build uses
// (see DESIGN.md section 2).  Pure streaming: 1.5 bytes read per pixel, 6 bytes written per macroblock.
// The auto-variance modes (2, 3; :352-392) need the frame means of qp = (energy+1)^(1/8) and of qp^2, which the reference sums in
// single precision in raster order: aq_means_kernel does exactly that, one thread, sequentially (65 us at 4K, once per picture,
// beside the lookahead) -- a parallel reduction would round differently -- and aq_final_kernel applies them per macroblock.
#include "ctx.h"
#include <math.h>

namespace {

__constant__ float c_aq_log2_lut[128];       // x264_log2_lut (common/tables.c:66-85)
__constant__ uint8_t c_aq_exp2_lut[64];      // x264_exp2_lut (common/tables.c:58-64)

__global__ void __launch_bounds__( 256 )
aq_kernel( const uint8_t *__restrict__ luma, intptr_t stride, const uint8_t *__restrict__ cb, const uint8_t *__restrict__ cr,
           intptr_t cstride, int width, int height, int mb_w, int mb_count, int active, int aq_mode, float strength,
           float *__restrict__ qp_offset_aq, uint16_t *__restrict__ inv_qscale, float *__restrict__ q4_out,
           unsigned long long *__restrict__ stats )
{
    const int lane = threadIdx.x & 31;
    const int mb = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const bool valid = mb < mb_count;
    const int mbc = valid ? mb : mb_count - 1;
    const int mb_y = mbc / mb_w, mb_x = mbc - mb_y * mb_w;
    const int cw = ( width + 1 ) >> 1, ch = ( height + 1 ) >> 1;
    // luma: lane = (row pair, 8-pixel half): rows lane>>1 ... 16 rows x 2 halves = 32 lanes, 8 pixels each
    unsigned sum[3] = { 0, 0, 0 }, sqr[3] = { 0, 0, 0 };
    {
        const int y = min( 16 * mb_y + ( lane >> 1 ), height - 1 ), x0 = 16 * mb_x + ( lane & 1 ) * 8;
        const uint8_t *row = luma + (intptr_t)y * stride;
        if( x0 + 8 <= width && !( ( (uintptr_t)row + x0 ) & 7 ) )
        {
            const uint2 v = __ldg( (const uint2 *)( row + x0 ) );
            sum[0] = __dp4a( v.x, 0x01010101u, __dp4a( v.y, 0x01010101u, 0u ) );
            sqr[0] = __dp4a( v.x, v.x, __dp4a( v.y, v.y, 0u ) );
        }
        else      // picture edge: replicated to the macroblock grid (x264_frame_expand_border_mod16, frame.c:640-665)
            for( int k = 0; k < 8; k++ ) { unsigned p = row[min( x0 + k, width - 1 )]; sum[0] += p; sqr[0] += p * p; }
    }
    {   // chroma: lanes 0-15 Cb, 16-31 Cr; lane = (row, 4-pixel half) of the 8x8 block
        const int l = lane & 15, pl = 1 + ( lane >> 4 );
        const uint8_t *base = pl == 1 ? cb : cr;
        const int y = min( 8 * mb_y + ( l >> 1 ), ch - 1 ), x0 = 8 * mb_x + ( l & 1 ) * 4;
        const uint8_t *row = base + (intptr_t)y * cstride;
        unsigned s = 0, q = 0;
        if( x0 + 4 <= cw && !( ( (uintptr_t)row + x0 ) & 3 ) )
        {
            const unsigned v = __ldg( (const unsigned *)( row + x0 ) );
            s = __dp4a( v, 0x01010101u, 0u ); q = __dp4a( v, v, 0u );
        }
        else
            for( int k = 0; k < 4; k++ ) { unsigned p = row[min( x0 + k, cw - 1 )]; s += p; q += p * p; }
        sum[pl] = s; sqr[pl] = q;
    }
#pragma unroll
    for( int m = 16; m; m >>= 1 )
#pragma unroll
        for( int i = 0; i < 3; i++ )
        {
            sum[i] += __shfl_xor_sync( 0xffffffffu, sum[i], m );
            sqr[i] += __shfl_xor_sync( 0xffffffffu, sqr[i], m );
        }
    if( lane == 0 && valid )
    {
        // ac_energy_var (ratecontrol.c:225-235): ssd - (sum*sum >> shift), 32-bit; shift 8 for 256 luma pixels, 6 for 64 chroma pixels
        unsigned energy = sqr[0] - (unsigned)( (unsigned long long)sum[0] * sum[0] >> 8 );
        energy += sqr[1] - (unsigned)( (unsigned long long)sum[1] * sum[1] >> 6 );
        energy += sqr[2] - (unsigned)( (unsigned long long)sum[2] * sum[2] >> 6 );
        float qp_adj = 0.f;
        unsigned q = 256;
        if( active && aq_mode >= 2 )
        {   // first pass of the auto-variance modes: powf( energy + 1, 0.125f ) as the reference's -ffast-math build computes it,
            // three correctly rounded square roots; the second one doubles as qp * qp in the frame mean
            const float q4 = __fsqrt_rn( __fsqrt_rn( __fadd_rn( __uint2float_rn( energy ), 1.f ) ) );
            qp_adj = __fsqrt_rn( q4 );
            q4_out[mb] = q4;
        }
        else if( active )
        {
            const unsigned e = max( energy, 1u );
            const int lz = __clz( e );
            const float ipart = (float)( 31 - lz ), frac = c_aq_log2_lut[( e << lz >> 24 ) & 0x7f];
            qp_adj = __fmul_rn( strength, __fadd_rn( __fsub_rn( ipart, 14.427f ), frac ) );
            // x264_exp2fix8
            const int i = __float2int_rz( __fadd_rn( __fmul_rn( qp_adj, -64.f / 6.f ), 512.5f ) );
            q = i < 0 ? 0u : i > 1023 ? 0xffffu : ( ( (unsigned)c_aq_exp2_lut[i & 63] + 256u ) << ( i >> 6 ) ) >> 8;
        }
        qp_offset_aq[mb] = qp_adj;
        inv_qscale[mb] = (uint16_t)q;
        if( stats )
#pragma unroll
            for( int i = 0; i < 3; i++ )
            {
                atomicAdd( &stats[i], (unsigned long long)sum[i] );
                atomicAdd( &stats[3 + i], (unsigned long long)sqr[i] );
            }
    }
}

// avg_adj / avg_adj_pow2 of ratecontrol.c:352-372, summed in the reference's order; out = { avg_adj (final), strength }
__global__ void aq_means_kernel( const float *__restrict__ qp, const float *__restrict__ q4, int mb_count, float aq_strength, float *out )
{
    if( threadIdx.x || blockIdx.x ) return;
    float avg = 0.f, pow2 = 0.f;
    for( int i = 0; i < mb_count; i++ )
    {
        avg = __fadd_rn( avg, qp[i] );
        pow2 = __fadd_rn( pow2, q4[i] );
    }
    avg = __fdiv_rn( avg, (float)mb_count );
    pow2 = __fdiv_rn( pow2, (float)mb_count );
    out[1] = __fmul_rn( avg, aq_strength );
    out[0] = __fadd_rn( __fdiv_rn( __fmul_rn( 0.5f, __fsub_rn( 14.f, pow2 ) ), avg ), avg );
}

__global__ void __launch_bounds__( 256 )
aq_final_kernel( int mb_count, int aq_mode, float bias_strength, const float *__restrict__ means, float *__restrict__ qp_offset_aq,
                 uint16_t *__restrict__ inv_qscale )
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if( mb >= mb_count ) return;
    const float avg = means[0], strength = means[1], qp = qp_offset_aq[mb];
    float qp_adj = __fmul_rn( __fsub_rn( qp, avg ), strength );
    if( aq_mode == 3 )
        qp_adj = __fadd_rn( qp_adj, __fmul_rn( __fsub_rn( 1.f, __fdiv_rn( 14.f, __fmul_rn( qp, qp ) ) ), bias_strength ) );
    const int i = __float2int_rz( __fadd_rn( __fmul_rn( qp_adj, -64.f / 6.f ), 512.5f ) );
    qp_offset_aq[mb] = qp_adj;
    inv_qscale[mb] = (uint16_t)( i < 0 ? 0u : i > 1023 ? 0xffffu : ( ( (unsigned)c_aq_exp2_lut[i & 63] + 256u ) << ( i >> 6 ) ) >> 8 );
}

} // namespace

// the launches on a stream of the caller's choice (the lookahead's upload stream); d_q4: mb_count + 2 floats of scratch for
// the auto-variance modes; d_stats: 6 x u64 raw sums (zeroed here) or NULL
int x264cu_adaptive_quant_frame_on( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, const uint8_t *d_cb,
                                    const uint8_t *d_cr, intptr_t chroma_stride, int width, int height, int aq_mode, float aq_strength,
                                    float *d_qp_offset_aq, uint16_t *d_inv_qscale, float *d_q4, unsigned long long *d_stats )
{
    if( !ctx->aq_tables )
    {
        float l2[128];
        uint8_t e2[64];
        for( int i = 0; i < 128; i++ )
        {   // x264_log2_lut: log2( 1 + i/128 ) printed with five decimals
            char buf[32];
            snprintf( buf, sizeof( buf ), "%.5f", log2( 1.0 + i / 128.0 ) );
            l2[i] = strtof( buf, nullptr );
        }
        for( int i = 0; i < 64; i++ ) e2[i] = (uint8_t)( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) + 0.5 );      // x264_exp2_lut
        CU_CHECK( ctx, cudaMemcpyToSymbol( c_aq_log2_lut, l2, sizeof( l2 ) ) );
        CU_CHECK( ctx, cudaMemcpyToSymbol( c_aq_exp2_lut, e2, sizeof( e2 ) ) );
        ctx->aq_tables = true;
    }
    const int mb_w = ( width + 15 ) >> 4, mb_h = ( height + 15 ) >> 4, mb_count = mb_w * mb_h;
    if( d_stats ) CU_CHECK( ctx, cudaMemsetAsync( d_stats, 0, 48, stream ) );
    const int active = aq_mode != 0 && aq_strength != 0.f;
    const bool two_pass = active && aq_mode >= 2;
    if( two_pass && !d_q4 ) return x264cu_fail( ctx, "adaptive_quant_frame: no scratch for the auto-variance modes" );
    aq_kernel<<<( mb_count + 7 ) / 8, 256, 0, stream>>>( d_luma, luma_stride, d_cb, d_cr, chroma_stride, width, height, mb_w, mb_count,
                                                         active, aq_mode, aq_strength * 1.0397f, d_qp_offset_aq, d_inv_qscale, d_q4, d_stats );
    CU_LAUNCH_CHECK( ctx );
    if( two_pass )
    {
        aq_means_kernel<<<1, 32, 0, stream>>>( d_qp_offset_aq, d_q4, mb_count, aq_strength, d_q4 + mb_count );
        CU_LAUNCH_CHECK( ctx );
        aq_final_kernel<<<( mb_count + 255 ) / 256, 256, 0, stream>>>( mb_count, aq_mode, aq_strength, d_q4 + mb_count, d_qp_offset_aq, d_inv_qscale );
        CU_LAUNCH_CHECK( ctx );
    }
    return 0;
}

extern "C" int x264cu_adaptive_quant_frame( x264cu_ctx_t *ctx, const uint8_t *d_luma, intptr_t luma_stride, const uint8_t *d_cb,
                                            const uint8_t *d_cr, intptr_t chroma_stride, int width, int height, int aq_mode,
                                            float aq_strength, float *d_qp_offset_aq, uint16_t *d_inv_qscale, uint64_t *h_stats )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( !d_luma || !d_cb || !d_cr || !d_qp_offset_aq || !d_inv_qscale || width < 1 || height < 1 )
        return x264cu_fail( ctx, "adaptive_quant_frame: bad arguments" );
    if( aq_mode < 0 || aq_mode > 3 )
        return x264cu_fail( ctx, "adaptive_quant_frame: aq-mode %d out of range", aq_mode );
    const int mb_w = ( width + 15 ) >> 4, mb_h = ( height + 15 ) >> 4, mb_count = mb_w * mb_h;
    unsigned long long *d_stats = nullptr;
    if( h_stats )
    {
        d_stats = (unsigned long long *)x264cu_scratch( ctx, 7, 48 );
        if( !d_stats ) return -1;
    }
    float *d_q4 = nullptr;
    if( aq_mode >= 2 && aq_strength != 0.f )
    {   // scratch: (energy+1)^(1/4) per macroblock, then the two frame-level scalars
        d_q4 = (float *)x264cu_scratch( ctx, 8, (size_t)( mb_count + 2 ) * 4 );
        if( !d_q4 ) return -1;
    }
    if( x264cu_adaptive_quant_frame_on( ctx, ctx->stream, d_luma, luma_stride, d_cb, d_cr, chroma_stride, width, height, aq_mode, aq_strength,
                                        d_qp_offset_aq, d_inv_qscale, d_q4, d_stats ) )
        return -1;
    if( h_stats )
    {
        unsigned long long raw[6];
        CU_CHECK( ctx, cudaMemcpyAsync( raw, d_stats, 48, cudaMemcpyDeviceToHost, ctx->stream ) );
        CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
        for( int i = 0; i < 3; i++ )
        {   // "remove mean from SSD calculation", ratecontrol.c:405-414
            const unsigned long long w = (unsigned long long)( 16 * mb_w ) >> ( i ? 1 : 0 ), h = (unsigned long long)( 16 * mb_h ) >> ( i ? 1 : 0 );
            h_stats[i] = raw[i];
            h_stats[3 + i] = raw[3 + i] - ( raw[i] * raw[i] + w * h / 2 ) / ( w * h );
        }
    }
    return 0;
}
