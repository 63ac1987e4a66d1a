// Input staging: the plane-copy entries of x264_mc_functions_t (plane_copy, plane_copy_swap, plane_copy_interleave,
// plane_copy_deinterleave; common/mc.c:294-332) as device kernels, and x264_frame_copy_picture (common/frame.c:363-480) for the
// 8-bit 4:2:0 colour spaces on top of them: a picture in host memory becomes the reference's internal frame layout in HBM -- a luma
// plane and ONE interleaved chroma plane (NV12) -- which is what the motion search's chroma ME and the encoder's later stages read.
// Pure streaming kernels: a thread moves 4 (interleave: 2 x 4 -> 8) bytes; copies ride on cudaMemcpy2DAsync.
#include "ctx.h"

namespace {

__global__ void __launch_bounds__( 256 )
interleave_kernel( uint8_t *__restrict__ dst, intptr_t dst_stride, const uint8_t *__restrict__ su, intptr_t su_stride,
                   const uint8_t *__restrict__ sv, intptr_t sv_stride, int w, int h, int fast )
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int x = gx * 4;
    if( x >= w || y >= h ) return;
    const uint8_t *pu = su + (intptr_t)y * su_stride + x, *pv = sv + (intptr_t)y * sv_stride + x;
    uint8_t *pd = dst + (intptr_t)y * dst_stride + 2 * x;
    if( fast && x + 4 <= w )
    {   // 4 + 4 bytes in, 8 out: dst[2x] = u[x], dst[2x+1] = v[x]
        const uint32_t u = *(const uint32_t *)pu, v = *(const uint32_t *)pv;
        *(uint2 *)pd = make_uint2( __byte_perm( u, v, 0x5140 ), __byte_perm( u, v, 0x7362 ) );
        return;
    }
    for( int i = 0; i < 4 && x + i < w; i++ ) { pd[2*i] = pu[i]; pd[2*i+1] = pv[i]; }
}

__global__ void __launch_bounds__( 256 )
deinterleave_kernel( uint8_t *__restrict__ da, intptr_t da_stride, uint8_t *__restrict__ db, intptr_t db_stride,
                     const uint8_t *__restrict__ src, intptr_t src_stride, int w, int h, int fast )
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int x = gx * 4;
    if( x >= w || y >= h ) return;
    const uint8_t *ps = src + (intptr_t)y * src_stride + 2 * x;
    uint8_t *pa = da + (intptr_t)y * da_stride + x, *pb = db + (intptr_t)y * db_stride + x;
    if( fast && x + 4 <= w )
    {
        const uint2 s = *(const uint2 *)ps;
        *(uint32_t *)pa = __byte_perm( s.x, s.y, 0x6420 );
        *(uint32_t *)pb = __byte_perm( s.x, s.y, 0x7531 );
        return;
    }
    for( int i = 0; i < 4 && x + i < w; i++ ) { pa[i] = ps[2*i]; pb[i] = ps[2*i+1]; }
}

// w = number of byte PAIRS per row
__global__ void __launch_bounds__( 256 )
swap_kernel( uint8_t *__restrict__ dst, intptr_t dst_stride, const uint8_t *__restrict__ src, intptr_t src_stride, int w, int h, int fast )
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int x = gx * 2;                                    // pairs
    if( x >= w || y >= h ) return;
    const uint8_t *ps = src + (intptr_t)y * src_stride + 2 * x;
    uint8_t *pd = dst + (intptr_t)y * dst_stride + 2 * x;
    if( fast && x + 2 <= w )
    {
        *(uint32_t *)pd = __byte_perm( *(const uint32_t *)ps, 0, 0x2301 );
        return;
    }
    for( int i = 0; i < 2 && x + i < w; i++ ) { pd[2*i] = ps[2*i+1]; pd[2*i+1] = ps[2*i]; }
}

inline bool aligned( const void *p, intptr_t stride, int a ) { return !( ( (uintptr_t)p | (uintptr_t)stride ) & ( a - 1 ) ); }

}

extern "C" {

/* h->mc.plane_copy_interleave (common/mc.c:317-327): dst[2x] = srcu[x], dst[2x+1] = srcv[x]; w x h pairs */
int x264cu_plane_copy_interleave( x264cu_ctx_t *ctx, uint8_t *d_dst, intptr_t dst_stride, const uint8_t *d_srcu, intptr_t srcu_stride,
                                  const uint8_t *d_srcv, intptr_t srcv_stride, int w, int h )
{
    X264CU_ENTER( ctx );
    if( !ctx || !d_dst || !d_srcu || !d_srcv ) return -1;
    if( w <= 0 || h <= 0 ) return 0;
    const int fast = aligned( d_dst, dst_stride, 8 ) && aligned( d_srcu, srcu_stride, 4 ) && aligned( d_srcv, srcv_stride, 4 );
    interleave_kernel<<<dim3( ( ( w + 3 ) / 4 + 255 ) / 256, h ), 256, 0, ctx->stream>>>( d_dst, dst_stride, d_srcu, srcu_stride, d_srcv, srcv_stride, w, h, fast );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

/* h->mc.plane_copy_deinterleave (common/mc.c:329-339): dsta[x] = src[2x], dstb[x] = src[2x+1] */
int x264cu_plane_copy_deinterleave( x264cu_ctx_t *ctx, uint8_t *d_dsta, intptr_t dsta_stride, uint8_t *d_dstb, intptr_t dstb_stride,
                                    const uint8_t *d_src, intptr_t src_stride, int w, int h )
{
    X264CU_ENTER( ctx );
    if( !ctx || !d_dsta || !d_dstb || !d_src ) return -1;
    if( w <= 0 || h <= 0 ) return 0;
    const int fast = aligned( d_src, src_stride, 8 ) && aligned( d_dsta, dsta_stride, 4 ) && aligned( d_dstb, dstb_stride, 4 );
    deinterleave_kernel<<<dim3( ( ( w + 3 ) / 4 + 255 ) / 256, h ), 256, 0, ctx->stream>>>( d_dsta, dsta_stride, d_dstb, dstb_stride, d_src, src_stride, w, h, fast );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

/* h->mc.plane_copy_swap (common/mc.c:305-315): w byte pairs per row, each swapped */
int x264cu_plane_copy_swap( x264cu_ctx_t *ctx, uint8_t *d_dst, intptr_t dst_stride, const uint8_t *d_src, intptr_t src_stride, int w, int h )
{
    X264CU_ENTER( ctx );
    if( !ctx || !d_dst || !d_src ) return -1;
    if( w <= 0 || h <= 0 ) return 0;
    const int fast = aligned( d_dst, dst_stride, 4 ) && aligned( d_src, src_stride, 4 );
    swap_kernel<<<dim3( ( ( w + 1 ) / 2 + 255 ) / 256, h ), 256, 0, ctx->stream>>>( d_dst, dst_stride, d_src, src_stride, w, h, fast );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

/* x264_frame_copy_picture (common/frame.c:363-480) for the 8-bit 4:2:0 colour spaces: i_csp = X264_CSP_I420 (2), YV12 (3), NV12 (4)
 * or NV21 (5), optionally | X264_CSP_VFLIP (0x1000).  h_plane / stride as in x264_image_t.  The frame in HBM gets the reference's
 * internal layout: luma in d_luma, Cb / Cr interleaved in d_chroma ((width/2) pairs x (height/2) rows).  Strides of negative sign
 * are not needed from the caller: VFLIP says it. */
int x264cu_frame_copy_picture( x264cu_ctx_t *ctx, int i_csp, const uint8_t *const h_plane[3], const int stride[3], int width, int height,
                               uint8_t *d_luma, intptr_t luma_stride, uint8_t *d_chroma, intptr_t chroma_stride )
{
    X264CU_ENTER( ctx );
    if( !ctx || !h_plane || !stride || !d_luma ) return -1;
    const int csp = i_csp & 0xff, vflip = ( i_csp & 0x1000 ) != 0;
    if( csp < 2 || csp > 5 || ( i_csp & 0x2000 ) ) return x264cu_fail( ctx, "frame_copy_picture: colour space 0x%x is not an 8-bit 4:2:0 one", i_csp );
    if( width < 2 || height < 2 ) return x264cu_fail( ctx, "frame_copy_picture: bad size" );
    const int cw = width >> 1, ch = height >> 1;
    if( width > abs( stride[0] ) ) return x264cu_fail( ctx, "frame_copy_picture: width %d is greater than stride %d", width, stride[0] );
    // get_plane_ptr (frame.c:342-358): a flipped picture is read from its last row upwards
    auto first_row = [&]( int plane, int rows ) { return h_plane[plane] + ( vflip ? (intptr_t)( rows - 1 ) * stride[plane] : 0 ); };
    auto pitch = [&]( int plane ) { return vflip ? -(intptr_t)stride[plane] : (intptr_t)stride[plane]; };
    // host rows -> device staging: cudaMemcpy2D wants a positive pitch, so a flipped plane goes row by row into the staging buffer
    auto upload = [&]( uint8_t *d, intptr_t d_stride, int plane, int bytes, int rows ) -> int {
        if( !h_plane[plane] ) return x264cu_fail( ctx, "frame_copy_picture: plane %d missing", plane );
        if( !vflip )
            CU_CHECK( ctx, cudaMemcpy2DAsync( d, d_stride, h_plane[plane], stride[plane], bytes, rows, cudaMemcpyHostToDevice, ctx->stream ) );
        else
            for( int y = 0; y < rows; y++ )
                CU_CHECK( ctx, cudaMemcpyAsync( d + (intptr_t)y * d_stride, first_row( plane, rows ) + (intptr_t)y * pitch( plane ), bytes,
                                                cudaMemcpyHostToDevice, ctx->stream ) );
        return 0;
    };
    if( upload( d_luma, luma_stride, 0, width, height ) ) return -1;
    if( !d_chroma ) return 0;
    if( csp == 4 )                                                         // NV12: as it is
        return upload( d_chroma, chroma_stride, 1, 2 * cw, ch );
    const intptr_t st = ( 2 * cw + 63 ) & ~(intptr_t)63;
    uint8_t *tmp = (uint8_t *)x264cu_scratch( ctx, 13, (size_t)st * ch + 64 );
    if( !tmp ) return -1;
    if( csp == 5 )
    {   // NV21: pairs swapped
        if( upload( tmp, st, 1, 2 * cw, ch ) ) return -1;
        return x264cu_plane_copy_swap( ctx, d_chroma, chroma_stride, tmp, st, cw, ch );
    }
    // I420 / YV12: two planes interleaved (YV12 carries Cr first)
    const int pu = csp == 3 ? 2 : 1, pv = csp == 3 ? 1 : 2;
    const intptr_t half = ( cw + 63 ) & ~(intptr_t)63;
    uint8_t *tu = tmp, *tv = (uint8_t *)x264cu_scratch( ctx, 14, (size_t)half * ch + 64 );
    if( !tv ) return -1;
    if( upload( tu, half, pu, cw, ch ) || upload( tv, half, pv, cw, ch ) ) return -1;
    return x264cu_plane_copy_interleave( ctx, d_chroma, chroma_stride, tu, half, tv, half, cw, ch );
}

} // extern "C"
