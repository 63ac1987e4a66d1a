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
__device__ __forceinline__ uint32_t filt4( const int *r0, const int *r1, int o )
{
    // FILTER(a,b,c,d) = ((((a+b+1)>>1)+((c+d+1)>>1)+1)>>1) with a,b vertical pair at x, c,d at x+1 (mc.c:494-500)
    uint32_t out = 0;
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int a = r0[2*i+o], b = r1[2*i+o], c = r0[2*i+1+o], d = r1[2*i+1+o];
        int v = ( ( ( a + b + 1 ) >> 1 ) + ( ( c + d + 1 ) >> 1 ) + 1 ) >> 1;
        out |= (uint32_t)v << ( 8*i );
    }
    return out;
}

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
    int rows[3][9];
    if( fast_ok && oy == y && ox >= 0 && ox + 4 <= wl && 2*( ox + 4 ) + 1 <= width && 2*y + 2 < height )
    {   // interior: rows 2y..2y+2, columns 2ox .. 2ox+8, all inside the picture; 8-byte aligned vector loads
#pragma unroll
        for( int r = 0; r < 3; r++ )
        {
            const uint8_t *p = src + (intptr_t)( 2*y + r ) * src_stride + 2*ox;
            uint2 v = *(const uint2 *)p;
            rows[r][0] = v.x & 255; rows[r][1] = ( v.x >> 8 ) & 255; rows[r][2] = ( v.x >> 16 ) & 255; rows[r][3] = v.x >> 24;
            rows[r][4] = v.y & 255; rows[r][5] = ( v.y >> 8 ) & 255; rows[r][6] = ( v.y >> 16 ) & 255; rows[r][7] = v.y >> 24;
            rows[r][8] = p[8];
        }
        *(uint32_t *)( d0 + (intptr_t)oy*dst_stride + ox ) = filt4( rows[0], rows[1], 0 );
        *(uint32_t *)( dh + (intptr_t)oy*dst_stride + ox ) = filt4( rows[0], rows[1], 1 );
        *(uint32_t *)( dv + (intptr_t)oy*dst_stride + ox ) = filt4( rows[1], rows[2], 0 );
        *(uint32_t *)( dc + (intptr_t)oy*dst_stride + ox ) = filt4( rows[1], rows[2], 1 );
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
    if( !ctx ) return -1;
    return x264cu_frame_init_lowres_on( ctx, ctx->stream, d_luma, luma_stride, width, height, d_lowres, lowres_stride );
}

int x264cu_hpel_filter( x264cu_ctx_t *ctx, uint8_t *d_src, intptr_t stride, int width, int height,
                        uint8_t *d_h, uint8_t *d_v, uint8_t *d_c, int expand_src )
{
    if( !ctx ) return -1;
    if( width < 1 || height < 1 ) return x264cu_fail( ctx, "hpel_filter: bad size" );
    dim3 grid( ( width + 2*X264CU_PAD + HT_W - 1 ) / HT_W, ( height + 2*X264CU_PAD + HT_H - 1 ) / HT_H );
    hpel_kernel<<<grid, 256, 0, ctx->stream>>>( d_src, stride, width, height, d_h, d_v, d_c, expand_src ? d_src : nullptr );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

} // extern "C"
