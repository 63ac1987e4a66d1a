// Device-side block-metric primitives shared by the batched pixel kernels, the motion search and the
// lookahead kernels.  Unit of work everywhere: ONE LANE OWNS ONE 4x4 SUB-BLOCK (four rows of four 8-bit
// pixels = four 32-bit words per operand), so the 4x4 Hadamard of SATD needs no shuffles at all and a WxH
// block is (W/4)*(H/4) lanes whose partial sums are combined with a short segmented xor-shuffle.
//
// Reference semantics (jpsdr/x264, 8-bit):
//   SAD   common/pixel.c:55-80      sum |a-b|
//   SSD   common/pixel.c:85-110     sum (a-b)^2
//   SATD  common/pixel.c:242-332    sum over 4x4 tiles of ( sum |H4 (a-b) H4^T| ) >> 1
//   SA8D  common/pixel.c:334-381    ( sum over 8x8 tiles of sum |H8 (a-b) H8^T| + 2 ) >> 2
#pragma once
#include <stdint.h>

namespace x264cu {

enum { M_SAD = 0, M_SSD = 1, M_SATD = 2, M_SA8D = 3 };

__device__ __forceinline__ int dp4a_us( uint32_t a, int32_t s, int32_t c )
{
    int d;
    asm( "dp4a.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( s ), "r"( c ) );
    return d;
}

__device__ __forceinline__ uint32_t vsad4( uint32_t a, uint32_t b, uint32_t c )
{
    uint32_t d;
    asm( "vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}

// 4 pixels starting at an arbitrary byte address, from two aligned words
__device__ __forceinline__ uint32_t funnel( uint32_t lo, uint32_t hi, uint32_t shift_bits )
{
    return __funnelshift_r( lo, hi, shift_bits );
}

// Horizontal 4-point Hadamard of the pixel differences of one row, without unpacking the bytes:
// coefficient k = dot(a, S_k) - dot(b, S_k) with S_k in {+1,-1}^4, i.e. two DP4A each.
__device__ __forceinline__ void hrow( uint32_t a, uint32_t b, int h[4] )
{
    h[0] = dp4a_us( a, 0x01010101, dp4a_us( b, (int)0xFFFFFFFF, 0 ) );   // + + + +
    h[1] = dp4a_us( a, (int)0xFFFF0101, dp4a_us( b, 0x0101FFFF, 0 ) );   // + + - -
    h[2] = dp4a_us( a, 0x01FFFF01, dp4a_us( b, (int)0xFF0101FF, 0 ) );   // + - - +
    h[3] = dp4a_us( a, (int)0xFF01FF01, dp4a_us( b, 0x01FF01FF, 0 ) );   // + - + -
}

// SATD of one 4x4 block, already halved: the last vertical butterfly stage is folded into
// |p+q| + |p-q| = 2*max(|p|,|q|), so sum(max) is exactly (sum |coef|) >> 1 (pixel.c:262-286).
__device__ __forceinline__ int satd4x4( const uint32_t a[4], const uint32_t b[4] )
{
    int h0[4], h1[4], h2[4], h3[4];
    hrow( a[0], b[0], h0 );
    hrow( a[1], b[1], h1 );
    hrow( a[2], b[2], h2 );
    hrow( a[3], b[3], h3 );
    int acc = 0;
#pragma unroll
    for( int k = 0; k < 4; k++ )
    {
        int p0 = h0[k] + h1[k], p1 = h0[k] - h1[k];
        int p2 = h2[k] + h3[k], p3 = h2[k] - h3[k];
        acc += max( abs( p0 ), abs( p2 ) ) + max( abs( p1 ), abs( p3 ) );
    }
    return acc;
}

// full 2-D 4x4 Hadamard coefficients (unnormalised) for SA8D's cross-lane stages
__device__ __forceinline__ void had4x4( const uint32_t a[4], const uint32_t b[4], int c[16] )
{
    int h0[4], h1[4], h2[4], h3[4];
    hrow( a[0], b[0], h0 );
    hrow( a[1], b[1], h1 );
    hrow( a[2], b[2], h2 );
    hrow( a[3], b[3], h3 );
#pragma unroll
    for( int k = 0; k < 4; k++ )
    {
        int p0 = h0[k] + h1[k], p1 = h0[k] - h1[k];
        int p2 = h2[k] + h3[k], p3 = h2[k] - h3[k];
        c[k] = p0 + p2; c[4+k] = p0 - p2; c[8+k] = p1 + p3; c[12+k] = p1 - p3;
    }
}

__device__ __forceinline__ int sad4x4( const uint32_t a[4], const uint32_t b[4] )
{
    uint32_t s = vsad4( a[0], b[0], 0 );
    s = vsad4( a[1], b[1], s );
    s = vsad4( a[2], b[2], s );
    s = vsad4( a[3], b[3], s );
    return (int)s;
}

__device__ __forceinline__ int ssd4x4( const uint32_t a[4], const uint32_t b[4] )
{
    uint32_t s = 0;
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        uint32_t d = __vabsdiffu4( a[r], b[r] );
        s = __dp4a( d, d, s );
    }
    return (int)s;
}

// Lane geometry of a warp task: the warp covers a 32x16-pixel region = 8x4 sub-blocks of 4x4;
// qx = lane&7, qy = lane>>3.  A WxH block occupies (W/4)x(H/4) of those lanes.
template <int BW, int BH> struct BlockGeom
{
    static constexpr int LX = BW / 4, LY = BH / 4;            // lanes per block in x / y
    static constexpr int CX = 8 / LX, CY = 4 / LY;            // blocks per warp task in x / y
    static constexpr int CPT = CX * CY;                       // candidates per warp task
    __device__ static __forceinline__ int cand_in_task( int lane ) { return ( ( lane >> 3 ) / LY ) * CX + ( lane & 7 ) / LX; }
    __device__ static __forceinline__ int sub_x( int lane ) { return ( ( lane & 7 ) % LX ) * 4; }
    __device__ static __forceinline__ int sub_y( int lane ) { return ( ( lane >> 3 ) % LY ) * 4; }
    __device__ static __forceinline__ bool leader( int lane ) { return ( lane & 7 ) % LX == 0 && ( lane >> 3 ) % LY == 0; }
    // sum over the lanes of one block; every lane of the block ends up with the total
    __device__ static __forceinline__ int reduce( int v )
    {
        if( LX >= 2 ) v += __shfl_xor_sync( 0xffffffffu, v, 1 );
        if( LX >= 4 ) v += __shfl_xor_sync( 0xffffffffu, v, 2 );
        if( LY >= 2 ) v += __shfl_xor_sync( 0xffffffffu, v, 8 );
        if( LY >= 4 ) v += __shfl_xor_sync( 0xffffffffu, v, 16 );
        return v;
    }
};

// Per-lane partial metric of one 4x4 sub-block; SA8D does its two cross-lane butterfly stages here
// (lane^1 = horizontal neighbour, lane^8 = vertical neighbour inside the same 8x8).
template <int METRIC>
__device__ __forceinline__ int metric4x4( const uint32_t a[4], const uint32_t b[4], int lane )
{
    if( METRIC == M_SAD ) return sad4x4( a, b );
    if( METRIC == M_SSD ) return ssd4x4( a, b );
    if( METRIC == M_SATD ) return satd4x4( a, b );
    int c[16];
    had4x4( a, b, c );
    int acc = 0;
#pragma unroll
    for( int i = 0; i < 16; i++ )
    {
        int t = __shfl_xor_sync( 0xffffffffu, c[i], 1 );
        int v = ( lane & 1 ) ? t - c[i] : c[i] + t;
        t = __shfl_xor_sync( 0xffffffffu, v, 8 );
        v = ( lane & 8 ) ? t - v : v + t;
        acc += abs( v );
    }
    return acc;
}

// final normalisation once the lanes of a block are summed
template <int METRIC> __device__ __forceinline__ int metric_finish( int v )
{
    return METRIC == M_SA8D ? ( v + 2 ) >> 2 : v;
}

} // namespace x264cu
