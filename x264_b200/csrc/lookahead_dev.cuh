// Device code of the lowres lookahead: one WARP runs one 8x8 lowres macroblock's motion search with exactly the
// reference's control flow (encoder/me.c:182-992, restricted to what lowres_context_init selects:
// DIA/HEX, subpel_refine 2 or 4, no chroma ME; encoder/slicetype.c:45-61), evaluating the candidates of each
// step in parallel: 4 lanes per candidate (one 4x4 quadrant each), up to 8 candidates per step.  Sequential
// "first strictly smaller wins" chains (COPY1_IF_LT on packed (cost<<k)+tag values, me.c:325-341, :369-418)
// become a warp-wide integer minimum over the same packed keys, which picks the same winner.
#pragma once
#include "pixel_dev.cuh"

namespace x264cu {

#define LA_COST_MAX ( 1 << 28 )                 /* encoder/me.h:30 */

// phase timing of the search kernel (build with LA_PROFILE=1 in the environment; tools/la_phase_profile.py)
#ifdef LA_PROFILE
#define LA_TICK( acc, t ) do { long long _n = clock64(); ( acc ) += _n - ( t ); ( t ) = _n; } while( 0 )
#else
#define LA_TICK( acc, t ) do {} while( 0 )
#endif

__device__ __forceinline__ uint32_t ldg4u( const uint8_t *p )      // 4 pixels at any byte address
{
    uintptr_t u = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)( u & ~(uintptr_t)3 );
    uint32_t sh = ( (uint32_t)u & 3u ) * 8u;
    uint32_t lo = __ldg( q ), hi = __ldg( q + 1 );                 // planes are padded: q+1 is always readable
    return __funnelshift_r( lo, hi, sh );
}

struct LaWeight { int enabled, scale, denom, offset; };

__device__ __forceinline__ uint32_t weight4( uint32_t v, const LaWeight &w )   // mc_weight, common/mc.c:117-137
{
    uint32_t out = 0;
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int p = ( v >> ( 8*i ) ) & 255;
        int r = w.denom >= 1 ? ( ( p * w.scale + ( 1 << ( w.denom - 1 ) ) ) >> w.denom ) + w.offset : p * w.scale + w.offset;
        r = min( max( r, 0 ), 255 );
        out |= (uint32_t)r << ( 8*i );
    }
    return out;
}

// four rows at once, out of line: the search kernel calls it from many inlined sites but only for weighted references
static __device__ __noinline__ uint4 weight4x4( uint4 v, int enabled, int scale, int denom, int offset )
{
    LaWeight w = { enabled, scale, denom, offset };
    return make_uint4( weight4( v.x, w ), weight4( v.y, w ), weight4( v.z, w ), weight4( v.w, w ) );
}

// Per-warp shared-memory window of the four reference planes around the warp's macroblock row.  The search warp walks
// its row right to left; the window is, per plane, LA_WIN_ROWS rows of a 64-byte ring of columns (8 chunks of 8 px, one
// per macroblock step: byte = biased column & 63) at a pitch of 72 bytes -- the 4 rows of a 4x4 read sit at immediate
// offsets, and the quadrant 4 rows further down falls into other banks.  9 KB per warp.  A candidate whose pixels lie outside the loaded chunks / rows (motion beyond
// +-12 lowres pixels vertically, about +-24 horizontally) is read from global memory instead (same values, slower).
#ifndef LA_WIN_ROWS
#define LA_WIN_ROWS 32                                 /* measured: 32 rows (motion within +-12) beat 48 by 16 % at 4K: a third fewer */
#define LA_WIN_VR 12                                   /* cp.async per step.  Rows above the MB row: window rows = -12 .. +19 */
#endif
#define LA_WIN_PITCH 72
#define LA_WIN_PLANE ( LA_WIN_ROWS * LA_WIN_PITCH )
#define LA_WIN_BYTES ( 4 * LA_WIN_PLANE )
struct LaWin
{
    uint32_t base;                // shared-space address of this warp's window
    int bx;                       // this lane's quadrant origin, biased column (plane x + 64 >= 0)
    int ry;                       // this lane's quadrant origin, window row
    int dxlo;                     // in the window <=> (unsigned)( dx - dxlo ) <= dxspan && (unsigned)( dy + ry ) <= LA_WIN_ROWS - 4
    unsigned dxspan;
    bool on;                      // window in use (search kernel) or not (finalize)
    bool p0w;                     // window plane 0 holds the WEIGHTED full-pel plane, not F
};

template <int OFF>
__device__ __forceinline__ uint32_t la_lds( uint32_t addr )
{
    uint32_t v;
    asm volatile( "ld.shared.u32 %0, [%1+%2];" : "=r"( v ) : "r"( addr ), "n"( OFF ) );
    return v;
}

// warp-uniform search context + per-lane fenc quadrant
struct LaMe
{
    LaWin win;
    const uint8_t *fref[4];       // F,H,V,C plane pointers at this lane's 4x4 quadrant of the MB
    const uint8_t *fref_w;        // weighted full-pel plane (== fref[0] without weights)
    int stride;
    uint32_t fenc[4];             // this lane's fenc quadrant rows
    const uint16_t *cost_mv;      // centred table in global memory (the rare far entries)
    uint32_t cost_s;              // shared-space address of entry 0 of the table's centre part, |index| <= LA_COST_HALF
    int mvpx, mvpy;
    int min_spel_x, min_spel_y, max_spel_x, max_spel_y;
    int x_min, y_min, x_max, y_max;
    LaWeight w;
    bool satd;                    // mbcmp is SATD
#ifdef LA_PROFILE
    long long prof[3];            // cycles in: predictors, full-pel search, sub-pel refine
#endif
};

#define LA_COST_HALF 512          /* mv - mvp distances (quarter-pel) kept in shared memory: +-128 pixels */

// p_cost_mv[idx] (analyse.c:143-202 table, centred): the entries a search normally touches are in shared memory; a vector
// further than 128 pixels from its predictor takes the out-of-line global read
static __device__ __noinline__ int la_cost_far( const uint16_t *cost_mv, int idx ) { return __ldg( cost_mv + idx ); }
__device__ __forceinline__ int la_cost( const LaMe &m, int idx )
{
    if( (unsigned)( idx + LA_COST_HALF ) <= 2u * LA_COST_HALF )
    {
        uint32_t v;
        asm volatile( "ld.shared.u16 %0, [%1];" : "=r"( v ) : "r"( m.cost_s + 2 * idx ) );
        return (int)v;
    }
    return la_cost_far( m.cost_mv, idx );
}

#define LA_PLANE_W 4              /* la_load4 plane selector: the (possibly weighted) full-pel search plane */

// The rare path of la_load4, out of line (one copy instead of one per call site): some lane's block lies outside the
// shared-memory window (motion beyond the window's reach, or the unweighted F plane while the window holds the weighted
// one).  Lanes inside the window still read it; the others read the plane itself.  Scalars only: a reference to the
// register-resident LaMe would force it into local memory.
static __device__ __noinline__ uint4 la_load4_mixed( bool in_win, uint32_t pb, int x, const uint8_t *g, int stride )
{
    uint4 v;
    if( in_win )
    {
        const uint32_t a0 = pb + ( x & 60 ), a1 = pb + ( ( x + 4 ) & 60 );
        const uint32_t sh = ( (uint32_t)x & 3u ) * 8u;
        v.x = __funnelshift_r( la_lds<0>( a0 ), la_lds<0>( a1 ), sh );
        v.y = __funnelshift_r( la_lds<LA_WIN_PITCH>( a0 ), la_lds<LA_WIN_PITCH>( a1 ), sh );
        v.z = __funnelshift_r( la_lds<2 * LA_WIN_PITCH>( a0 ), la_lds<2 * LA_WIN_PITCH>( a1 ), sh );
        v.w = __funnelshift_r( la_lds<3 * LA_WIN_PITCH>( a0 ), la_lds<3 * LA_WIN_PITCH>( a1 ), sh );
    }
    else
    {
        v.x = ldg4u( g ); v.y = ldg4u( g + stride ); v.z = ldg4u( g + 2 * stride ); v.w = ldg4u( g + 3 * stride );
    }
    return v;
}

// this lane's 4x4 of plane `plane` (0..3 = F,H,V,C unweighted, LA_PLANE_W = fref_w) displaced by (dx,dy) full pixels.
// With the window on (search kernel) every lane of the warp must call it together (warp votes inside).
__device__ __forceinline__ void la_load4( const LaMe &m, int plane, int dx, int dy, uint32_t b[4] )
{
    const LaWin &w = m.win;
    if( !w.on )
    {   // finalize / weight kernels: straight from the planes (L2-resident)
        const uint8_t *s = plane == LA_PLANE_W ? m.fref_w : plane == 0 ? m.fref[0] : plane == 1 ? m.fref[1] : plane == 2 ? m.fref[2] : m.fref[3];
        s += dy * m.stride + dx;
#pragma unroll
        for( int i = 0; i < 4; i++ ) b[i] = ldg4u( s + i * m.stride );
        return;
    }
    const int x = w.bx + dx, r = w.ry + dy;
    const bool in_win = (unsigned)( dx - w.dxlo ) <= w.dxspan && (unsigned)r <= LA_WIN_ROWS - 4 && !( plane == 0 && w.p0w );
    const int slot = plane == LA_PLANE_W ? 0 : plane;
    const uint32_t pb = w.base + slot * LA_WIN_PLANE + r * LA_WIN_PITCH;
    if( __all_sync( 0xffffffffu, in_win ) )
    {
        const uint32_t a0 = pb + ( x & 60 ), a1 = pb + ( ( x + 4 ) & 60 );
        const uint32_t sh = ( (uint32_t)x & 3u ) * 8u;
        b[0] = __funnelshift_r( la_lds<0>( a0 ), la_lds<0>( a1 ), sh );
        b[1] = __funnelshift_r( la_lds<LA_WIN_PITCH>( a0 ), la_lds<LA_WIN_PITCH>( a1 ), sh );
        b[2] = __funnelshift_r( la_lds<2 * LA_WIN_PITCH>( a0 ), la_lds<2 * LA_WIN_PITCH>( a1 ), sh );
        b[3] = __funnelshift_r( la_lds<3 * LA_WIN_PITCH>( a0 ), la_lds<3 * LA_WIN_PITCH>( a1 ), sh );
    }
    else
    {
        const uint8_t *s = plane == LA_PLANE_W ? m.fref_w : plane == 0 ? m.fref[0] : plane == 1 ? m.fref[1] : plane == 2 ? m.fref[2] : m.fref[3];
        const uint4 v = la_load4_mixed( in_win, pb, x, s + dy * m.stride + dx, m.stride );
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    }
}

// get_ref (common/mc.c:198-249, tables.c:183-184): this lane's 4x4 of the block interpolated at quarter-pel mv.
// Warp-uniform control flow: the second plane is read by every lane as soon as one lane needs it (a lane that does not
// reads its first block again: (a+a+1)>>1 == a).
__device__ __forceinline__ void qpel4x4( const LaMe &m, int mvx, int mvy, uint32_t b[4] )
{
    // hpel_ref0 = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1}, hpel_ref1 = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2}: 2 bits each
    const uint32_t R0 = 0x54FE5454u;    // idx 0..15, 2 bits per entry: 0,1,1,1, 0,1,1,1, 2,3,3,3, 0,1,1,1
    const uint32_t R1 = 0xBABABA10u;    // 0,0,1,0, 2,2,3,2, 2,2,3,2, 2,2,3,2
    const int idx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    const int fx = mvx >> 2, fy = mvy >> 2;
    const int p0 = ( R0 >> ( 2*idx ) ) & 3, y0 = fy + ( ( mvy & 3 ) == 3 );
    la_load4( m, p0, fx, y0, b );
    const bool two = ( idx & 5 ) != 0;
    if( !m.win.on ? two : __any_sync( 0xffffffffu, two ) )
    {
        uint32_t c[4];
        la_load4( m, two ? ( R1 >> ( 2*idx ) ) & 3 : p0, two ? fx + ( ( mvx & 3 ) == 3 ) : fx, two ? fy : y0, c );
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = __vavgu4( b[r], c[r] );      // (a+b+1)>>1
    }
    if( m.w.enabled )
    {
        const uint4 v = weight4x4( make_uint4( b[0], b[1], b[2], b[3] ), 1, m.w.scale, m.w.denom, m.w.offset );
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    }
}

// same as qpel4x4 with the plane pointers passed individually (selected with ?: instead of an indexed local array)
__device__ __forceinline__ void qpel4x4_p( const uint8_t *f0, const uint8_t *f1, const uint8_t *f2, const uint8_t *f3, int stride,
                                           const LaWeight &w, int mvx, int mvy, uint32_t b[4] )
{
    const uint32_t R0 = 0x54FE5454u, R1 = 0xBABABA10u;
    const int idx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    const int off = ( mvy >> 2 ) * stride + ( mvx >> 2 );
    const int k0 = ( R0 >> ( 2*idx ) ) & 3, k1 = ( R1 >> ( 2*idx ) ) & 3;
    const uint8_t *s1 = ( k0 == 0 ? f0 : k0 == 1 ? f1 : k0 == 2 ? f2 : f3 ) + off + ( ( mvy & 3 ) == 3 ? stride : 0 );
    if( idx & 5 )
    {
        const uint8_t *s2 = ( k1 == 0 ? f0 : k1 == 1 ? f1 : k1 == 2 ? f2 : f3 ) + off + ( ( mvx & 3 ) == 3 ? 1 : 0 );
#pragma unroll
        for( int r = 0; r < 4; r++ )
            b[r] = __vavgu4( ldg4u( s1 + r * stride ), ldg4u( s2 + r * stride ) );
    }
    else
    {
#pragma unroll
        for( int r = 0; r < 4; r++ )
            b[r] = ldg4u( s1 + r * stride );
    }
    if( w.enabled )
    {
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = weight4( b[r], w );
    }
}

__device__ __forceinline__ int quad_sum( int v )          // sum over the 4 lanes of one candidate slot
{
    v += __shfl_xor_sync( 0xffffffffu, v, 1 );
    v += __shfl_xor_sync( 0xffffffffu, v, 2 );
    return v;
}

// SAD of this slot's full-pel candidate against the weighted full-pel plane (COST_MV without the mv bits)
__device__ __forceinline__ int la_sad_fpel( const LaMe &m, int mx, int my )
{
    uint32_t b[4];
    la_load4( m, LA_PLANE_W, mx, my, b );
    return quad_sum( sad4x4( m.fenc, b ) );
}
__device__ __forceinline__ int la_bits_fpel( const LaMe &m, int mx, int my )      // BITS_MVD, me.c:60-61
{
    return la_cost( m, mx*4 - m.mvpx ) + la_cost( m, my*4 - m.mvpy );
}
// quarter-pel candidate, SAD (fpelcmp) or mbcmp (COST_MV_HPEL / COST_MV_SAD / COST_MV_SATD), incl. mv bits
__device__ __forceinline__ int la_cost_qpel( const LaMe &m, int mx, int my, bool use_mbcmp )
{
    uint32_t b[4];
    qpel4x4( m, mx, my, b );
    int d = ( use_mbcmp && m.satd ) ? satd4x4( m.fenc, b ) : sad4x4( m.fenc, b );
    return quad_sum( d ) + la_cost( m, mx - m.mvpx ) + la_cost( m, my - m.mvpy );
}

__device__ __forceinline__ int warp_min( int v ) { return __reduce_min_sync( 0xffffffffu, v ); }
__device__ __forceinline__ int clip3i( int v, int lo, int hi ) { return min( max( v, lo ), hi ); }
__device__ __forceinline__ uint32_t pack_mv( int x, int y ) { return ( (uint32_t)x & 0xFFFF ) + ( (uint32_t)y << 16 ); }
#define LA_FPEL( v ) ( ( ( v ) + 2 ) >> 2 )

// x264_me_search_ref for one lowres MB.  All control flow is warp-uniform; `slot` = lane>>2.
// mvc: up to 4 candidate vectors (qpel), i_mvc of them valid.  Results: mv (qpel) and cost, uniform.
static __device__ __forceinline__ void la_me_search( LaMe &m, int me_method, int subpel_refine, int me_range,
                                           const int *mvc_x, const int *mvc_y, int i_mvc, int lane,
                                           int &out_mvx, int &out_mvy, int &out_cost )
{
    const int slot = lane >> 2;
    int bmx, bmy, bcost = LA_COST_MAX, bpred_cost = LA_COST_MAX;
    uint32_t pmv, bpred_mv = 0;
#ifdef LA_PROFILE
    long long t_prof = clock64();
#endif

    if( subpel_refine >= 3 )
    {   // me.c:216-275: sub-pel predictors
        int bpx = clip3i( m.mvpx, m.x_min*4, m.x_max*4 ), bpy = clip3i( m.mvpy, m.y_min*4, m.y_max*4 );
        pmv = pack_mv( bpx, bpy );
        // x264_predictor_clip (common.h:791-806): compact the candidates that are neither zero nor == pmv
        int cx[4], cy[4], n = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ )
        {
            cx[i] = cy[i] = 0;
        }
#pragma unroll
        for( int i = 0; i < 4; i++ )
            if( i < i_mvc )
            {
                uint32_t mv = pack_mv( mvc_x[i], mvc_y[i] );
                if( mv && mv != pmv )
                {
                    int vx = clip3i( mvc_x[i], m.x_min*4, m.x_max*4 ), vy = clip3i( mvc_y[i], m.y_min*4, m.y_max*4 );
                    if( n == 0 ) { cx[0] = vx; cy[0] = vy; }
                    else if( n == 1 ) { cx[1] = vx; cy[1] = vy; }
                    else if( n == 2 ) { cx[2] = vx; cy[2] = vy; }
                    else { cx[3] = vx; cy[3] = vy; }
                    n++;
                }
            }
        // slot 0 = the predictor, slots 1..n = candidates.  The two full-pel probes that follow in the reference (the
        // rounded winner and the zero vector, me.c:247-272) do not depend on anything but the winner's identity: every
        // slot also measures ITS candidate rounded to full-pel (slot 5: the zero vector), the winner's is picked after.
        int sx = bpx, sy = bpy;
        if( slot == 1 ) { sx = cx[0]; sy = cy[0]; }
        if( slot == 2 ) { sx = cx[1]; sy = cy[1]; }
        if( slot == 3 ) { sx = cx[2]; sy = cy[2]; }
        if( slot == 4 ) { sx = cx[3]; sy = cy[3]; }
        if( slot >= 5 ) { sx = 0; sy = 0; }
        int c = la_cost_qpel( m, sx, sy, false );
        const int fx = LA_FPEL( sx ), fy = LA_FPEL( sy );
        const int fc = la_sad_fpel( m, fx, fy ) + la_bits_fpel( m, fx, fy );
        int pmv_cost = __shfl_sync( 0xffffffffu, c, 0 );
        bpred_cost = pmv_cost;
        int w = 0;
        if( n > 0 )
        {
            int key = slot <= n ? ( c << 4 ) + slot : 0x7fffffff;
            key = warp_min( key );
            w = key & 15;
            if( w == 1 ) { bpx = cx[0]; bpy = cy[0]; }
            if( w == 2 ) { bpx = cx[1]; bpy = cy[1]; }
            if( w == 3 ) { bpx = cx[2]; bpy = cy[2]; }
            if( w == 4 ) { bpx = cx[3]; bpy = cy[3]; }
            bpred_cost = key >> 4;
        }
        bmx = LA_FPEL( bpx ); bmy = LA_FPEL( bpy );
        bpred_mv = pack_mv( bpx, bpy );
        int c_bm = __shfl_sync( 0xffffffffu, fc, w * 4 ), c_zero = __shfl_sync( 0xffffffffu, fc, 20 );
        if( bpred_mv & 0x00030003 ) { if( c_bm < bcost ) bcost = c_bm; }
        else bcost = bpred_cost;
        if( pmv )
        {
            if( ( bmx | bmy ) && c_zero < bcost ) { bcost = c_zero; bmx = 0; bmy = 0; }
        }
        else if( pmv_cost < bcost ) { bcost = pmv_cost; bmx = 0; bmy = 0; }
    }
    else
    {   // me.c:277-318: rounded predictor measured WITHOUT its mv cost, rounded candidates, zero vector
        bmx = clip3i( LA_FPEL( m.mvpx ), m.x_min, m.x_max );
        bmy = clip3i( LA_FPEL( m.mvpy ), m.y_min, m.y_max );
        pmv = pack_mv( bmx, bmy );
        int cx[4], cy[4], n = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ ) cx[i] = cy[i] = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ )
            if( i < i_mvc )
            {   // x264_predictor_roundclip, common.h:774-789
                int rx = ( mvc_x[i] + 2 ) >> 2, ry = ( mvc_y[i] + 2 ) >> 2;
                uint32_t mv = pack_mv( rx, ry );
                if( mv && mv != pmv )
                {
                    int vx = clip3i( rx, m.x_min, m.x_max ), vy = clip3i( ry, m.y_min, m.y_max );
                    if( n == 0 ) { cx[0] = vx; cy[0] = vy; }
                    else if( n == 1 ) { cx[1] = vx; cy[1] = vy; }
                    else if( n == 2 ) { cx[2] = vx; cy[2] = vy; }
                    else { cx[3] = vx; cy[3] = vy; }
                    n++;
                }
            }
        int sx = bmx, sy = bmy;                                   // slot 0: predictor; 1..4: candidates; 5: zero
        if( slot == 1 ) { sx = cx[0]; sy = cy[0]; }
        if( slot == 2 ) { sx = cx[1]; sy = cy[1]; }
        if( slot == 3 ) { sx = cx[2]; sy = cy[2]; }
        if( slot == 4 ) { sx = cx[3]; sy = cy[3]; }
        if( slot >= 5 ) { sx = 0; sy = 0; }
        int sad = la_sad_fpel( m, sx, sy );
        int c = sad + la_bits_fpel( m, sx, sy );
        bcost = __shfl_sync( 0xffffffffu, sad, 0 );
        int c_zero = __shfl_sync( 0xffffffffu, c, 20 );
        if( n > 0 )
        {
            int key = slot == 0 ? ( bcost << 4 ) : ( slot <= n ? ( c << 4 ) + slot : 0x7fffffff );
            key = warp_min( key );
            int w = key & 15;
            if( w == 1 ) { bmx = cx[0]; bmy = cy[0]; }
            if( w == 2 ) { bmx = cx[1]; bmy = cy[1]; }
            if( w == 3 ) { bmx = cx[2]; bmy = cy[2]; }
            if( w == 4 ) { bmx = cx[3]; bmy = cy[3]; }
            bcost = key >> 4;
        }
        if( pmv && c_zero < bcost ) { bcost = c_zero; bmx = 0; bmy = 0; }
    }

    auto in_range = [&]( int x, int y ) { return x >= m.x_min && x <= m.x_max && y >= m.y_min && y <= m.y_max; };
    LA_TICK( m.prof[0], t_prof );

    if( me_method == X264CU_ME_DIA )
    {   // me.c:322-342
        int i = me_range;
        do
        {
            const int dx = slot == 2 ? -1 : slot == 3 ? 1 : 0;
            const int dy = slot == 0 ? -1 : slot == 1 ? 1 : 0;
            const int tag = slot == 0 ? 1 : slot == 1 ? 3 : slot == 2 ? 4 : 12;
            int c = la_sad_fpel( m, bmx + dx, bmy + dy ) + la_bits_fpel( m, bmx + dx, bmy + dy );
            int key = slot < 4 ? ( c << 4 ) + tag : 0x7fffffff;
            key = min( warp_min( key ), bcost << 4 );
            if( !( key & 15 ) )
                break;
            bmx -= (int32_t)( (uint32_t)key << 28 ) >> 30;
            bmy -= (int32_t)( (uint32_t)key << 30 ) >> 30;
            bcost = key >> 4;
        } while( --i && in_range( bmx, bmy ) );
    }
    else
    {   // hexagon + square refine, me.c:344-420
        // hex2[k] = {-1,-2} {-2,0} {-1,2} {1,2} {2,0} {1,-2} {-1,-2} {-2,0} (me.c:55), packed as signed nibbles
        auto hex2x = []( int k ) { return (int32_t)( 0xEF121FEFu << ( 28 - 4*k ) ) >> 28; };
        auto hex2y = []( int k ) { return (int32_t)( 0x0EE0220Eu << ( 28 - 4*k ) ) >> 28; };
        // first ring: tags 2..7 = (-2,0) (-1,2) (1,2) (2,0) (1,-2) (-1,-2) = hex2[1..6]
        int key;
        {
            int k = min( slot + 1, 7 );
            int c = la_sad_fpel( m, bmx + hex2x( k ), bmy + hex2y( k ) ) + la_bits_fpel( m, bmx + hex2x( k ), bmy + hex2y( k ) );
            key = slot < 6 ? ( c << 3 ) + slot + 2 : 0x7fffffff;
            key = min( warp_min( key ), bcost << 3 );
        }
        if( key & 7 )
        {
            int dir = ( key & 7 ) - 2;
            bmx += hex2x( dir + 1 );
            bmy += hex2y( dir + 1 );
            for( int i = ( me_range >> 1 ) - 1; i > 0 && in_range( bmx, bmy ); i-- )
            {
                int k = min( dir + slot, 7 );
                int c = la_sad_fpel( m, bmx + hex2x( k ), bmy + hex2y( k ) ) + la_bits_fpel( m, bmx + hex2x( k ), bmy + hex2y( k ) );
                int k2 = slot < 3 ? ( c << 3 ) + slot + 1 : 0x7fffffff;
                key = min( warp_min( k2 ), key & ~7 );
                if( !( key & 7 ) )
                    break;
                dir += ( key & 7 ) - 2;
                dir = dir < 0 ? 5 : dir > 5 ? dir - 6 : dir;            // mod6m1[dir+1]
                bmx += hex2x( dir + 1 );
                bmy += hex2y( dir + 1 );
            }
        }
        bcost = key >> 3;
        // square1[1..8] = (0,-1) (0,1) (-1,0) (1,0) (-1,-1) (-1,1) (1,-1) (1,1)
        const int sqx = slot == 2 || slot == 4 || slot == 5 ? -1 : ( slot == 3 || slot == 6 || slot == 7 ? 1 : 0 );
        const int sqy = slot == 0 || slot == 4 || slot == 6 ? -1 : ( slot == 1 || slot == 5 || slot == 7 ? 1 : 0 );
        int c = la_sad_fpel( m, bmx + sqx, bmy + sqy ) + la_bits_fpel( m, bmx + sqx, bmy + sqy );
        int k3 = min( warp_min( ( c << 4 ) + slot + 1 ), bcost << 4 );
        int w = k3 & 15;
        if( w )
        {
            bmx += ( w == 3 || w == 5 || w == 6 ) ? -1 : ( w == 4 || w == 7 || w == 8 ) ? 1 : 0;
            bmy += ( w == 1 || w == 5 || w == 7 ) ? -1 : ( w == 2 || w == 6 || w == 8 ) ? 1 : 0;
        }
        bcost = k3 >> 4;
    }

    LA_TICK( m.prof[1], t_prof );
    // -> quarter-pel, me.c:774-789
    int mvx, mvy, cost;
    if( subpel_refine < 3 )
    {
        cost = bcost;
        if( pack_mv( bmx, bmy ) == pmv )
            cost += la_cost( m, bmx*4 - m.mvpx ) + la_cost( m, bmy*4 - m.mvpy );
        mvx = bmx*4; mvy = bmy*4;
    }
    else if( bpred_cost < bcost )
    {
        mvx = (int16_t)( bpred_mv & 0xFFFF ); mvy = (int16_t)( bpred_mv >> 16 );
        cost = bpred_cost;
    }
    else { mvx = bmx*4; mvy = bmy*4; cost = bcost; }

    // refine_subpel, me.c:865-992 with subpel_iterations[2] = {.,.,1,0}, [4] = {.,.,1,1}
    {
        int qx = mvx, qy = mvy, qcost = cost;
        // half-pel stage: optional predictor probe (slot 4) is evaluated together with the diamond only when
        // it cannot change the diamond's centre, i.e. never: it runs first, as in the reference
        if( subpel_refine < 3 )
        {
            int px = clip3i( m.mvpx, m.min_spel_x + 2, m.max_spel_x - 2 ), py = clip3i( m.mvpy, m.min_spel_y + 2, m.max_spel_y - 2 );
            if( ( px - qx ) | ( py - qy ) )
            {
                int c = la_cost_qpel( m, px, py, false );
                c = __shfl_sync( 0xffffffffu, c, 0 );
                if( c < qcost ) { qcost = c; qx = px; qy = py; }
            }
        }
        {   // one half-pel diamond iteration: (0,-2) (0,2) (-2,0) (2,0), tags 2,6,16,48
            const int dx = slot == 2 ? -2 : slot == 3 ? 2 : 0;
            const int dy = slot == 0 ? -2 : slot == 1 ? 2 : 0;
            const int tag = slot == 0 ? 2 : slot == 1 ? 6 : slot == 2 ? 16 : 48;
            int c = la_cost_qpel( m, qx + dx, qy + dy, false );
            int key = slot < 4 ? ( c << 6 ) + tag : 0x7fffffff;
            key = min( warp_min( key ), qcost << 6 );
            if( key & 63 )
            {
                qx -= (int32_t)( (uint32_t)key << 26 ) >> 29;
                qy -= (int32_t)( (uint32_t)key << 29 ) >> 29;
            }
            qcost = key >> 6;
        }
        // me.c:925-963: the winner is re-measured with mbcmp (when that is SATD) and one quarter-pel diamond iteration
        // follows (bdir = -1: nothing is skipped).  The re-measurement and the four diamond points only depend on the
        // half-pel winner: ONE step, slot 4 taking the centre.
        const bool do_qpel = subpel_refine >= 4 && !( qy <= m.min_spel_y || qy >= m.max_spel_y || qx <= m.min_spel_x || qx >= m.max_spel_x );
        if( m.satd || do_qpel )
        {
            const int dx = !do_qpel ? 0 : slot == 2 ? -1 : slot == 3 ? 1 : 0;
            const int dy = !do_qpel ? 0 : slot == 0 ? -1 : slot == 1 ? 1 : 0;
            int c = la_cost_qpel( m, qx + dx, qy + dy, true );
            if( m.satd ) qcost = __shfl_sync( 0xffffffffu, c, 16 );
            if( do_qpel )
            {
                int key = slot < 4 ? ( c << 2 ) + slot : 0x7fffffff;
                key = warp_min( key );
                if( ( key >> 2 ) < qcost )
                {
                    int w = key & 3;
                    qcost = key >> 2;
                    qx += w == 2 ? -1 : w == 3 ? 1 : 0;
                    qy += w == 0 ? -1 : w == 1 ? 1 : 0;
                }
            }
        }
        mvx = qx; mvy = qy; cost = qcost;
    }
    LA_TICK( m.prof[2], t_prof );
    out_mvx = mvx; out_mvy = mvy; out_cost = cost;
}

} // namespace x264cu
