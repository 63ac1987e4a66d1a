// Generic warp-per-search block matcher: x264_me_search_ref + refine_subpel (encoder/me.c:182-992) for every luma
// partition size, DIA / HEX / UMH, every sub-pel level (subpel_iterations, me.c:38-50), optional weighted reference and
// half-pel early-termination threshold; no chroma ME, no ESA/TESA.
//
// One warp runs one search with the reference's exact control flow (warp-uniform); every step's candidates are evaluated
// in parallel: a WxH block is covered by L = (W/4)*(H/4) lanes (one 4x4 each), so S = 32/L candidates per round.
// The reference's sequential "strictly smaller wins, first wins ties" chains (COPY1/3/4_IF_LT, me.c) are reproduced by a
// warp-wide minimum over packed (cost<<8 | order) keys, applied round by round.
#pragma once
#include "lookahead_dev.cuh"
#include <stdio.h>

namespace x264cu {

// i-th signed 4-bit entry of a packed table (keeps the small offset tables in immediates instead of local memory)
__device__ __forceinline__ int nib( unsigned long long tab, int i ) { return (int)( (long long)( tab << ( 60 - 4*i ) ) >> 60 ); }

struct MeShared                       // per-launch constants
{
    const uint8_t *fenc; int fenc_stride;
    const uint8_t *fref[4]; const uint8_t *fref_w; int stride;
    const uint16_t *cost_mv;          // centred table in global memory
    int me_method, subpel_refine, me_range;
    int satd;                         // mbcmp is SATD (encoder subme > 1)
    LaWeight w;
};

template <int BW, int BH>
struct MeWarp
{
    static constexpr int LX = BW / 4, LY = BH / 4, L = LX * LY, S = 32 / L;
    // per-lane
    uint32_t fenc[4];
    const uint8_t *fref0, *fref1, *fref2, *fref3, *fref_w;
    int stride;
    const uint16_t *cost_mv;
    int mvpx, mvpy;
    int x_min, y_min, x_max, y_max;
    int min_spel_x, min_spel_y, max_spel_x, max_spel_y;
    LaWeight w;
    bool satd;
    int slot;
    // search state (uniform)
    int bmx, bmy, bcost;

    __device__ __forceinline__ int group_sum( int v ) const
    {
#pragma unroll
        for( int m = 1; m < L; m <<= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, m );
        return v;
    }
    __device__ __forceinline__ bool in_range( int x, int y ) const { return x >= x_min && x <= x_max && y >= y_min && y <= y_max; }
    __device__ __forceinline__ int bits_fpel( int mx, int my ) const { return __ldg( cost_mv + ( mx*4 - mvpx ) ) + __ldg( cost_mv + ( my*4 - mvpy ) ); }
    __device__ __forceinline__ int sad_fpel( int mx, int my ) const
    {
        uint32_t b[4];
        const uint8_t *s = fref_w + my * stride + mx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = ldg4u( s + r * stride );
        return group_sum( sad4x4( fenc, b ) );
    }
    __device__ __forceinline__ int cost_fpel( int mx, int my ) const { return sad_fpel( mx, my ) + bits_fpel( mx, my ); }
    __device__ __forceinline__ int cost_qpel( int mx, int my, bool use_mbcmp ) const
    {
        uint32_t b[4];
        qpel4x4_p( fref0, fref1, fref2, fref3, stride, w, mx, my, b );
        int d = ( use_mbcmp && satd ) ? satd4x4( fenc, b ) : sad4x4( fenc, b );
        return group_sum( d ) + __ldg( cost_mv + ( mx - mvpx ) ) + __ldg( cost_mv + ( my - mvpy ) );
    }

    // Ordered candidate list, full-pel: candidate i = (ox + dx(i), oy + dy(i)); `checked` = skip out-of-range ones
    // (CHECK_MVRANGE'd COST_MV) -- COST_MV_X4 / COST_MV in order, me.c:63-118
    template <typename FX, typename FY, typename FV>
    __device__ __forceinline__ void try_list_v( int n, int ox, int oy, FX dx, FY dy, FV valid )
    {
        for( int base = 0; base < n; base += S )
        {
            const int i = base + slot;
            const int cx = ox + dx( min( i, n - 1 ) ), cy = oy + dy( min( i, n - 1 ) );
            const bool ok = i < n && valid( cx, cy );
            // lanes of skipped candidates still take part in the shuffles: probe a safe position
            const int c = cost_fpel( ok ? cx : bmx, ok ? cy : bmy );
            int key = ok ? ( c << 8 ) | i : 0x7fffffff;
            key = warp_min( key );
#ifdef ME_DEBUG
            if( blockIdx.x == 0 && threadIdx.x < 32 && ( threadIdx.x % L ) == 0 )
                printf( "  try lane %d i %d n %d cand %d %d ok %d c %d key %x bcost %d\n", threadIdx.x, i, n, cx, cy, (int)ok, c, key, bcost );
#endif
            if( key != 0x7fffffff && ( key >> 8 ) < bcost )
            {
                const int w = key & 255;
                bcost = key >> 8;
                bmx = ox + dx( w ); bmy = oy + dy( w );
            }
        }
    }

    template <typename FX, typename FY>
    __device__ __forceinline__ void try_list( int n, int ox, int oy, FX dx, FY dy, bool checked )
    {
        try_list_v( n, ox, oy, dx, dy, [=]( int cx, int cy ) { return !checked || in_range( cx, cy ); } );
    }

    __device__ void dia1( int ox, int oy )                              // DIA1_ITER, me.c:143-150
    {
        try_list( 4, ox, oy, []( int i ) { return i == 2 ? -1 : i == 3 ? 1 : 0; }, []( int i ) { return i == 0 ? -1 : i == 1 ? 1 : 0; }, false );
    }

    // CROSS, me.c:152-176: +i, -i for i = start, start+2, ... < max on x, then on y.  Each candidate is checked only
    // against the limit it moves towards (the unrolled x4 part of the macro runs only where those checks hold anyway).
    __device__ void cross( int ox, int oy, int start, int xmax, int ymax )
    {
        const int xlo = x_min, xhi = x_max, ylo = y_min, yhi = y_max;
        const int nx = xmax > start ? ( ( xmax - start + 1 ) / 2 ) * 2 : 0;
        try_list_v( nx, ox, oy, [=]( int i ) { int d = start + ( i >> 1 ) * 2; return ( i & 1 ) ? -d : d; }, []( int ) { return 0; },
                    [=]( int cx, int ) { return cx > ox ? cx <= xhi : cx >= xlo; } );
        const int ny = ymax > start ? ( ( ymax - start + 1 ) / 2 ) * 2 : 0;
        try_list_v( ny, ox, oy, []( int ) { return 0; }, [=]( int i ) { int d = start + ( i >> 1 ) * 2; return ( i & 1 ) ? -d : d; },
                    [=]( int, int cy ) { return cy > oy ? cy <= yhi : cy >= ylo; } );
    }

    __device__ void hex_refine( int me_range )                           // me.c:344-420
    {
        auto hex2x = []( int k ) { return (int32_t)( 0xEF121FEFu << ( 28 - 4*k ) ) >> 28; };
        auto hex2y = []( int k ) { return (int32_t)( 0x0EE0220Eu << ( 28 - 4*k ) ) >> 28; };
        auto ring = [&]( int n, int first, int tag0, int cur ) {
            // candidates hex2[first + j], tags tag0 + j; returns packed (cost<<3)+tag minimum against `cur`
            int best = cur;
            for( int base = 0; base < n; base += S )
            {
                const int j = base + slot;
                const int k = min( first + min( j, n - 1 ), 7 );
                const int c = cost_fpel( bmx + hex2x( k ), bmy + hex2y( k ) );
                int key = j < n ? ( c << 3 ) + tag0 + j : 0x7fffffff;
                best = min( best, warp_min( key ) );
            }
            return best;
        };
        int key = ring( 6, 1, 2, bcost << 3 );
        if( key & 7 )
        {
            int dir = ( key & 7 ) - 2;
            bmx += hex2x( dir + 1 ); bmy += hex2y( dir + 1 );
            for( int i = ( me_range >> 1 ) - 1; i > 0 && in_range( bmx, bmy ); i-- )
            {
                key = ring( 3, dir, 1, key & ~7 );
                if( !( key & 7 ) )
                    break;
                dir += ( key & 7 ) - 2;
                dir = dir < 0 ? 5 : dir > 5 ? dir - 6 : dir;
                bmx += hex2x( dir + 1 ); bmy += hex2y( dir + 1 );
            }
        }
        bcost = key >> 3;
        // square1[1..8]
        auto sqx = []( int t ) { return ( t == 3 || t == 5 || t == 6 ) ? -1 : ( t == 4 || t == 7 || t == 8 ) ? 1 : 0; };
        auto sqy = []( int t ) { return ( t == 1 || t == 5 || t == 7 ) ? -1 : ( t == 2 || t == 6 || t == 8 ) ? 1 : 0; };
        int best = bcost << 4;
        for( int base = 0; base < 8; base += S )
        {
            const int t = base + slot + 1;
            const int tt = min( t, 8 );
            const int c = cost_fpel( bmx + sqx( tt ), bmy + sqy( tt ) );
            int k2 = t <= 8 ? ( c << 4 ) + t : 0x7fffffff;
            best = min( best, warp_min( k2 ) );
        }
        const int w = best & 15;
        bmx += sqx( w ); bmy += sqy( w );
        bcost = best >> 4;
    }
};

// mvc: up to 8 candidate vectors; thresh_io: half-pel early-termination threshold (< 0 = none)
template <int BW, int BH>
__device__ void me_search_generic( const MeShared &g, int i_pixel, uint32_t fenc_off, uint32_t ref_off, int mvpx, int mvpy,
                                   const int16_t *mvc, int i_mvc, const int16_t *limits /* min_x,min_y,max_x,max_y spel */,
                                   int &thresh_io, int lane, int &out_mvx, int &out_mvy, int &out_cost, int &out_cost_mv )
{
    using M = MeWarp<BW, BH>;
    M m;
    const int gl = lane % M::L;
    const int sx = ( gl % M::LX ) * 4, sy = ( gl / M::LX ) * 4;
    m.slot = lane / M::L;
    m.stride = g.stride; m.cost_mv = g.cost_mv; m.w = g.w; m.satd = g.satd != 0;
    m.mvpx = mvpx; m.mvpy = mvpy;
    {
        const uint8_t *f = g.fenc + fenc_off + sy * g.fenc_stride + sx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) m.fenc[r] = ldg4u( f + r * g.fenc_stride );
        const int o = ref_off + sy * g.stride + sx;
        m.fref0 = g.fref[0] + o; m.fref1 = g.fref[1] + o; m.fref2 = g.fref[2] + o; m.fref3 = g.fref[3] + o;
        m.fref_w = g.fref_w + o;
    }
    m.min_spel_x = limits[0]; m.min_spel_y = limits[1]; m.max_spel_x = limits[2]; m.max_spel_y = limits[3];
    m.x_min = m.min_spel_x >> 2; m.y_min = m.min_spel_y >> 2; m.x_max = m.max_spel_x >> 2; m.y_max = m.max_spel_y >> 2;
    m.bcost = LA_COST_MAX;
    int me_range = g.me_range;
    const int subpel = g.subpel_refine;
    int bpred_cost = LA_COST_MAX, pmx, pmy;
    uint32_t pmv, bpred_mv = 0;

    // ---- predictor stage, me.c:216-318 ----
    // candidates kept as packed (x | y<<16) words in ONE local array (two dynamically indexed local arrays were seen
    // to alias in the generated code)
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;      // packed (x | y<<16), kept in registers
    int n = 0;
    auto cget = [&]( int k ) { return k == 0 ? c0 : k == 1 ? c1 : k == 2 ? c2 : k == 3 ? c3 : k == 4 ? c4 : k == 5 ? c5 : k == 6 ? c6 : c7; };
    auto cput = [&]( int k, uint32_t v ) {
        if( k == 0 ) c0 = v; else if( k == 1 ) c1 = v; else if( k == 2 ) c2 = v; else if( k == 3 ) c3 = v;
        else if( k == 4 ) c4 = v; else if( k == 5 ) c5 = v; else if( k == 6 ) c6 = v; else c7 = v; };
#define CX( k ) ( (int)(int16_t)( cget( k ) & 0xffff ) )
#define CY( k ) ( (int)(int16_t)( cget( k ) >> 16 ) )
    if( subpel >= 3 )
    {
        int bpx = clip3i( mvpx, m.x_min*4, m.x_max*4 ), bpy = clip3i( mvpy, m.y_min*4, m.y_max*4 );
        pmv = pack_mv( bpx, bpy );
        pmx = LA_FPEL( bpx ); pmy = LA_FPEL( bpy );
        for( int i = 0; i < 8; i++ )
        {
            if( i >= i_mvc ) break;
            int vx = mvc[2*i], vy = mvc[2*i+1];
            uint32_t mv = pack_mv( vx, vy );
            if( !mv || mv == pmv ) continue;
            cput( n, pack_mv( clip3i( vx, m.x_min*4, m.x_max*4 ), clip3i( vy, m.y_min*4, m.y_max*4 ) ) );
            n++;
        }
        int pmv_cost = m.cost_qpel( bpx, bpy, false );           // every slot computes it: uniform value
        bpred_cost = pmv_cost;
        if( n > 0 )
        {
            int best = bpred_cost << 4;
            for( int base = 0; base < n; base += M::S )
            {
                const int i = base + m.slot;
                const int ii = min( i, n - 1 );
                const int c = m.cost_qpel( CX( ii ), CY( ii ), false );
                int key = i < n ? ( c << 4 ) + i + 1 : 0x7fffffff;
                best = min( best, warp_min( key ) );
            }
            if( best & 15 ) { bpx = CX( ( best & 15 ) - 1 ); bpy = CY( ( best & 15 ) - 1 ); }
            bpred_cost = best >> 4;
        }
        m.bmx = LA_FPEL( bpx ); m.bmy = LA_FPEL( bpy );
        bpred_mv = pack_mv( bpx, bpy );
        if( bpred_mv & 0x00030003 )
        {
            int c = m.cost_fpel( m.bmx, m.bmy );
            if( c < m.bcost ) m.bcost = c;
        }
        else
            m.bcost = bpred_cost;
        if( pmv )
        {
            if( m.bmx | m.bmy )
            {
                int c = m.cost_fpel( 0, 0 );
                if( c < m.bcost ) { m.bcost = c; m.bmx = 0; m.bmy = 0; }
            }
        }
        else if( pmv_cost < m.bcost ) { m.bcost = pmv_cost; m.bmx = 0; m.bmy = 0; }
    }
    else
    {
        m.bmx = pmx = clip3i( LA_FPEL( mvpx ), m.x_min, m.x_max );
        m.bmy = pmy = clip3i( LA_FPEL( mvpy ), m.y_min, m.y_max );
        pmv = pack_mv( m.bmx, m.bmy );
        m.bcost = m.sad_fpel( m.bmx, m.bmy );                     // no mv cost on the rounded predictor (me.c:283-291)
        for( int i = 0; i < 8; i++ )
        {
            if( i >= i_mvc ) break;
            int rx = ( mvc[2*i] + 2 ) >> 2, ry = ( mvc[2*i+1] + 2 ) >> 2;
            uint32_t mv = pack_mv( rx, ry );
            if( !mv || mv == pmv ) continue;
            cput( n, pack_mv( clip3i( rx, m.x_min, m.x_max ), clip3i( ry, m.y_min, m.y_max ) ) );
            n++;
        }
        if( n > 0 )
        {
            int best = m.bcost << 4;
            for( int base = 0; base < n; base += M::S )
            {
                const int i = base + m.slot;
                const int ii = min( i, n - 1 );
                const int c = m.cost_fpel( CX( ii ), CY( ii ) );
                int key = i < n ? ( c << 4 ) + i + 1 : 0x7fffffff;
#ifdef ME_DEBUG
                if( blockIdx.x == 0 && threadIdx.x < 32 && ( threadIdx.x % M::L ) == 0 )
                    printf( "  pred lane %d i %d ii %d n %d cand %d %d c %d key %x bcost %d\n", threadIdx.x, i, ii, n, CX( ii ), CY( ii ), c, key, m.bcost );
#endif
                best = min( best, warp_min( key ) );
            }
            if( best & 15 ) { m.bmx = CX( ( best & 15 ) - 1 ); m.bmy = CY( ( best & 15 ) - 1 ); }
            m.bcost = best >> 4;
        }
        if( pmv )
        {
            int c = m.cost_fpel( 0, 0 );
            if( c < m.bcost ) { m.bcost = c; m.bmx = 0; m.bmy = 0; }
        }
    }

#ifdef ME_DEBUG
    if( blockIdx.x == 0 && threadIdx.x == 0 )
        printf( "dbg method %d subpel %d range %d i_mvc %d mvp %d %d lim %d %d %d %d n %d start %d %d cost %d pmv %x L %d S %d\n", g.me_method, subpel, me_range, i_mvc,
                mvpx, mvpy, m.x_min, m.y_min, m.x_max, m.y_max, n, m.bmx, m.bmy, m.bcost, pmv, M::L, M::S );
#endif
    // ---- integer search ----
    if( g.me_method == X264CU_ME_DIA )
    {
        int i = me_range;
        do
        {
            const int ox = m.bmx, oy = m.bmy;
            m.dia1( ox, oy );
            if( m.bmx == ox && m.bmy == oy )
                break;
        } while( --i && m.in_range( m.bmx, m.bmy ) );
    }
    else if( g.me_method == X264CU_ME_HEX )
        m.hex_refine( me_range );
    else if( g.me_method == X264CU_ME_ESA )
    {   // me.c:618-771, the "just ADS and SAD" branch.  The reference's ADS prefilter (pixf.ads over the integral image) only
        // drops positions whose lower bound |sum(fenc) - sum(ref)| + mv cost already reaches the best cost: sum|a-b| >= |sum a -
        // sum b|, so none of them could be strictly better.  The result is therefore the first strictly smaller cost in raster
        // order over the window, which the warp computes by brute force, S positions per round, row after row (32/L SADs per
        // round: no integral images, no prefilter).  The window's width is rounded up to a multiple of 4 as in the reference
        // (up to 3 positions past mv_x_max, never range-checked).
        const int min_x = max( m.bmx - me_range, m.x_min ), min_y = max( m.bmy - me_range, m.y_min );
        const int max_x = min( m.bmx + me_range, m.x_max ), max_y = min( m.bmy + me_range, m.y_max );
        const int width = ( max_x - min_x + 3 ) & ~3;
        for( int my = min_y; my <= max_y; my++ )
            m.try_list_v( width, min_x, my, []( int i ) { return i; }, []( int ) { return 0; }, []( int, int ) { return true; } );
    }
    else
    {   // UMH, me.c:422-616
        const int shift = i_pixel == 0 ? 0 : i_pixel <= 2 ? 1 : i_pixel == 3 ? 2 : i_pixel <= 5 ? 3 : 4;   // pixel_size_shift
        int ucost1 = m.bcost, ucost2, cross_start = 1;
        bool done = false;
        m.dia1( pmx, pmy );
        if( pmx | pmy )
            m.dia1( 0, 0 );
        if( i_pixel == X264CU_PIXEL_4x4 )
            m.hex_refine( me_range );
        else
        {
            ucost2 = m.bcost;
            if( ( m.bmx | m.bmy ) && ( ( m.bmx - pmx ) | ( m.bmy - pmy ) ) )
                m.dia1( m.bmx, m.bmy );
            if( m.bcost == ucost2 )
                cross_start = 3;
            int omx = m.bmx, omy = m.bmy;
            if( m.bcost == ucost2 && m.bcost < ( 2000 >> shift ) )
            {
                // octagon: (0,-2) (-1,-1) (1,-1) (-2,0) (2,0) (-1,1) (1,1) (0,2)
                m.try_list( 8, omx, omy, []( int i ) { return nib( 0x01F2E1F0ull, i ); },
                            []( int i ) { return nib( 0x21100FFEull, i ); }, false );
                if( m.bcost == ucost1 && m.bcost < ( 500 >> shift ) )
                    done = true;
                else if( m.bcost == ucost2 )
                {
                    const int range = ( me_range >> 1 ) | 1;
                    m.cross( omx, omy, 3, range, range );
                    // (-1,-2) (1,-2) (-2,-1) (2,-1) (-2,1) (2,1) (-1,2) (1,2)
                    m.try_list( 8, omx, omy, []( int i ) { return nib( 0x1F2E2E1Full, i ); },
                                []( int i ) { return nib( 0x2211FFEEull, i ); }, false );
                    if( m.bcost == ucost2 )
                        done = true;
                    cross_start = range + 2;
                }
            }
            if( !done )
            {
                if( i_mvc )
                {   // adaptive search range, me.c:469-519
                    int mvd, denom = 1;
                    if( i_mvc == 1 )
                        mvd = i_pixel == X264CU_PIXEL_16x16 ? 25 : abs( mvpx - mvc[0] ) + abs( mvpy - mvc[1] );
                    else
                    {
                        denom = i_mvc - 1;
                        mvd = 0;
                        if( i_pixel != X264CU_PIXEL_16x16 )
                        {
                            mvd = abs( mvpx - mvc[0] ) + abs( mvpy - mvc[1] );
                            denom++;
                        }
                        for( int i = 0; i < i_mvc - 1; i++ )
                            mvd += abs( mvc[2*i] - mvc[2*i+2] ) + abs( mvc[2*i+1] - mvc[2*i+3] );
                    }
                    const int sad_ctx = m.bcost < ( 1000 >> shift ) ? 0 : m.bcost < ( 2000 >> shift ) ? 1 : m.bcost < ( 4000 >> shift ) ? 2 : 3;
                    const int mvd_ctx = mvd < 10*denom ? 0 : mvd < 20*denom ? 1 : mvd < 40*denom ? 2 : 3;
                    // range_mul[mvd_ctx][sad_ctx] = {3,3,4,4},{3,4,4,4},{4,4,4,5},{4,4,5,6}
                    const uint32_t rm = mvd_ctx == 0 ? 0x4433u : mvd_ctx == 1 ? 0x4443u : mvd_ctx == 2 ? 0x5444u : 0x6544u;
                    me_range = me_range * (int)( ( rm >> ( 4*sad_ctx ) ) & 15 ) >> 2;
                }
                m.cross( omx, omy, cross_start, me_range, me_range >> 1 );
                m.try_list( 4, omx, omy, []( int i ) { return i < 2 ? -2 : 2; }, []( int i ) { return ( i & 1 ) ? 2 : -2; }, false );
                // hexagon grid: 16 points scaled by i = 1 .. range/4, centre fixed at the best so far
                omx = m.bmx; omy = m.bmy;
                int i = 1;
                do
                {
                    const int sc = i;
                    m.try_list( 16, omx, omy,
                                [=]( int j ) { return nib( 0x2E4C4C4C4C4C2E00ull, j ) * sc; },
                                [=]( int j ) { return nib( 0x33221100FFEEDD4Cull, j ) * sc; }, true );
                } while( ++i <= me_range >> 2 );
                if( m.bmy <= m.y_max && m.bmy >= m.y_min && m.bmx <= m.x_max && m.bmx >= m.x_min )
                    m.hex_refine( me_range );
            }
        }
    }

#ifdef ME_DEBUG
    if( blockIdx.x == 0 && threadIdx.x == 0 )
        printf( "dbg after int search %d %d cost %d\n", m.bmx, m.bmy, m.bcost );
#endif
    // ---- back to quarter-pel, me.c:774-789 ----
    int qx, qy, qcost;
    if( subpel < 3 )
    {
        qcost = m.bcost;
        if( pack_mv( m.bmx, m.bmy ) == pmv )
            qcost += m.bits_fpel( m.bmx, m.bmy );
        qx = m.bmx*4; qy = m.bmy*4;
    }
    else if( bpred_cost < m.bcost ) { qx = (int16_t)( bpred_mv & 0xFFFF ); qy = (int16_t)( bpred_mv >> 16 ); qcost = bpred_cost; }
    else { qx = m.bmx*4; qy = m.bmy*4; qcost = m.bcost; }

    // ---- refine_subpel, me.c:865-992 ----
    if( subpel >= 2 )
    {
        // subpel_iterations[subme][2..3]: me_hpel, me_qpel
        const int hpel_iters = subpel < 8 ? ( subpel >= 6 ? 2 : 1 ) : 4;
        const int qpel_iters = subpel < 4 ? 0 : subpel == 4 ? 1 : subpel < 8 ? 2 : 10;
        auto diamond = [&]( int step, bool use_mbcmp, int skip_dir ) {
            // candidates (0,-s) (0,+s) (-s,0) (+s,0); returns the winning direction or -1; updates qx,qy,qcost
            int best = 0x7fffffff;
            for( int base = 0; base < 4; base += M::S )
            {
                const int dir = base + m.slot;
                const int dd = min( dir, 3 );
                const int dx = dd == 2 ? -step : dd == 3 ? step : 0, dy = dd == 0 ? -step : dd == 1 ? step : 0;
                const bool ok = dir < 4 && dir != skip_dir;
                const int c = m.cost_qpel( qx + dx, qy + dy, use_mbcmp );
                int key = ok ? ( c << 2 ) + dir : 0x7fffffff;
                best = min( best, warp_min( key ) );
            }
            if( best != 0x7fffffff && ( best >> 2 ) < qcost )
            {
                const int dir = best & 3;
                qcost = best >> 2;
                qx += dir == 2 ? -step : dir == 3 ? step : 0;
                qy += dir == 0 ? -step : dir == 1 ? step : 0;
                return dir;
            }
            return -1;
        };
        if( hpel_iters )
        {
            if( subpel < 3 )
            {
                int px = clip3i( mvpx, m.min_spel_x + 2, m.max_spel_x - 2 ), py = clip3i( mvpy, m.min_spel_y + 2, m.max_spel_y - 2 );
                if( ( px - qx ) | ( py - qy ) )
                {
                    int c = m.cost_qpel( px, py, false );
                    if( c < qcost ) { qcost = c; qx = px; qy = py; }
                }
            }
            for( int i = hpel_iters; i > 0; i-- )
                if( diamond( 2, false, -1 ) < 0 )
                    break;
        }
        if( m.satd )
            qcost = m.cost_qpel( qx, qy, true );
        bool early = false;
        if( thresh_io >= 0 )
        {
            if( ( qcost * 7 ) >> 3 > thresh_io ) early = true;
            else if( qcost < thresh_io ) thresh_io = qcost;
        }
        if( !early )
        {
            if( subpel != 1 )
            {
                int bdir = -1;
                for( int i = qpel_iters; i > 0; i-- )
                {
                    if( qy <= m.min_spel_y || qy >= m.max_spel_y || qx <= m.min_spel_x || qx >= m.max_spel_x )
                        break;
                    const int odir = bdir;
                    const int d = diamond( 1, true, odir >= 0 ? ( odir ^ 1 ) : -1 );
                    if( d < 0 )
                        break;
                    bdir = d;
                }
            }
        }
    }
    out_mvx = qx; out_mvy = qy; out_cost = qcost;
    out_cost_mv = __ldg( g.cost_mv + ( qx - mvpx ) ) + __ldg( g.cost_mv + ( qy - mvpy ) );
}

#undef CX
#undef CY

} // namespace x264cu
