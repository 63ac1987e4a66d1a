// Generic warp-per-search block matcher: x264_me_search_ref + refine_subpel (encoder/me.c:182-992) for every luma
// partition size, DIA / HEX / UMH / ESA / TESA, every sub-pel level (subpel_iterations, me.c:38-50), optional weighted reference,
// half-pel early-termination threshold and chroma ME (the chroma branch of COST_MV_SATD, me.c:826-857, with mc_chroma,
// common/mc.c:251-283; 4:2:0).  Also x264_me_refine_bidir_satd (me.c:1027-1183).
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

// pair j of x264_me_refine_bidir_satd's table (me.c:1063-1074) as base-3 digits, offset = digit - 1, order (m0x, m0y, m1x, m1y)
__constant__ uint8_t c_bidir_pairs[33] =
    { 40, 67, 13, 49, 31, 43, 37, 41, 39, 76, 4, 52, 28, 44, 36, 68, 12, 70, 10, 50, 30, 58, 22, 46, 34, 42, 38, 14, 66, 64, 16, 48, 32 };
__device__ __forceinline__ int bidir_code( int j ) { return c_bidir_pairs[j]; }

struct MeShared                       // per-launch constants
{
    const uint8_t *fenc; int fenc_stride;
    const uint8_t *fref[4]; const uint8_t *fref_w; int stride;
    const uint16_t *cost_mv;          // centred table in global memory
    int me_method, subpel_refine, me_range;
    int satd;                         // mbcmp is SATD (encoder subme > 1)
    int fpel_border;                  // mv_limit_fpel = (mv_min_spel >> 2) + border .. (mv_max_spel >> 2) - border (analyse.c:333-349)
    int fpel_satd;                    // fpelcmp is SATD too: TESA with subme > 1 (encoder.c:1409-1427)
    LaWeight w;
    uint2 *tesa_list; int tesa_cap;   // TESA: per-warp candidate lists (sad, packed mv), tesa_cap entries each
    // chroma ME (h->mb.b_chroma_me): NV12 planes, pixel (0,0) of each; the weights of the two chroma planes (m->weight[1..2])
    int chroma;
    const uint8_t *fenc_uv; int fenc_uv_stride;
    const uint8_t *fref_uv; int ref_uv_stride;
    LaWeight wc[2];
};

// the two cost functions every search step ends in: inlined at every call site, or one out-of-line copy per partition size
#ifndef ME_COST_INLINE
#define ME_COST_INLINE __forceinline__
#endif

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
    bool satd, fpel_satd;
    int slot;
    int fsum, gl;                     // ADS: pixel sum of this lane's 4x4 of fenc; lane index within the candidate group
    // chroma ME: the first L/4 lanes of a candidate group take the 4x4 blocks of U, the next L/4 those of V
    bool chroma, cact;                // chroma ME on for this search (uniform) / this lane holds a chroma block
    uint32_t cfenc[4];                // the lane's 4x4 of the source's chroma plane
    const uint8_t *cref;              // its corner in the reference's NV12 plane (component offset included)
    int cstride;
    LaWeight cw;
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
    __device__ ME_COST_INLINE int sad_fpel( int mx, int my ) const
    {
        uint32_t b[4];
        const uint8_t *s = fref_w + my * stride + mx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = ldg4u( s + r * stride );
        return group_sum( fpel_satd ? satd4x4( fenc, b ) : sad4x4( fenc, b ) );        // fpelcmp
    }
    __device__ __forceinline__ int cost_fpel( int mx, int my ) const { return sad_fpel( mx, my ) + bits_fpel( mx, my ); }
    // mc_chroma (common/mc.c:251-283) of this lane's 4x4 chroma block at the luma vector read in eighth-pels, the plane's
    // explicit weight, mbcmp against the source: this lane's share of COST_MV_SATD's chroma terms (me.c:843-855).  The reference
    // stops adding once the sum reaches the best cost; partial sums only grow, so what is accepted is the same with the whole sum.
    __device__ __noinline__ int chroma_cost( int mx, int my ) const
    {
        const int dx = mx & 7, dy = my & 7;
        const uint32_t w00 = ( 8 - dx ) * ( 8 - dy ), w01 = dx * ( 8 - dy ), w10 = ( 8 - dx ) * dy, w11 = dx * dy;
        const uint8_t *s = cref + ( my >> 3 ) * cstride + ( mx >> 3 ) * 2;
        // a row of five samples as two words of 16-bit fields per tap position: samples (0,2) / (1,3) for the left tap, (1,3) / (2,4)
        // for the right one -- the weighted sum of four samples is at most 64 * 255 + 32, so both fields of a word are summed at once
        uint32_t e0, o0, e1;                       // fields (s0,s2), (s1,s3), (s2,s4) of the row above the one being produced
        {
            const uint32_t a = ldg4u( s ), b = ldg4u( s + 4 ), c = ldg4u( s + 8 ) & 0xff;
            const uint32_t lo = __byte_perm( a, b, 0x6420 );
            e0 = __byte_perm( lo, 0, 0x4240 ); o0 = __byte_perm( lo, 0, 0x4341 ); e1 = __byte_perm( lo, c, 0x5452 );
        }
        uint32_t p[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            s += cstride;
            const uint32_t a = ldg4u( s ), b = ldg4u( s + 4 ), c = ldg4u( s + 8 ) & 0xff;
            const uint32_t lo = __byte_perm( a, b, 0x6420 );
            const uint32_t ne0 = __byte_perm( lo, 0, 0x4240 ), no0 = __byte_perm( lo, 0, 0x4341 ), ne1 = __byte_perm( lo, c, 0x5452 );
            // pixels 0 and 2: left taps e0 / ne0, right taps o0 / no0; pixels 1 and 3: left o0 / no0, right e1 / ne1
            const uint32_t p02 = ( e0 * w00 + o0 * w01 + ne0 * w10 + no0 * w11 + 0x00200020u ) >> 6;
            const uint32_t p13 = ( o0 * w00 + e1 * w01 + no0 * w10 + ne1 * w11 + 0x00200020u ) >> 6;
            const uint32_t out = __byte_perm( p02, p13, 0x6240 );
            p[r] = cw.enabled ? weight4( out, cw ) : out;
            e0 = ne0; o0 = no0; e1 = ne1;
        }
        return satd ? satd4x4( cfenc, p ) : sad4x4( cfenc, p );
    }
    __device__ ME_COST_INLINE int cost_qpel( int mx, int my, bool use_mbcmp ) const
    {
        uint32_t b[4];
        qpel4x4_p( fref0, fref1, fref2, fref3, stride, w, mx, my, b );
        int d = ( use_mbcmp ? satd : fpel_satd ) ? satd4x4( fenc, b ) : sad4x4( fenc, b );
        if( BW >= 8 && BH >= 8 && chroma && use_mbcmp )
        {
            const int dc = chroma_cost( mx, my );          // every lane runs it (lanes without a block read a valid address)
            d += cact ? dc : 0;
        }
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

    // ---- exhaustive searches, me.c:618-771 ----
    // pixf.sad (always SAD) of the block at a full-pel position of the weighted plane, and pixf.ads' value without the mv cost
    // (pixel.c:759-803): sum over the block's 8x8 (4x4 for partitions below 8x8) sub-blocks of |dc(fenc) - dc(ref)|, the
    // reference's dc taken from the UNWEIGHTED plane like the integral image it is read from (mc.c:748-783)
    __device__ __forceinline__ void sad_ads( int mx, int my, int &sad, int &ads ) const
    {
        uint32_t b[4];
        const int o = my * stride + mx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = ldg4u( fref_w + o + r * stride );
        sad = group_sum( sad4x4( fenc, b ) );
        if( fref_w != fref0 )
        {
#pragma unroll
            for( int r = 0; r < 4; r++ ) b[r] = ldg4u( fref0 + o + r * stride );
        }
        int d = fsum;
#pragma unroll
        for( int r = 0; r < 4; r++ ) d -= __dp4a( b[r], 0x01010101u, 0u );
        if( BW >= 8 && BH >= 8 )
        {
            d += __shfl_xor_sync( 0xffffffffu, d, 1 );
            d += __shfl_xor_sync( 0xffffffffu, d, LX );
            d = ( ( gl % LX ) | ( gl / LX ) ) & 1 ? 0 : abs( d );
        }
        else
            d = abs( d );
        ads = group_sum( d );
    }

    // ESA, me.c:751-768: per row the positions whose ads + x mv cost stay below the best cost (less the row's y cost) are
    // measured in ascending x.  Without a weighted reference ads is a lower bound of the SAD, the prefilter cannot drop a
    // winner and it is skipped; with one (the sums come from the unweighted plane) it decides and is applied.
    __device__ void esa( int me_range )
    {
        const int min_x = max( bmx - me_range, x_min ), min_y = max( bmy - me_range, y_min );
        const int max_x = min( bmx + me_range, x_max ), max_y = min( bmy + me_range, y_max );
        const int width = ( max_x - min_x + 3 ) & ~3;      // rounded up as in the reference: up to 3 positions past mv_x_max
        const bool filter = fref_w != fref0;
        for( int my = min_y; my <= max_y; my++ )
        {
            const int ycost = __ldg( cost_mv + ( my*4 - mvpy ) );
            if( bcost <= ycost )
                continue;
            const int thresh = bcost - ycost;
            for( int base = 0; base < width; base += S )
            {
                const int i = base + slot;
                const int cx = min_x + min( i, width - 1 );
                const int xcost = __ldg( cost_mv + ( cx*4 - mvpx ) );
                int sad, ads = 0;
                if( filter ) sad_ads( cx, my, sad, ads );
                else sad = sad_fpel( cx, my );
                const bool ok = i < width && ads + xcost < thresh;
                int key = ok ? ( ( sad + xcost + ycost ) << 8 ) | slot : 0x7fffffff;
                key = warp_min( key );
                if( key != 0x7fffffff && ( key >> 8 ) < bcost )
                {
                    bcost = key >> 8;
                    bmx = min_x + base + ( key & 255 ); bmy = my;
                }
            }
        }
    }

    // TESA, me.c:656-747: ADS threshold (17/16 of the best SAD), SAD threshold (sad_thresh/8 of the running best), the list of
    // survivors thinned to me_range/2 entries, then fpelcmp (SATD) on those.  `list` = this warp's scratch in global memory.
    __device__ void tesa( int me_range, uint2 *list )
    {
        const int lane = slot * L + gl;
        const int min_x = max( bmx - me_range, x_min ), min_y = max( bmy - me_range, y_min );
        const int max_x = min( bmx + me_range, x_max ), max_y = min( bmy + me_range, y_max );
        const int width = ( max_x - min_x + 3 ) & ~3;
        int n = 0;
        int sad_thresh = me_range <= 16 ? 10 : me_range <= 24 ? 11 : 12;
        int bsad, unused;
        sad_ads( bmx, bmy, bsad, unused );
        bsad += bits_fpel( bmx, bmy );
        for( int my = min_y; my <= max_y; my++ )
        {
            const int ycost = __ldg( cost_mv + ( my*4 - mvpy ) );
            if( bsad <= ycost )
                continue;
            bsad -= ycost;
            const int athresh = bsad * 17 >> 4;
            for( int base = 0; base < width; base += S )
            {
                const int i = min( base + slot, width - 1 );
                const int cx = min_x + i;
                int sad, ads;
                sad_ads( cx, my, sad, ads );
                const int pass = base + slot < width && ads + __ldg( cost_mv + ( cx*4 - mvpx ) ) < athresh;
                // the listing step reads the x mv cost at the position's index within the row (cost_fpel_mvx[xs[i]], me.c:683, :699)
                const int v = sad + __ldg( cost_mv + ( i*4 - mvpx ) );
#pragma unroll 1
                for( int k = 0; k < S; k++ )
                {
                    const int pk = __shfl_sync( 0xffffffffu, pass, k * L ), vk = __shfl_sync( 0xffffffffu, v, k * L );
                    if( pk && vk < ( bsad * sad_thresh >> 3 ) )
                    {
                        bsad = min( bsad, vk );
                        if( lane == 0 ) list[n] = make_uint2( (uint32_t)( vk + ycost ), pack_mv( min_x + base + k, my ) );
                        n++;
                    }
                }
            }
            bsad += ycost;
        }
        __syncwarp();
        const int limit = me_range >> 1;
        sad_thresh = bsad * sad_thresh >> 3;
        while( n > limit*2 && sad_thresh > bsad )
        {   // halve the admitted range; keep, in order, what is still inside it
            sad_thresh = ( sad_thresh + bsad ) >> 1;
            int k = 0;
            for( int base = 0; base < n; base += 32 )
            {
                const int i = base + lane;
                uint2 e = make_uint2( 0, 0 );
                if( i < n ) e = list[i];
                const bool keep = i < n && (int)e.x <= sad_thresh;
                const uint32_t mask = __ballot_sync( 0xffffffffu, keep );
                __syncwarp();
                if( keep ) list[k + __popc( mask & ( ( 1u << lane ) - 1 ) )] = e;
                k += __popc( mask );
                __syncwarp();
            }
            n = k;
        }
        while( n > limit )
        {   // drop the first worst entry, the last one takes its place
            int msad = -1, midx = 0x7fffffff;
            for( int i = lane; i < n; i += 32 )
            {
                const int sd = (int)list[i].x;
                if( sd > msad ) { msad = sd; midx = i; }
            }
            const int wsad = __reduce_max_sync( 0xffffffffu, msad );
            const int bi = __reduce_min_sync( 0xffffffffu, msad == wsad ? midx : 0x7fffffff );
            n--;
            if( lane == 0 ) list[bi] = list[n];
            __syncwarp();
        }
        for( int base = 0; base < n; base += S )
        {
            const int i = base + slot;
            const uint2 e = list[min( i, n - 1 )];
            const int cx = (int16_t)( e.y & 0xffff ), cy = (int16_t)( e.y >> 16 );
            const int c = cost_fpel( cx, cy );
            int key = i < n ? ( c << 8 ) | slot : 0x7fffffff;
            key = warp_min( key );
            if( key != 0x7fffffff && ( key >> 8 ) < bcost )
            {
                const uint2 w = list[base + ( key & 255 )];
                bcost = key >> 8;
                bmx = (int16_t)( w.y & 0xffff ); bmy = (int16_t)( w.y >> 16 );
            }
        }
        __syncwarp();
    }
};

// per-lane state of one search: the lane's 4x4 of fenc, its corner in the reference planes, the window
template <int BW, int BH>
__device__ __forceinline__ void me_warp_setup( MeWarp<BW, BH> &m, const MeShared &g, uint32_t fenc_off, uint32_t ref_off, int mvpx, int mvpy,
                                               const int16_t *limits /* min_x,min_y,max_x,max_y spel */, int lane )
{
    using M = MeWarp<BW, BH>;
    const int gl = lane % M::L;
    m.gl = gl; m.fpel_satd = g.fpel_satd != 0;
    const int sx = ( gl % M::LX ) * 4, sy = ( gl / M::LX ) * 4;
    m.slot = lane / M::L;
    m.stride = g.stride; m.cost_mv = g.cost_mv; m.w = g.w; m.satd = g.satd != 0;
    m.mvpx = mvpx; m.mvpy = mvpy;
    {
        const uint8_t *f = g.fenc + fenc_off + sy * g.fenc_stride + sx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) m.fenc[r] = ldg4u( f + r * g.fenc_stride );
        m.fsum = 0;
#pragma unroll
        for( int r = 0; r < 4; r++ ) m.fsum = __dp4a( m.fenc[r], 0x01010101u, (uint32_t)m.fsum );
        const int o = ref_off + sy * g.stride + sx;
        m.fref0 = g.fref[0] + o; m.fref1 = g.fref[1] + o; m.fref2 = g.fref[2] + o; m.fref3 = g.fref[3] + o;
        m.fref_w = g.fref_w + o;
    }
    m.chroma = false; m.cact = false;
    if( BW >= 8 && BH >= 8 && g.chroma )
    {   // the block's chroma origin from its luma offset: (x, y) -> row y/2, byte 2*(x/2) of the interleaved plane
        constexpr int QL = M::L / 4, CLX = M::LX / 2;
        m.chroma = true;
        m.cact = gl < 2 * QL;
        const int comp = gl < QL ? 0 : 1, idx = m.cact ? gl - comp * QL : 0;
        const int cbx = ( idx % CLX ) * 4, cby = ( idx / CLX ) * 4;
        const int fy = fenc_off / g.fenc_stride, fx = fenc_off - fy * g.fenc_stride;
        const int ry = ref_off / g.stride, rx = ref_off - ry * g.stride;
        const uint8_t *f = g.fenc_uv + ( ( fy >> 1 ) + cby ) * g.fenc_uv_stride + ( fx & ~1 ) + 2 * cbx + comp;
#pragma unroll
        for( int r = 0; r < 4; r++ )
            m.cfenc[r] = __byte_perm( ldg4u( f + r * g.fenc_uv_stride ), ldg4u( f + r * g.fenc_uv_stride + 4 ), 0x6420 );
        m.cref = g.fref_uv + ( ( ry >> 1 ) + cby ) * g.ref_uv_stride + ( rx & ~1 ) + 2 * cbx + comp;
        m.cstride = g.ref_uv_stride;
        m.cw = g.wc[comp];
    }
    m.min_spel_x = limits[0]; m.min_spel_y = limits[1]; m.max_spel_x = limits[2]; m.max_spel_y = limits[3];
    m.x_min = ( m.min_spel_x >> 2 ) + g.fpel_border; m.y_min = ( m.min_spel_y >> 2 ) + g.fpel_border;
    m.x_max = ( m.max_spel_x >> 2 ) - g.fpel_border; m.y_max = ( m.max_spel_y >> 2 ) - g.fpel_border;
}

// refine_subpel( h, m, hpel_iters, qpel_iters, p_halfpel_thresh, b_refine_qpel ), me.c:865-992, from (qx, qy, qcost)
template <int BW, int BH>
__device__ __forceinline__ void me_refine_subpel( MeWarp<BW, BH> &m, int subpel, int hpel_iters, int qpel_iters, bool b_refine_qpel, int &thresh_io,
                                  int &qx, int &qy, int &qcost )
{
    using M = MeWarp<BW, BH>;
    const int mvpx = m.mvpx, mvpy = m.mvpy;
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
    if( !b_refine_qpel && ( ( m.satd && !m.fpel_satd ) || ( BW >= 8 && BH >= 8 && m.chroma ) ) )    // mbcmp != fpelcmp, or chroma to add:
        qcost = m.cost_qpel( qx, qy, true );                                                         // re-measure the winner, me.c:925-929
    if( thresh_io >= 0 )
    {
        if( ( qcost * 7 ) >> 3 > thresh_io ) return;
        else if( qcost < thresh_io ) thresh_io = qcost;
    }
    const bool inside = qy > m.min_spel_y && qy < m.max_spel_y && qx > m.min_spel_x && qx < m.max_spel_x;
    if( subpel != 1 )
    {
        int bdir = -1;
        for( int i = qpel_iters; i > 0; i-- )
        {
            if( qy <= m.min_spel_y || qy >= m.max_spel_y || qx <= m.min_spel_x || qx >= m.max_spel_x )
                break;
            const int odir = bdir;
            const int d = diamond( 1, true, !b_refine_qpel && odir >= 0 ? ( odir ^ 1 ) : -1 );     // never straight back, me.c:828
            if( d < 0 )
                break;
            bdir = d;
        }
    }
    else if( inside )
        diamond( 1, false, -1 );                                 // subme 1: one fpelcmp quarter-pel diamond, me.c:965-985
}

// mvc: up to 9 candidate vectors; thresh_io: half-pel early-termination threshold (< 0 = none)
// EXH: the exhaustive searches (ESA / TESA) are compiled in; the kernel of the other methods is built without them (their
// candidate-list code costs it 48 registers: 1.40 -> 1.61 ms per 4K picture of UMH merange-64 searches)
template <int BW, int BH, bool EXH>
__device__ void me_search_generic( const MeShared &g, int i_pixel, uint32_t fenc_off, uint32_t ref_off, int mvpx, int mvpy,
                                   const int16_t *mvc, int i_mvc, const int16_t *limits /* min_x,min_y,max_x,max_y spel */,
                                   int &thresh_io, int lane, int &out_mvx, int &out_mvy, int &out_cost, int &out_cost_mv, uint2 *tesa_list )
{
    using M = MeWarp<BW, BH>;
    M m;
    me_warp_setup<BW, BH>( m, g, fenc_off, ref_off, mvpx, mvpy, limits, lane );
    m.bcost = LA_COST_MAX;
    int me_range = g.me_range;
    const int subpel = g.subpel_refine;
    int bpred_cost = LA_COST_MAX, pmx, pmy;
    uint32_t pmv, bpred_mv = 0;

    // ---- predictor stage, me.c:216-318 ----
    // candidates kept as packed (x | y<<16) words in ONE local array (two dynamically indexed local arrays were seen
    // to alias in the generated code)
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0;      // packed (x | y<<16), kept in registers
    int n = 0;
    auto cget = [&]( int k ) { return k == 0 ? c0 : k == 1 ? c1 : k == 2 ? c2 : k == 3 ? c3 : k == 4 ? c4 : k == 5 ? c5 : k == 6 ? c6 : k == 7 ? c7 : c8; };
    auto cput = [&]( int k, uint32_t v ) {
        if( k == 0 ) c0 = v; else if( k == 1 ) c1 = v; else if( k == 2 ) c2 = v; else if( k == 3 ) c3 = v;
        else if( k == 4 ) c4 = v; else if( k == 5 ) c5 = v; else if( k == 6 ) c6 = v; else if( k == 7 ) c7 = v; else c8 = v; };
#define CX( k ) ( (int)(int16_t)( cget( k ) & 0xffff ) )
#define CY( k ) ( (int)(int16_t)( cget( k ) >> 16 ) )
    if( subpel >= 3 )
    {
        int bpx = clip3i( mvpx, m.x_min*4, m.x_max*4 ), bpy = clip3i( mvpy, m.y_min*4, m.y_max*4 );
        pmv = pack_mv( bpx, bpy );
        pmx = LA_FPEL( bpx ); pmy = LA_FPEL( bpy );
        for( int i = 0; i < 9; i++ )
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
        for( int i = 0; i < 9; i++ )
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
    else if( EXH && g.me_method == X264CU_ME_ESA )
        m.esa( me_range );
    else if( EXH && g.me_method == X264CU_ME_TESA )
        m.tesa( me_range, tesa_list );
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
        me_refine_subpel<BW, BH>( m, subpel, hpel_iters, qpel_iters, false, thresh_io, qx, qy, qcost );
    }
    out_mvx = qx; out_mvy = qy; out_cost = qcost;
    out_cost_mv = __ldg( g.cost_mv + ( qx - mvpx ) ) + __ldg( g.cost_mv + ( qy - mvpy ) );
}

#undef CX
#undef CY

// ---- x264_me_refine_bidir_satd, me.c:1027-1183 (rd = 0) ---------------------------------------------------------------------
// One warp refines one (list 0, list 1) vector pair.  A pass measures the pairs that differ from the current one by +-1
// quarter-pel in at most two of the four components -- 33 pairs (the reference's dia4d table, me.c:1063-1074, kept here as
// base-3 digits), S = 32/L of them per round -- skipping pairs already measured (the reference's 4096-bit map indexed by the low
// three bits of each component, in shared memory), cost = mbcmp( fenc, avg( ref0, ref1, weight ) ) + four mv costs; strictly
// smaller wins, first in table order on ties; up to 8 passes, until the centre stays.
__device__ __forceinline__ uint32_t avg4_bipred( uint32_t a, uint32_t b, int weight )       // pixel_avg_wxh / _weight_wxh, mc.c:49-75
{
    if( weight == 32 )
        return __vavgu4( a, b );
    uint32_t out = 0;
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        const int p = ( a >> ( 8*i ) ) & 255, q = ( b >> ( 8*i ) ) & 255;
        const int v = ( p * weight + q * ( 64 - weight ) + 32 ) >> 6;
        out |= (uint32_t)min( max( v, 0 ), 255 ) << ( 8*i );
    }
    return out;
}

struct BidirShared
{
    const uint8_t *fenc; int fenc_stride;
    const uint8_t *fref0[4], *fref1[4]; int stride;
    const uint16_t *cost_mv;
    int satd;
};

template <int BW, int BH>
__device__ void me_refine_bidir( const BidirShared &g, uint32_t fenc_off, uint32_t ref0_off, uint32_t ref1_off, const int16_t *mv_in /* m0x m0y m1x m1y */,
                                 const int16_t *mvp /* same order */, const int16_t *limits /* min_x,min_y,max_x,max_y spel */, int weight,
                                 int lane, uint32_t *visited /* 128 words of shared memory, this warp's */, int bm[4], int &out_cost )
{
    constexpr int LX = BW / 4, L = LX * ( BH / 4 ), S = 32 / L;
    const int gl = lane % L, slot = lane / L;
    const int sx = ( gl % LX ) * 4, sy = ( gl / LX ) * 4;
    uint32_t fenc[4];
    {
        const uint8_t *f = g.fenc + fenc_off + sy * g.fenc_stride + sx;
#pragma unroll
        for( int r = 0; r < 4; r++ ) fenc[r] = ldg4u( f + r * g.fenc_stride );
    }
    const int o0 = ref0_off + sy * g.stride + sx, o1 = ref1_off + sy * g.stride + sx;
#pragma unroll
    for( int k = 0; k < 4; k++ ) bm[k] = mv_in[k];
    out_cost = LA_COST_MAX;
    for( int k = 0; k < 4; k++ )                                   // me.c:1076-1080: too close to the window's edge
        if( bm[k] < limits[k & 1] + 8 || bm[k] > limits[2 + ( k & 1 )] - 8 )
            return;
    for( int i = lane; i < 128; i += 32 ) visited[i] = 0;
    __syncwarp();
    LaWeight none; none.enabled = 0; none.scale = 0; none.denom = 0; none.offset = 0;
    int bcost = LA_COST_MAX;
    for( int pass = 0; pass < 8; pass++ )
    {
        int best = 0x7fffffff;
        const int first = pass ? 1 : 0;
        for( int base = first; base < 33; base += S )
        {
            const int j = min( base + slot, 32 );
            int code = bidir_code( j );
            int v[4];
#pragma unroll
            for( int k = 0; k < 4; k++, code /= 3 ) v[k] = bm[k] + code % 3 - 1;
            const int word = ( ( v[0] & 7 ) << 4 ) | ( ( v[1] & 7 ) << 1 ) | ( ( v[2] & 7 ) >> 2 );
            const uint32_t bit = 1u << ( ( ( v[2] & 3 ) << 3 ) | ( v[3] & 7 ) );
            const bool fresh = base + slot < 33 && ( !pass || !( visited[word] & bit ) );
            uint32_t p0[4], p1[4];
            qpel4x4_p( g.fref0[0] + o0, g.fref0[1] + o0, g.fref0[2] + o0, g.fref0[3] + o0, g.stride, none, v[0], v[1], p0 );
            qpel4x4_p( g.fref1[0] + o1, g.fref1[1] + o1, g.fref1[2] + o1, g.fref1[3] + o1, g.stride, none, v[2], v[3], p1 );
#pragma unroll
            for( int r = 0; r < 4; r++ ) p0[r] = avg4_bipred( p0[r], p1[r], weight );
            int d = g.satd ? satd4x4( fenc, p0 ) : sad4x4( fenc, p0 );
#pragma unroll
            for( int m = 1; m < L; m <<= 1 ) d += __shfl_xor_sync( 0xffffffffu, d, m );
            d += __ldg( g.cost_mv + ( v[0] - mvp[0] ) ) + __ldg( g.cost_mv + ( v[1] - mvp[1] ) )
               + __ldg( g.cost_mv + ( v[2] - mvp[2] ) ) + __ldg( g.cost_mv + ( v[3] - mvp[3] ) );
            __syncwarp();
            if( fresh && gl == 0 ) atomicOr( &visited[word], bit );
            __syncwarp();
            const int key = fresh ? ( d << 6 ) | j : 0x7fffffff;
            best = min( best, warp_min( key ) );
        }
        if( best == 0x7fffffff || ( best >> 6 ) >= bcost )
            break;
        const int bestj = best & 63;
        bcost = best >> 6;
        if( !bestj )
            break;
        int code = bidir_code( bestj );
#pragma unroll
        for( int k = 0; k < 4; k++, code /= 3 ) bm[k] += code % 3 - 1;
    }
    out_cost = bcost;
}

} // namespace x264cu
