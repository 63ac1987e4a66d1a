// Lowres lookahead on the GPU: the slot that common/opencl.c + encoder/slicetype-cl.c fill in the reference, with the
// results of the reference's CPU path (encoder/slicetype.c:514-995, one lookahead thread), bit for bit.
//
// slicetype_mb_cost is split along its data dependencies:
//   intra_kernel    lowres intra cost of every MB (3 or 10 modes)            -- no dependencies, once per frame
//   search_kernel   lowres_mvs / lowres_mv_costs of one (frame, list, distance) -- reverse-raster dependency on the
//                   right / below / below-left / below-right neighbours (slicetype.c:662-680): one warp per MB ROW,
//                   rows pipelined two MBs apart through self-validating (generation, mv) records; the reference planes
//                   are staged in a per-warp sliding shared-memory window (cp.async); many jobs per launch
//   finalize_kernel bidir candidates, list choice, intra-vs-inter, AQ scaling, row and frame sums, lowres_costs
//                   -- per-MB parallel once the vectors exist
// The memoisation / order logic of slicetype_frame_cost (sentinels, do_search, b_intra_calculated) stays on the host,
// in x264cu_lookahead_frame_cost below.
#include "ctx.h"
#include "lookahead_dev.cuh"
#include <math.h>
#include <string.h>
#include <vector>

using namespace x264cu;

struct LaHostStats { double put_sync = 0, cost_sync = 0, weight_sync = 0, ev_sync = 0; long n_put = 0, n_cost = 0, n_weight = 0, n_batch = 0, n_jobs = 0; };
#ifdef X264CU_TUNING
// tuning build only: host-side wait accounting (printed at close when X264CU_STATS is set): where the calling thread blocks
#include <time.h>
static inline double la_now() { timespec t; clock_gettime( CLOCK_MONOTONIC, &t ); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define LA_TIMED( acc, cnt, stmt ) do { double _t = la_now(); stmt; ( acc ) += la_now() - _t; ( cnt )++; } while( 0 )
#else
#define LA_TIMED( acc, cnt, stmt ) do { stmt; } while( 0 )
#endif

#define LOWRES_COST_MASK 0x3fff          /* common/frame.h:107-112 */
#define LOWRES_COST_SHIFT 14
#define LA_MAX_B ( X264CU_BFRAME_MAX )
#define LA_PROPAGATE_MAX 32767u          /* MC_CLIP_ADD's ceiling, common/mc.h:29 */

struct LaDims
{
    int mb_w, mb_h, mb_count;
    int stride;                      // lowres stride
    int B;                           // bframes
    int subme;                       // param subpel_refine
    int me_method, subpel, me_range; // what lowres_context_init selects (slicetype.c:45-61)
    int mv_range2;                   // 2 * i_mv_range
    int bipred_weighted, aq, do_edges;
    int cost_len;                    // half length of the cost_mv table
};

struct LaSlotDev
{
    uint8_t *planes[4];              // origins of F,H,V,C
    int16_t *mvs;                    // [2][B+1][mb_count][2]
    int32_t *mv_costs;               // [2][B+1][mb_count]
    uint16_t *costs;                 // [(B+2)*(B+2)][mb_count]
    int32_t *intra;                  // [mb_count]
    uint16_t *qscale;                // [mb_count]
    int32_t *row_satds;              // [(B+2)*(B+2)][mb_h]
    unsigned long long *recs;        // [2][B+1][mb_count] (generation << 32 | mv): how the rows of a search hand over vectors
    unsigned int *propagate;         // [mb_count] i_propagate_cost as a 32-bit accumulator (the reference's u16 saturates at 32767: clamped when read)
    float *qp_offset, *qp_offset_aq; // [mb_count] f_qp_offset / f_qp_offset_aq
};

struct LaSearchJob
{
    const uint8_t *fenc;             // F plane origin of the frame being searched
    const uint8_t *ref[4];
    int16_t *mvs;                    // [mb_count][2]
    int32_t *mv_costs;
    unsigned long long *recs;        // [mb_count] self-validating (gen << 32 | mv) records, see search_kernel
    unsigned int gen;                // the slot's generation: a record is this search's iff its high word equals it
    const uint8_t *ref_w;            // weighted full-pel plane (origin) or NULL
    int w_enabled, w_scale, w_denom, w_offset;
};

#ifndef LA_PACK
#define LA_PACK 128                 /* 96 bytes each: 12 KB of kernel parameters */
#endif
struct LaJobPack { LaSearchJob j[LA_PACK]; };     // passed by value as a kernel parameter: no staging copy, no sync

struct LaFinalizeArgs
{
    const uint8_t *fenc;
    const uint8_t *ref0[4], *ref1[4];
    const int16_t *mvs0, *mvs1;      // this frame's vectors for the two lists (mvs1 NULL for P)
    const int32_t *cost0, *cost1;
    const int16_t *mvr;              // fref1's L0 vectors at distance p1-p0 (temporal direct) or NULL
    int32_t *intra;
    const uint16_t *qscale;
    uint16_t *costs;                 // lowres_costs[b-p0][p1-b]
    int32_t *row_inter, *row_intra;  // row_satds slots
    int32_t *record;                 // {cost_est, cost_est_aq, intra_mbs, intra_cost_est, intra_cost_est_aq}
    int b_bidir, b_inter, dist_scale_factor, bipred_weight;
};

// ------------------------------------------------------------------------------------------------
// lane geometry shared by the per-MB-parallel kernels: 4 lanes per MB (4x4 quadrants), 8 MBs per warp
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_quad( const uint8_t *p, int stride, uint32_t a[4] )
{
#pragma unroll
    for( int r = 0; r < 4; r++ ) a[r] = ldg4u( p + r * stride );
}
// the same for a 4-byte aligned address (a quadrant of a macroblock at vector zero: plane origins and strides are multiples
// of 4): plain word loads, nothing that has to wait for them at the point of issue
__device__ __forceinline__ void load_quad_aligned( const uint8_t *p, int stride, uint32_t a[4] )
{
#pragma unroll
    for( int r = 0; r < 4; r++ ) a[r] = __ldg( (const uint32_t *)( p + r * stride ) );
}

// ------------------------------------------------------------------------------------------------
// intra: slicetype.c:714-757 + predict.c (8x8c DC/H/V/P, 8x8 filtered DDL..HU)
// ------------------------------------------------------------------------------------------------
#define F1( a, b ) ( ( ( a ) + ( b ) + 1 ) >> 1 )
#define F2( a, b, c ) ( ( ( a ) + 2 * ( b ) + ( c ) + 2 ) >> 2 )

__global__ void __launch_bounds__( 256 )
intra_kernel( LaDims d, const uint8_t *__restrict__ plane, int32_t *__restrict__ intra )
{
    // per MB edge arrays in shared memory: raw top t[-1..15], raw left l[0..7], filtered ft[-1..15], fl[-1..7]
    __shared__ uint8_t s_t[64][20], s_l[64][8], s_ft[64][20], s_fl[64][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, mbl = warp * 8 + ( lane >> 2 );
    const int mb = blockIdx.x * 64 + mbl;
    const bool valid = mb < d.mb_count;
    const int mbc = valid ? mb : d.mb_count - 1;
    const int mb_y = mbc / d.mb_w, mb_x = mbc - mb_y * d.mb_w;
    const uint8_t *src = plane + ( mb_y * 8 ) * d.stride + mb_x * 8;
    // stage edges: lane q loads a quarter
    for( int i = q; i < 17; i += 4 ) s_t[mbl][i] = src[-d.stride + i - 1];
    for( int i = q; i < 8; i += 4 ) s_l[mbl][i] = src[i * d.stride - 1];
    __syncwarp();
    {
        const uint8_t *t = &s_t[mbl][1];               // t[-1] = corner
        const uint8_t *l = s_l[mbl];
        uint8_t *ft = &s_ft[mbl][1], *fl = &s_fl[mbl][1];
        if( q == 0 )
        {   // predict_8x8_filter with every neighbour available, predict.c:632-680
            int lt = t[-1];
            ft[-1] = fl[-1] = ( t[0] + 2*lt + l[0] + 2 ) >> 2;
            fl[0] = ( lt + 2*l[0] + l[1] + 2 ) >> 2;
            for( int y = 1; y < 7; y++ ) fl[y] = F2( l[y-1], l[y], l[y+1] );
            fl[7] = ( l[6] + 3*l[7] + 2 ) >> 2;
            ft[0] = ( lt + 2*t[0] + t[1] + 2 ) >> 2;
            for( int x = 1; x < 15; x++ ) ft[x] = F2( t[x-1], t[x], t[x+1] );
            ft[15] = ( t[14] + 3*t[15] + 2 ) >> 2;
        }
    }
    __syncwarp();
    const uint8_t *t = &s_t[mbl][1], *l = s_l[mbl];
    const uint8_t *ft = &s_ft[mbl][1], *fl = &s_fl[mbl][1];
    const int qx = ( q & 1 ) * 4, qy = ( q >> 1 ) * 4;
    uint32_t a[4];
    load_quad( src + qy * d.stride + qx, d.stride, a );
    const bool satd = d.subme > 1;
    int best = LA_COST_MAX;
    const int n_modes = satd ? 10 : 3;
    // plane-prediction parameters (predict.c:282-310)
    int pb = 0, pc = 0, pi00 = 0;
    if( satd )
    {
        int H = 0, V = 0;
        for( int i = 0; i < 4; i++ )
        {
            H += ( i + 1 ) * ( t[4+i] - t[2-i] );
            V += ( i + 1 ) * ( l[i+4] - ( i == 3 ? t[-1] : l[2-i] ) );
        }
        int pa = 16 * ( l[7] + t[7] );
        pb = ( 17*H + 16 ) >> 5; pc = ( 17*V + 16 ) >> 5;
        pi00 = pa - 3*pb - 3*pc + 16;
    }
    // 8x8c DC values (predict.c:221-257)
    int s0 = t[0] + t[1] + t[2] + t[3], s1 = t[4] + t[5] + t[6] + t[7];
    int s2 = l[0] + l[1] + l[2] + l[3], s3 = l[4] + l[5] + l[6] + l[7];
    const int dcq = q == 0 ? ( s0 + s2 + 4 ) >> 3 : q == 1 ? ( s1 + 2 ) >> 2 : q == 2 ? ( s3 + 2 ) >> 2 : ( s1 + s3 + 4 ) >> 3;
    for( int mode = 0; mode < n_modes; mode++ )
    {
        uint32_t b[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const int y = qy + r;
            uint32_t w = 0;
#pragma unroll
            for( int c = 0; c < 4; c++ )
            {
                const int x = qx + c;
                int v;
                switch( mode )
                {
                case 0: v = dcq; break;
                case 1: v = l[y]; break;                                          /* 8x8c H */
                case 2: v = t[x]; break;                                          /* 8x8c V */
                case 3: v = min( max( ( pi00 + pb*x + pc*y ) >> 5, 0 ), 255 ); break;   /* 8x8c P */
                case 4: /* DDL */
                    v = ( x == 7 && y == 7 ) ? F2( ft[14], ft[15], ft[15] ) : F2( ft[x+y], ft[x+y+1], ft[x+y+2] );
                    break;
                case 5: /* DDR */
                {
                    int dd = x - y;
                    v = dd > 0 ? F2( ft[dd-2], ft[dd-1], ft[dd] ) : dd == 0 ? F2( fl[0], ft[-1], ft[0] ) : F2( fl[-dd], fl[-dd-1], fl[-dd-2] );
                    break;
                }
                case 6: /* VR */
                {
                    int z = 2*x - y, k = x - ( y >> 1 );
                    if( z >= 0 ) v = ( z & 1 ) ? F2( ft[k-2], ft[k-1], ft[k] ) : F1( ft[k-1], ft[k] );
                    else if( z == -1 ) v = F2( fl[0], ft[-1], ft[0] );
                    else v = F2( fl[y-2*x-1], fl[y-2*x-2], fl[y-2*x-3] );
                    break;
                }
                case 7: /* HD */
                {
                    int i = 2*( 7 - y ) + x;
                    if( i == 15 ) v = F2( fl[0], ft[-1], ft[0] );
                    else if( i < 16 ) { int k = 6 - ( i >> 1 ); v = ( i & 1 ) ? F2( fl[k-1], fl[k], fl[k+1] ) : F1( fl[k], fl[k+1] ); }
                    else { int k = i - 16; v = F2( ft[k+1], ft[k], ft[k-1] ); }
                    break;
                }
                case 8: /* VL */
                {
                    int k = x + ( y >> 1 );
                    v = ( y & 1 ) ? F2( ft[k], ft[k+1], ft[k+2] ) : F1( ft[k], ft[k+1] );
                    break;
                }
                default: /* HU */
                {
                    int i = 2*y + x, k = i >> 1;
                    if( i >= 14 ) v = fl[7];
                    else if( i == 13 ) v = F2( fl[6], fl[7], fl[7] );
                    else v = ( i & 1 ) ? F2( fl[k], fl[k+1], fl[k+2] ) : F1( fl[k], fl[k+1] );
                    break;
                }
                }
                w |= (uint32_t)v << ( 8*c );
            }
            b[r] = w;
        }
        int c = quad_sum( satd ? satd4x4( a, b ) : sad4x4( a, b ) );
        best = min( best, c );
    }
    if( valid && q == 0 )
        intra[mb] = best + 5 + 4;                 // + intra_penalty (5*lambda, lambda = 1) + lowres_penalty (slicetype.c:745)
}

// ------------------------------------------------------------------------------------------------
// search: one warp = one MB row, rows bottom-up; slicetype.c:654-705 + me.c
// ------------------------------------------------------------------------------------------------
// 64-bit relaxed accesses are single-copy atomic: a record (generation << 32 | mv) is either entirely there or not,
// so the rows of a search synchronise on the data itself -- no fences, no cache invalidation
__device__ __forceinline__ unsigned long long ld_relaxed64( const unsigned long long *p )
{
    unsigned long long v;
    asm volatile( "ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ void st_relaxed64( unsigned long long *p, unsigned long long v )
{
    asm volatile( "st.relaxed.gpu.global.u64 [%0], %1;" ::"l"( p ), "l"( v ) : "memory" );
}

__device__ __forceinline__ void cp_async8( uint32_t saddr, const void *g )
{
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;" ::"r"( saddr ), "l"( g ) : "memory" );
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile( "cp.async.wait_all;" ::: "memory" ); }

// One warp per macroblock ROW, rows pipelined two columns apart (the reverse-raster predictor dependency of
// slicetype.c:658-680).  Work items (row group, job) are handed out by an atomic ticket, row groups first, so that an
// item only ever waits on items with smaller tickets, i.e. CTAs that are already running -- the grid may be larger than
// what is resident.  Each warp keeps a sliding shared-memory window of the reference planes (LaWin) fed by cp.async.
#ifdef LA_PROFILE
// cycles summed over all search warps: {poll, setup, predictors, full-pel, sub-pel, window slide, macroblocks, searched macroblocks}
__device__ unsigned long long g_la_prof[8];
extern "C" int x264cu_debug_la_profile( unsigned long long *out8, int reset )
{
    if( cudaMemcpyFromSymbol( out8, g_la_prof, sizeof( g_la_prof ) ) != cudaSuccess ) return -1;
    if( reset ) { unsigned long long z[8] = { 0 }; cudaMemcpyToSymbol( g_la_prof, z, sizeof( z ) ); }
    return 0;
}
#endif

#ifndef LA_NW
#define LA_NW 8                                        /* warps (= macroblock rows) per CTA */
#endif
#ifndef LA_MIN_CTAS
#define LA_MIN_CTAS 2                                  /* resident CTAs per SM the register allocation is sized for */
#endif
#ifdef LA_MAXNREG
#define LA_SEARCH_BOUNDS __maxnreg__( LA_MAXNREG )
#else
#define LA_SEARCH_BOUNDS __launch_bounds__( NW * 32, LA_MIN_CTAS )
#endif
template <int NW>
__global__ void LA_SEARCH_BOUNDS
search_kernel( const LaDims d, const __grid_constant__ LaJobPack jobs, const uint16_t *__restrict__ cost_mv_g, int n_jobs,
               unsigned int *__restrict__ ticket )
{
    extern __shared__ __align__( 16 ) uint8_t s_raw[];
    __shared__ unsigned int s_ticket;
    __shared__ uint16_t s_cost[2 * LA_COST_HALF + 2];
    if( threadIdx.x == 0 ) s_ticket = atomicAdd( ticket, 1u );
    const uint16_t *cost_mv = cost_mv_g + d.cost_len;
    for( int i = threadIdx.x; i <= 2 * LA_COST_HALF; i += blockDim.x )
    {   // the table is symmetric around its centre and at least cost_len long on either side
        const int idx = i - LA_COST_HALF;
        s_cost[i] = abs( idx ) <= d.cost_len ? cost_mv[idx] : 0xFFFF;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t win_base = (uint32_t)__cvta_generic_to_shared( s_raw ) + warp * LA_WIN_BYTES;
    const uint32_t cost_s = (uint32_t)__cvta_generic_to_shared( s_cost + LA_COST_HALF );

    const int row_group = s_ticket / n_jobs;
    const LaSearchJob &job = jobs.j[s_ticket - row_group * n_jobs];
    const int start_y = min( d.mb_h - 1, d.mb_h - 2 + d.do_edges ), end_y = max( 0, 1 - d.do_edges );
    const int start_x = d.mb_w - 2 + d.do_edges, end_x = 1 - d.do_edges;
    const int mb_y = start_y - ( row_group * NW + warp );
    if( mb_y < end_y ) return;
    const int q = lane & 3, qx = ( q & 1 ) * 4, qy = ( q >> 1 ) * 4;

    // ---- reference window ---------------------------------------------------------------------------
    const uint8_t *g0 = job.ref_w ? job.ref_w : job.ref[0], *g1 = job.ref[1], *g2 = job.ref[2], *g3 = job.ref[3];
    const int n_chunks = d.stride >> 3;
    const int win_row0 = mb_y * 8 - LA_WIN_VR;
    auto load_chunk = [&]( int k ) {
#pragma unroll
        for( int j = 0; j < 4 * LA_WIN_ROWS / 32; j++ )
        {
            const int i = lane + 32 * j, p = i / LA_WIN_ROWS, r = i - p * LA_WIN_ROWS;
            const uint8_t *g = ( p == 0 ? g0 : p == 1 ? g1 : p == 2 ? g2 : g3 ) + (ptrdiff_t)( win_row0 + r ) * d.stride + k * 8 - 64;
            cp_async8( win_base + p * LA_WIN_PLANE + r * LA_WIN_PITCH + ( k & 7 ) * 8, g );
        }
    };
    // chunk of the macroblock itself = mb_x + 8 (biased column >> 3); chunks [kc-3, kc+3] are valid while kc-4 is in flight
    int kc = start_x + 8;
    for( int k = max( kc - 3, 0 ); k <= min( kc + 3, n_chunks - 1 ); k++ ) load_chunk( k );
    cp_async_wait_all();
    __syncwarp();
    if( kc - 4 >= 0 ) load_chunk( kc - 4 );

    int16_t *mvs = job.mvs;
    // right neighbour's vector is produced by this warp itself; the never-searched edge column keeps what the frame
    // reset left there: zeros (frame.c:287-293)
    int right_x = 0, right_y = 0;
    const int mv_range = d.mv_range2;
    // Vectors of the row below (below, below-left, below-right) come from another warp, through self-validating
    // records.  That row runs right to left too, so one NEW record is needed per macroblock -- below-left -- and the
    // other two are the previous macroblock's below-left and below.  Its load is issued one macroblock early
    // (speculatively: it is re-polled if the generation is not there yet), as is the load of the next fenc block, so
    // that no L2 round trip sits between two searches of a row.
    const bool has_below = mb_y < d.mb_h - 1;
    const unsigned long long *rrow = job.recs + ( mb_y + 1 ) * d.mb_w;       // dereferenced only where rec_needed()
    // columns / rows outside the searched range are never written: they read as the reset value, zero
    auto rec_needed = [&]( int c ) { return has_below && mb_y + 1 <= start_y && c >= end_x && c <= start_x; };
#ifndef LA_POLL_SPINS
#define LA_POLL_SPINS 16
#define LA_POLL_NS 40
#endif
    auto rec_poll = [&]( unsigned long long r, const unsigned long long *p ) {
        for( int spins = 0; (unsigned int)( r >> 32 ) != job.gen; r = ld_relaxed64( p ) )
            if( ++spins > LA_POLL_SPINS ) __nanosleep( LA_POLL_NS );
        return (int)(unsigned int)r;
    };
    int nb0 = 0, nb2 = 0;                                                      // below, below-right
    unsigned long long spec = 0;
    if( rec_needed( start_x ) ) nb0 = rec_poll( 0, rrow + start_x );
    if( rec_needed( start_x - 1 ) ) spec = ld_relaxed64( rrow + start_x - 1 );
    uint32_t fenc_next[4];
    load_quad_aligned( job.fenc + ( mb_y * 8 + qy ) * d.stride + start_x * 8 + qx, d.stride, fenc_next );
#ifdef LA_PROFILE
    long long pr[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, t_prof = clock64();
#endif
    for( int mb_x = start_x; mb_x >= end_x; mb_x-- )
    {
        const int mb_xy = mb_y * d.mb_w + mb_x;
        LaMe m;
#ifdef LA_PROFILE
        m.prof[0] = m.prof[1] = m.prof[2] = 0;
#endif
        const int pel = ( mb_y * 8 ) * d.stride + mb_x * 8;
#pragma unroll
        for( int i = 0; i < 4; i++ ) m.fenc[i] = fenc_next[i];
        if( mb_x + 8 != kc )
        {   // slide the window: chunk kc-3 (requested one macroblock ago) becomes valid, chunk kc-4 is requested
            kc = mb_x + 8;
            cp_async_wait_all();
            __syncwarp();
            if( kc - 4 >= 0 ) load_chunk( kc - 4 );
        }
        LA_TICK( pr[5], t_prof );
        int nb1 = 0;                                                           // below-left
        if( rec_needed( mb_x - 1 ) ) nb1 = rec_poll( spec, rrow + mb_x - 1 );
        if( rec_needed( mb_x - 2 ) ) spec = ld_relaxed64( rrow + mb_x - 2 );
        if( mb_x > end_x ) load_quad_aligned( job.fenc + pel - 8 + qy * d.stride + qx, d.stride, fenc_next );
        LA_TICK( pr[0], t_prof );
        m.win.base = win_base;
        m.win.bx = mb_x * 8 + qx + 64;
        m.win.ry = qy + LA_WIN_VR;
        {   // loaded chunks [vlo, vhi): both words of a 4-px read lie inside <=> 8*vlo <= x <= 8*vhi - 5
            const int vlo = max( kc - 3, 0 ), vhi = min( kc + 4, n_chunks );
            m.win.dxlo = 8 * vlo - m.win.bx;
            m.win.dxspan = (unsigned)( 8 * ( vhi - vlo ) - 5 );
        }
        m.win.on = true;
        m.win.p0w = job.ref_w != nullptr;
#pragma unroll
        for( int i = 0; i < 4; i++ ) m.fref[i] = job.ref[i] + pel + qy * d.stride + qx;
        m.fref_w = job.ref_w ? job.ref_w + pel + qy * d.stride + qx : m.fref[0];
        m.stride = d.stride;
        m.cost_mv = cost_mv;
        m.cost_s = cost_s;
        m.w.enabled = job.w_enabled; m.w.scale = job.w_scale; m.w.denom = job.w_denom; m.w.offset = job.w_offset;
        m.satd = d.subme > 1;
        m.min_spel_x = max( 4*( -8*mb_x - 12 ), -mv_range );
        m.max_spel_x = min( 4*( 8*( d.mb_w - mb_x - 1 ) + 12 ), mv_range - 1 );
        m.min_spel_y = max( 4*( -8*mb_y - 12 ), -mv_range );
        m.max_spel_y = min( 4*( 8*( d.mb_h - mb_y - 1 ) + 12 ), mv_range - 1 );
        m.x_min = m.min_spel_x >> 2; m.x_max = m.max_spel_x >> 2;
        m.y_min = m.min_spel_y >> 2; m.y_max = m.max_spel_y >> 2;

        // reverse-order predictors: right, below, below-left, below-right (slicetype.c:658-680)
        int mvcx[4] = { 0, 0, 0, 0 }, mvcy[4] = { 0, 0, 0, 0 }, i_mvc = 0;
        if( mb_x < d.mb_w - 1 ) { mvcx[0] = right_x; mvcy[0] = right_y; i_mvc = 1; }
        if( mb_y < d.mb_h - 1 )
        {
            int bx = (int16_t)( nb0 & 0xffff ), by = (int16_t)( nb0 >> 16 );
            if( i_mvc == 0 ) { mvcx[0] = bx; mvcy[0] = by; } else { mvcx[1] = bx; mvcy[1] = by; }
            i_mvc++;
            if( mb_x > 0 )
            {
                int cx = (int16_t)( nb1 & 0xffff ), cy = (int16_t)( nb1 >> 16 );
                if( i_mvc == 1 ) { mvcx[1] = cx; mvcy[1] = cy; } else { mvcx[2] = cx; mvcy[2] = cy; }
                i_mvc++;
            }
            if( mb_x < d.mb_w - 1 )
            {
                int cx = (int16_t)( nb2 & 0xffff ), cy = (int16_t)( nb2 >> 16 );
                if( i_mvc == 1 ) { mvcx[1] = cx; mvcy[1] = cy; } else if( i_mvc == 2 ) { mvcx[2] = cx; mvcy[2] = cy; } else { mvcx[3] = cx; mvcy[3] = cy; }
                i_mvc++;
            }
        }
        if( i_mvc <= 1 ) { m.mvpx = mvcx[0]; m.mvpy = mvcy[0]; }
        else
        {   // x264_median_mv of the first three (the third is zero when only two exist), base.h:232-246
            m.mvpx = max( min( mvcx[0], mvcx[1] ), min( max( mvcx[0], mvcx[1] ), mvcx[2] ) );
            m.mvpy = max( min( mvcy[0], mvcy[1] ), min( max( mvcy[0], mvcy[1] ), mvcy[2] ) );
        }
        int mvx = 0, mvy = 0, cost = 0;
        bool skip = false;
        if( !( m.mvpx | m.mvpy ) )
        {   // zero-predictor fast skip, slicetype.c:684-692
            uint32_t b[4];
            la_load4( m, 0, 0, 0, b );
            cost = quad_sum( m.satd ? satd4x4( m.fenc, b ) : sad4x4( m.fenc, b ) );
            cost = __shfl_sync( 0xffffffffu, cost, 0 );
            skip = cost < 64;
        }
        LA_TICK( pr[1], t_prof );
        if( !skip )
        {
            la_me_search( m, d.me_method, d.subpel, d.me_range, mvcx, mvcy, i_mvc, lane, mvx, mvy, cost );
            cost -= la_cost( m, 0 );
            if( mvx | mvy ) cost += 5;                              // 5 * lambda
        }
        if( lane == 0 )
        {
            const unsigned int mv = pack_mv( mvx, mvy );
            st_relaxed64( job.recs + mb_xy, ( (unsigned long long)job.gen << 32 ) | mv );
            *(int *)( mvs + 2 * mb_xy ) = (int)mv;
            job.mv_costs[mb_xy] = cost;
        }
        right_x = mvx; right_y = mvy;
        nb2 = nb0; nb0 = nb1;
#ifdef LA_PROFILE
        t_prof = clock64();
        pr[2] += m.prof[0]; pr[3] += m.prof[1]; pr[4] += m.prof[2]; pr[6]++; pr[7] += !skip;
#endif
    }
#ifdef LA_PROFILE
    if( lane == 0 )
        for( int i = 0; i < 8; i++ ) atomicAdd( &g_la_prof[i], (unsigned long long)pr[i] );
#endif
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------
// finalize: slicetype.c:579-652, :706-712, :758-790
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t avg4w( uint32_t a, uint32_t b, int weight )     // pixel_avg_weight_wxh, mc.c:63-75
{
    uint32_t out = 0;
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int p = ( a >> ( 8*i ) ) & 255, qv = ( b >> ( 8*i ) ) & 255;
        int v = ( p * weight + qv * ( 64 - weight ) + 32 ) >> 6;
        v = min( max( v, 0 ), 255 );
        out |= (uint32_t)v << ( 8*i );
    }
    return out;
}

// 128-thread blocks at <= 64 registers: one fits on an SM beside two resident search CTAs (LA_FIN_THREADS * 64 registers are
// what the search kernel's register cap leaves free), so a cost request does not wait for a search CTA to retire
#define LA_FIN_THREADS 128
__device__ __forceinline__ void finalize_body( const LaDims &d, const LaFinalizeArgs &A )
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3;
    const int mb = ( blockIdx.x * ( blockDim.x >> 5 ) + warp ) * 8 + ( lane >> 2 );
    const bool in_frame = mb < d.mb_count;
    const int mbc = in_frame ? mb : d.mb_count - 1;
    const int mb_y = mbc / d.mb_w, mb_x = mbc - mb_y * d.mb_w;
    // MBs the reference visits (slicetype.c:823-833)
    const bool visited = in_frame && ( d.do_edges || d.mb_w <= 2 || d.mb_h <= 2 ||
                                       ( mb_x >= 1 && mb_x <= d.mb_w - 2 && mb_y >= 1 && mb_y <= d.mb_h - 2 ) );
    const bool score_mb = ( mb_x > 0 && mb_x < d.mb_w - 1 && mb_y > 0 && mb_y < d.mb_h - 1 ) || d.mb_w <= 2 || d.mb_h <= 2;
    const int qx = ( q & 1 ) * 4, qy = ( q >> 1 ) * 4;
    const int pel = ( mb_y * 8 ) * d.stride + mb_x * 8 + qy * d.stride + qx;
    const bool satd = d.subme > 1;
    int bcost = LA_COST_MAX, list_used = 0;

    if( A.b_inter )
    {
        uint32_t a[4];
        load_quad_aligned( A.fenc + pel, d.stride, a );
        LaMe m0, m1;
        m0.stride = m1.stride = d.stride;
        m0.w.enabled = m1.w.enabled = 0;
        m0.win.on = m1.win.on = false; m0.win.p0w = m1.win.p0w = false;
#pragma unroll
        for( int i = 0; i < 4; i++ ) { m0.fref[i] = A.ref0[i] + pel; m1.fref[i] = ( A.b_bidir ? A.ref1[i] : A.ref0[i] ) + pel; }
        const int v0 = *(const int *)( A.mvs0 + 2 * mbc );
        const int mv0x = (int16_t)( v0 & 0xffff ), mv0y = (int16_t)( v0 >> 16 );
        const int c0 = A.cost0[mbc];
        auto try_bidir = [&]( int ax, int ay, int bx, int by, int penalty ) {
            uint32_t p1[4], p2[4];
            if( d.subme <= 1 )
            {   // half-pel planes addressed directly, slicetype.c:589-596
                int h1 = ( ( ax & 2 ) >> 1 ) + ( ay & 2 ), h2 = ( ( bx & 2 ) >> 1 ) + ( by & 2 );
                load_quad( m0.fref[h1] + ( ax >> 2 ) + ( ay >> 2 ) * d.stride, d.stride, p1 );
                load_quad( m1.fref[h2] + ( bx >> 2 ) + ( by >> 2 ) * d.stride, d.stride, p2 );
            }
            else
            {
                qpel4x4( m0, ax, ay, p1 );
                qpel4x4( m1, bx, by, p2 );
            }
#pragma unroll
            for( int r = 0; r < 4; r++ )
                p1[r] = A.bipred_weight == 32 ? __vavgu4( p1[r], p2[r] ) : avg4w( p1[r], p2[r], A.bipred_weight );
            int c = penalty + quad_sum( satd ? satd4x4( a, p1 ) : sad4x4( a, p1 ) );
            if( c < bcost ) { bcost = c; list_used = 3; }
        };
        int mv1x = 0, mv1y = 0, c1 = 0;
        if( A.b_bidir )
        {
            const int v1 = *(const int *)( A.mvs1 + 2 * mbc );
            mv1x = (int16_t)( v1 & 0xffff ); mv1y = (int16_t)( v1 >> 16 );
            c1 = A.cost1[mbc];
            int d0x = 0, d0y = 0, d1x = 0, d1y = 0;
            if( A.mvr )
            {   // temporal direct, slicetype.c:629-642
                const int vr = *(const int *)( A.mvr + 2 * mbc );
                const int rx = (int16_t)( vr & 0xffff ), ry = (int16_t)( vr >> 16 );
                const int mv_range = d.mv_range2;
                const int minx = max( 4*( -8*mb_x - 12 ), -mv_range ), maxx = min( 4*( 8*( d.mb_w - mb_x - 1 ) + 12 ), mv_range - 1 );
                const int miny = max( 4*( -8*mb_y - 12 ), -mv_range ), maxy = min( 4*( 8*( d.mb_h - mb_y - 1 ) + 12 ), mv_range - 1 );
                d0x = ( rx * A.dist_scale_factor + 128 ) >> 8;
                d0y = ( ry * A.dist_scale_factor + 128 ) >> 8;
                d1x = d0x - rx; d1y = d0y - ry;
                d0x = (int16_t)d0x; d0y = (int16_t)d0y; d1x = (int16_t)d1x; d1y = (int16_t)d1y;
                d0x = clip3i( d0x, minx, maxx ); d0y = clip3i( d0y, miny, maxy );
                d1x = clip3i( d1x, minx, maxx ); d1y = clip3i( d1y, miny, maxy );
                if( d.subme <= 1 ) { d0x &= ~1; d0y &= ~1; d1x &= ~1; d1y &= ~1; }
            }
            try_bidir( d0x, d0y, d1x, d1y, 0 );
            // the zero-vector pair is tried only when the temporal pair is not already zero; computed for all
            // lanes (uniform code), applied per MB
            {
                uint32_t p1[4], p2[4];
                load_quad_aligned( m0.fref[0], d.stride, p1 );
                load_quad_aligned( m1.fref[0], d.stride, p2 );
#pragma unroll
                for( int r = 0; r < 4; r++ )
                    p1[r] = A.bipred_weight == 32 ? __vavgu4( p1[r], p2[r] ) : avg4w( p1[r], p2[r], A.bipred_weight );
                int c = quad_sum( satd ? satd4x4( a, p1 ) : sad4x4( a, p1 ) );
                if( ( d0x | d0y | d1x | d1y ) && c < bcost ) { bcost = c; list_used = 3; }
            }
        }
        if( c0 < bcost ) { bcost = c0; list_used = 1; }
        if( A.b_bidir )
        {
            if( c1 < bcost ) { bcost = c1; list_used = 2; }
            // (mv0, mv1) pair with the 5*lambda penalty; evaluated by every lane group, applied where a vector is non-zero
            int sb = bcost, sl = list_used;
            try_bidir( mv0x, mv0y, mv1x, mv1y, 5 );
            if( !( mv0x | mv0y | mv1x | mv1y ) ) { bcost = sb; list_used = sl; }
        }
    }

    const int icost = A.intra[mbc];
    bcost += 4;                                                     // lowres_penalty
    int b_intra = 0;
    if( !A.b_bidir )
    {
        b_intra = icost < bcost;
        if( b_intra ) { bcost = icost; list_used = 0; }
    }
    const int qs = d.aq ? A.qscale[mbc] : 256;
    const int icost_aq = d.aq ? ( icost * qs + 128 ) >> 8 : icost;
    const int bcost_aq = d.aq ? ( bcost * qs + 128 ) >> 8 : bcost;
    const bool lead = visited && q == 0;
    // frame-level sums: {cost_est, cost_est_aq, intra_mbs, intra cost_est, intra cost_est_aq}
    int sums[5];
    sums[0] = ( lead && score_mb && A.b_inter ) ? bcost : 0;
    sums[1] = ( lead && score_mb && A.b_inter ) ? bcost_aq : 0;
    sums[2] = ( lead && score_mb && !A.b_bidir ) ? b_intra : 0;
    sums[3] = ( lead && score_mb ) ? icost : 0;
    sums[4] = ( lead && score_mb ) ? icost_aq : 0;
#pragma unroll
    for( int i = 0; i < 5; i++ )
    {
        int v = __reduce_add_sync( 0xffffffffu, sums[i] );
        if( lane == 0 && v ) atomicAdd( &A.record[i], v );
    }
    if( lead )
    {
        if( A.b_inter && A.row_inter ) atomicAdd( &A.row_inter[mb_y], bcost_aq );
        if( A.row_intra ) atomicAdd( &A.row_intra[mb_y], icost_aq );
        A.costs[mb] = (uint16_t)( min( bcost, LOWRES_COST_MASK ) + ( list_used << LOWRES_COST_SHIFT ) );
        // i_intra_cost IS lowres_costs[0][0] in the reference (frame.c:287): an I request leaves the clipped value behind
        if( !A.b_inter ) A.intra[mb] = min( bcost, LOWRES_COST_MASK );
    }
}

__global__ void __launch_bounds__( LA_FIN_THREADS, 8 )
finalize_kernel( LaDims d, LaFinalizeArgs A )
{
    finalize_body( d, A );
}

// many cost requests in one launch (x264cu_lookahead_finalize_batch): blockIdx.y selects the request
__global__ void __launch_bounds__( LA_FIN_THREADS, 8 )
finalize_batch_kernel( LaDims d, const LaFinalizeArgs *__restrict__ args )
{
    const LaFinalizeArgs A = args[blockIdx.y];
    finalize_body( d, A );
}

// ------------------------------------------------------------------------------------------------
// lookahead weighted prediction (slicetype.c:284-501, luma, b_lookahead = 1)
// ------------------------------------------------------------------------------------------------
// frame statistics of x264_adaptive_quant_frame (ratecontrol.c:225-233): sum and sum of squares over the mod-16 picture
__global__ void __launch_bounds__( 256 )
luma_stats_kernel( const uint8_t *__restrict__ src, intptr_t stride, int width, int height, int w16, int h16, unsigned long long *out )
{
    unsigned long long sum = 0, sqr = 0;
    const int groups = w16 / 4;
    for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < groups * h16; i += gridDim.x * blockDim.x )
    {
        int y = i / groups, x = ( i - y * groups ) * 4;
        const uint8_t *row = src + (intptr_t)min( y, height - 1 ) * stride;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            unsigned v = row[min( x + k, width - 1 )];
            sum += v; sqr += v * v;
        }
    }
    for( int m = 16; m; m >>= 1 )
    {
        sum += __shfl_xor_sync( 0xffffffffu, sum, m );
        sqr += __shfl_xor_sync( 0xffffffffu, sqr, m );
    }
    if( ( threadIdx.x & 31 ) == 0 ) { atomicAdd( &out[0], sum ); atomicAdd( &out[1], sqr ); }
}

// weight_cost_luma, slicetype.c:191-222 (without the header bits): sum over MBs of min( mbcmp( w(ref), fenc ), intra )
__global__ void __launch_bounds__( 256 )
weight_cost_kernel( LaDims d, const uint8_t *__restrict__ fenc, const uint8_t *__restrict__ ref, const int32_t *__restrict__ intra,
                    LaWeight w, unsigned int *out )
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3;
    const int mb = ( blockIdx.x * 8 + warp ) * 8 + ( lane >> 2 );
    const bool valid = mb < d.mb_count;
    const int mbc = valid ? mb : d.mb_count - 1;
    const int mb_y = mbc / d.mb_w, mb_x = mbc - mb_y * d.mb_w;
    const int pel = ( mb_y * 8 + ( q >> 1 ) * 4 ) * d.stride + mb_x * 8 + ( q & 1 ) * 4;
    uint32_t a[4], b[4];
    load_quad_aligned( fenc + pel, d.stride, a );
    load_quad_aligned( ref + pel, d.stride, b );
    if( w.enabled )
    {
#pragma unroll
        for( int r = 0; r < 4; r++ ) b[r] = weight4( b[r], w );
    }
    int cmp = quad_sum( d.subme > 1 ? satd4x4( b, a ) : sad4x4( b, a ) );
    int icost = (int)(uint16_t)intra[mbc];                  // i_intra_cost is a u16 array in the reference (frame.h:135)
    // border MBs are never costed without do_edges: their i_intra_cost keeps its initial 0xFFFF (frame.c:288)
    if( !d.do_edges && ( mb_x == 0 || mb_y == 0 || mb_x == d.mb_w - 1 || mb_y == d.mb_h - 1 ) ) icost = 0xFFFF;
    int v = ( valid && q == 0 ) ? min( cmp, icost ) : 0;
    v = __reduce_add_sync( 0xffffffffu, v );
    if( lane == 0 && v ) atomicAdd( out, (unsigned int)v );
}

// x264_weight_scale_plane over a whole padded plane (slicetype.c:489-500, frame.c:825-841)
__global__ void __launch_bounds__( 256 )
weight_plane_kernel( const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t n_words, LaWeight w )
{
    for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x )
        dst[i] = weight4( src[i], w );
}

// ------------------------------------------------------------------------------------------------
// MB-tree: macroblock_tree_propagate (slicetype.c:1050-1089) = mbtree_propagate_cost + mbtree_propagate_list
// (common/mc.c:511-598) and macroblock_tree_finish (slicetype.c:1029-1048), one thread per macroblock.
// The reference adds into u16 arrays saturating at (1<<15)-1 (MC_CLIP_ADD, common/mc.h:29), MB after MB; every addend is >= 0,
// so the result is min( sum, 32767 ) whatever the order: here 32-bit atomics, clamped where the value is read.
// Float expressions are evaluated operation by operation (no FMA contraction) in the reference's order.
// ------------------------------------------------------------------------------------------------
struct LaMbtreeArgs
{
    const unsigned int *prop_in;     // frame b's own accumulator (referenced frames) or NULL (all zero)
    unsigned int *ref_costs[2];      // accumulators of p0 / p1 (the second NULL for P frames)
    const int16_t *mvs[2];
    const int32_t *intra;
    const uint16_t *lowres_costs, *qscale;
    int bipred_weight[2];
    float fps_factor;
};

__global__ void __launch_bounds__( 256 )
mbtree_propagate_kernel( LaDims d, LaMbtreeArgs A )
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if( mb >= d.mb_count ) return;
    const int mb_y = mb / d.mb_w, mb_x = mb - mb_y * d.mb_w;
    // mbtree_propagate_cost, mc.c:511-527
    const int intra_cost = (uint16_t)A.intra[mb];
    const int lc = A.lowres_costs[mb];
    const int inter_cost = min( intra_cost, lc & LOWRES_COST_MASK );
    const float propagate_intra = (float)( intra_cost * (int)A.qscale[mb] );
    const float propagate_in = (float)( A.prop_in ? min( A.prop_in[mb], LA_PROPAGATE_MAX ) : 0u );
    const float propagate_amount = __fadd_rn( propagate_in, __fmul_rn( propagate_intra, A.fps_factor ) );
    const float num = (float)( intra_cost - inter_cost ), denom = (float)intra_cost;
    const int amount = min( __float2int_rz( __fadd_rn( __fdiv_rn( __fmul_rn( propagate_amount, num ), denom ), 0.5f ) ), 32767 );
    const int lists_used = lc >> LOWRES_COST_SHIFT;
    // mbtree_propagate_list, mc.c:529-598
#pragma unroll
    for( int list = 0; list < 2; list++ )
    {
        unsigned int *ref_costs = A.ref_costs[list];
        if( !ref_costs || !( lists_used & ( 1 << list ) ) ) continue;
        int listamount = amount;
        if( lists_used == 3 ) listamount = ( listamount * A.bipred_weight[list] + 32 ) >> 6;
        const int v = *(const int *)( A.mvs[list] + 2 * mb );
        int x = (int16_t)( v & 0xffff ), y = (int16_t)( v >> 16 );
        if( !v ) { if( listamount ) atomicAdd( &ref_costs[mb], (unsigned)listamount ); continue; }
        const unsigned mbx = (unsigned)( ( x >> 5 ) + mb_x ), mby = (unsigned)( ( y >> 5 ) + mb_y );
        const unsigned idx0 = mbx + mby * d.mb_w, idx2 = idx0 + d.mb_w;
        x &= 31; y &= 31;
        const int w0 = ( ( 32 - y ) * ( 32 - x ) * listamount + 512 ) >> 10, w1 = ( ( 32 - y ) * x * listamount + 512 ) >> 10;
        const int w2 = ( y * ( 32 - x ) * listamount + 512 ) >> 10, w3 = ( y * x * listamount + 512 ) >> 10;
        if( mby < (unsigned)d.mb_h )
        {
            if( mbx < (unsigned)d.mb_w && w0 ) atomicAdd( &ref_costs[idx0], (unsigned)w0 );
            if( mbx + 1 < (unsigned)d.mb_w && w1 ) atomicAdd( &ref_costs[idx0 + 1], (unsigned)w1 );
        }
        if( mby + 1 < (unsigned)d.mb_h )
        {
            if( mbx < (unsigned)d.mb_w && w2 ) atomicAdd( &ref_costs[idx2], (unsigned)w2 );
            if( mbx + 1 < (unsigned)d.mb_w && w3 ) atomicAdd( &ref_costs[idx2 + 1], (unsigned)w3 );
        }
    }
}

__constant__ float c_log2_lut[128];          // x264_log2_lut, common/tables.c:66-85

__global__ void __launch_bounds__( 256 )
mbtree_finish_kernel( int mb_count, const int32_t *__restrict__ intra, const uint16_t *__restrict__ qscale,
                      const unsigned int *__restrict__ propagate, const float *__restrict__ qp_aq, float *__restrict__ qp,
                      int fps_factor, float weightdelta, float strength )
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if( mb >= mb_count ) return;
    const int intra_cost = ( (int)(uint16_t)intra[mb] * (int)qscale[mb] + 128 ) >> 8;
    if( !intra_cost ) return;
    const int propagate_cost = (int)( ( min( propagate[mb], LA_PROPAGATE_MAX ) * (unsigned)fps_factor + 128u ) >> 8 );
    // x264_log2( a ) - x264_log2( b ) + weightdelta (base.h:226-230).  The reference is built with -ffast-math (configure:1413),
    // which lets the compiler re-associate the five float terms; this is the association gcc 13 -O3 emits for
    // slicetype.c:1044 (read from the disassembly of the compiled reference): bit-identical to that build,
    // a few ulp of 16.0 from any other
    const uint32_t a = intra_cost + propagate_cost, b = intra_cost;
    const int lza = __clz( a ), lzb = __clz( b );
    const float fa = c_log2_lut[( a << lza >> 24 ) & 0x7f], fb = c_log2_lut[( b << lzb >> 24 ) & 0x7f];
    const float ia = (float)( 31 - lza ), ib = (float)( 31 - lzb );
    const float ratio = __fsub_rn( __fadd_rn( __fsub_rn( fa, ib ), __fadd_rn( ia, weightdelta ) ), fb );
    qp[mb] = __fsub_rn( qp_aq[mb], __fmul_rn( strength, ratio ) );
}

// slicetype_frame_cost_recalculate, slicetype.c:999-1024: one CTA per macroblock row; integer sums (order-free)
__constant__ uint8_t c_exp2_lut[64];         // x264_exp2_lut, common/tables.c:58-64

__global__ void __launch_bounds__( 128 )
recalculate_kernel( int mb_w, int mb_h, const uint16_t *__restrict__ costs, const int32_t *__restrict__ intra, const float *__restrict__ qp_offset,
                    int32_t *__restrict__ row_satd, int *__restrict__ score )
{
    __shared__ int s_row[4], s_in[4];
    const int y = blockIdx.x;
    const bool all = mb_w <= 2 || mb_h <= 2, row_in = y > 0 && y < mb_h - 1;
    int row = 0, in = 0;
    for( int x = threadIdx.x; x < mb_w; x += blockDim.x )
    {
        const int mb = x + y * mb_w;
        // x264_exp2fix8 (base.h:218-224): the product and the sum are rounded separately, as the reference's build does
        const int i = __float2int_rz( __fadd_rn( __fmul_rn( qp_offset[mb], -64.f / 6.f ), 512.5f ) );
        const int q = i < 0 ? 0 : i > 1023 ? 0xffff : (int)( ( ( (unsigned)c_exp2_lut[i & 63] + 256u ) << ( i >> 6 ) ) >> 8 );
        // lowres_costs[0][0] IS i_intra_cost (a u16 array) in the reference (frame.c:287): the intra request reads that one
        const int lc = intra ? ( intra[mb] & 0xffff ) : costs[mb];
        const int cost = ( ( lc & 0x3fff ) * q + 128 ) >> 8;
        row += cost;
        if( all || ( row_in && x > 0 && x < mb_w - 1 ) ) in += cost;
    }
#pragma unroll
    for( int m = 16; m; m >>= 1 ) { row += __shfl_xor_sync( 0xffffffffu, row, m ); in += __shfl_xor_sync( 0xffffffffu, in, m ); }
    if( !( threadIdx.x & 31 ) ) { s_row[threadIdx.x >> 5] = row; s_in[threadIdx.x >> 5] = in; }
    __syncthreads();
    if( !threadIdx.x )
    {
        row_satd[y] = s_row[0] + s_row[1] + s_row[2] + s_row[3];
        atomicAdd( score, s_in[0] + s_in[1] + s_in[2] + s_in[3] );
    }
}

// ================================================================================================
// host side
// ================================================================================================
struct LaSlotHost
{
    LaSlotDev dev;
    uint8_t *plane_buf;              // 4 planes
    bool in_use = false;
    bool intra_on_device = false;    // intra costs computed for the current picture
    int b_intra_calculated = 0;
    int cost_est[LA_MAX_B + 2][LA_MAX_B + 2], cost_est_aq[LA_MAX_B + 2][LA_MAX_B + 2];
    int intra_mbs[LA_MAX_B + 2];
    bool row_satds_valid[LA_MAX_B + 2][LA_MAX_B + 2];
    unsigned int gen = 0;            // bumped at every reset: stale (gen << 32 | mv) records of earlier pictures are invalid
    bool searched[2][LA_MAX_B + 1];  // a search of this (list, distance) has been launched (on demand or ahead of time)
    bool requested[2][LA_MAX_B + 1]; // ... and a cost request has asked for it: the reference's 0x7FFF sentinel of lowres_mvs[l][d][0][0] is gone
    int pending[2][LA_MAX_B + 1];    // event index of a prefetched search still in flight on the search stream, or -1
    unsigned long long *d_stats;     // {sum, sum of squares} of the mod-16 luma
    unsigned long long pixel_sum, pixel_ssd;   // i_pixel_sum[0] / i_pixel_ssd[0] (ratecontrol.c:405-414), valid once stats_ready
    bool stats_ready;
    LaWeight weight;                 // fenc->weight[0][0] of the last lookahead analysis
    float weighted_cost_delta[LA_MAX_B + 1];   // f_weighted_cost_delta (slicetype.c:462-463, X264_WEIGHTP_FAKE only)
    // event (ring index, sequence number) of the last launch that reads or writes this slot on each of the two search streams
    // ([0], [1]: they are not ordered against each other) and of the last import on the exchange stream ([2])
    // ... and of the last batch of speculative cost requests ([3])
    int last_search_ev[4] = { -1, -1, -1, -1 };
    unsigned long long last_search_seq[4] = { 0, 0, 0, 0 };
    // cost requests answered ahead of time (x264cu_lookahead_finalize_batch): which batch holds the record of (b-p0, p1-b)
    struct Spec { unsigned long long batch; int index; } spec[LA_MAX_B + 2][LA_MAX_B + 2];
    cudaEvent_t ev_ready = nullptr;  // recorded on the upload stream when the picture's planes / reset arrays are in place
    bool main_waited = true;         // the context's stream has been ordered after ev_ready
    bool xch_dirty = false;
    unsigned long long mt_last = 0;  // sequence number of the last MB-tree operation that touched this slot's arrays
};

struct x264cu_lookahead
{
    x264cu_ctx *ctx;
    x264cu_lookahead_params_t p;
    LaDims d;
    int wl, ll;
    size_t plane_bytes;              // one padded lowres plane
    std::vector<LaSlotHost> slots;
    uint16_t *d_cost_mv = nullptr;
    uint8_t *d_chroma = nullptr;     // staging for the two chroma planes of an I420 picture (frame_put_i420)
    float *d_aq_q4 = nullptr;        // scratch of the auto-variance AQ modes
    uint8_t *d_luma = nullptr;       // staging for one full-res luma picture
    size_t luma_bytes = 0;
    uint8_t *h_luma[2] = { nullptr, nullptr };   // pinned staging ring for pictures handed over in pageable memory
    cudaEvent_t h_luma_ev[2] = {};   // ... each guarded by the event of its last copy
    unsigned int h_luma_next = 0;
    cudaStream_t mt_stream = nullptr;    // MB-tree stream (propagate / finish / their read-backs)
    cudaEvent_t ev_mt_dep = nullptr;
    // checkpoints on the MB-tree stream: an upload into a slot waits for the first one recorded after the last MB-tree operation
    // that touched the slot (long complete in steady state), not for whatever MB-tree work happens to be queued
    cudaEvent_t ev_mt_ckpt[32] = {};
    unsigned long long mt_ckpt_seq[32] = {}, mt_seq = 0;
    unsigned int mt_ckpt_next = 0;
    cudaStream_t xch_stream = nullptr;   // exchange stream: export / import of search results between GPUs (sharded stream)
    cudaStream_t up_stream = nullptr;    // upload stream: H2D copy, lowres planes and slot reset of a queued picture run beside the analysis
    cudaEvent_t ev_up_guard = nullptr;   // main-stream work queued before a put (it may still read the slot's previous picture)
#define LA_ZC_DEPTH 32
    cudaEvent_t ev_zero_copy[LA_ZC_DEPTH] = {};  // the last copies that read the caller's own page-locked buffers (ring)
    bool zero_copy_live[LA_ZC_DEPTH] = {};
    unsigned int zc_next = 0;
    int async_upload = 0;                // x264cu_lookahead_set_async_upload: page-locked uploads in flight (0 = wait for each)
    bool search_attr_set = false;
    struct XchMark { int slot, list, dm1; };
    std::vector<XchMark> xch_marks;      // searches imported since the last x264cu_lookahead_import_done
    int32_t *d_record = nullptr, *h_record = nullptr;
    LaJobPack pack;                  // jobs being assembled for the next launch
    cudaStream_t search_streams[2] = {};     // prefetch launches alternate: the drain of one wavefront overlaps the fill of the next
    cudaStream_t search_stream = nullptr;     // the one the current search_batch call uses
    unsigned int batch_no = 0;
    int last_ev_of[2] = { -1, -1 };
    cudaEvent_t ev[64];
    unsigned long long ev_seq[64] = {}, ev_seq_next = 1;    // which recording each ring entry currently holds
    int ev_next = 0, n_ev = 0;
    cudaEvent_t ev_main = nullptr;
    int last_prefetch_ev = -1;
    uint16_t *h_qscale = nullptr;
    LaHostStats st;
    bool stats_on = false;
    cudaEvent_t tm_ev[64][2] = {};       // X264CU_STATS: device time of each search launch
    bool tm_live[64] = {};
    double search_busy_ms = 0;
    uint16_t *d_qscale_flat = nullptr;   // mb_count x 256: the factors of a picture without AQ
    cudaEvent_t qs_ev[4] = {};           // guards of the pinned qscale staging ring
    unsigned int qs_next = 0;
    unsigned int *d_tickets = nullptr;   // work-distribution counters of the search launches (ring of 64)
    unsigned int ticket_next = 0;
    uint8_t *d_weight_plane = nullptr;   // h->mb.p_weight_buf[0]: weighted copy of one reference F plane (padded)
    // speculative cost requests: a ring of batches, each with its request descriptors and result records on both sides
#define LA_SPEC_RING 128
#define LA_SPEC_MAX 4096             /* requests per batch */
    cudaStream_t spec_stream = nullptr;
    LaFinalizeArgs *h_spec_args2[2] = {}, *d_spec_args2[2] = {};      // [LA_SPEC_MAX] each: descriptor staging, alternating per batch
    cudaEvent_t spec_args_ev2[2] = {};                                // ... and the kernel that last read each
    cudaEvent_t spec_done_ev = nullptr;                               // sharded: a batch's results are in the send buffer
    int32_t *d_spec_rec = nullptr, *h_spec_rec = nullptr;             // [LA_SPEC_RING][LA_SPEC_MAX][8]
    cudaEvent_t spec_ev[LA_SPEC_RING] = {};                           // batch complete, records on the host
    unsigned long long spec_seq[LA_SPEC_RING] = {}, spec_next = 1;    // which batch each ring entry holds
    long spec_hits = 0, spec_misses = 0, spec_launched = 0;
    unsigned long long *h_stats = nullptr;
};

static int la_stride_lowres( int wl )
{
    int s = ( wl + 96 + 63 ) & ~63;          // align_stride( width + PADH2, 64, 2048 ), common/frame.c:30-36, :125
    if( !( s & 2047 ) ) s += 64;
    return s;
}

extern "C" void x264cu_lookahead_close_internal( x264cu_ctx *ctx ) { (void)ctx; }
x264cu_ctx *x264cu_lookahead_ctx( x264cu_lookahead *la ) { return la->ctx; }

extern "C" {

void x264cu_lookahead_close( x264cu_lookahead_t *la )
{
    if( !la ) return;
    cudaSetDevice( la->ctx->device );
    cudaStreamSynchronize( la->ctx->stream );
    if( la->up_stream ) cudaStreamSynchronize( la->up_stream );
    if( la->xch_stream ) cudaStreamSynchronize( la->xch_stream );
    if( la->mt_stream ) cudaStreamSynchronize( la->mt_stream );
    if( la->spec_stream ) cudaStreamSynchronize( la->spec_stream );
    for( int i = 0; i < 2; i++ ) if( la->search_streams[i] ) cudaStreamSynchronize( la->search_streams[i] );
    if( la->stats_on )
    {
        cudaDeviceSynchronize();
        for( int i = 0; i < 64; i++ )
            if( la->tm_live[i] )
            {
                float ms = 0;
                cudaEventElapsedTime( &ms, la->tm_ev[i][0], la->tm_ev[i][1] );
                la->search_busy_ms += ms;
            }
        fprintf( stderr, "x264cu lookahead: search kernels busy %.1f ms on the device\n", la->search_busy_ms );
    }
    if( la->stats_on )
        fprintf( stderr, "x264cu lookahead host waits: frame_put %.1f ms / %ld, frame_cost %.1f ms / %ld, weights %.1f ms / %ld, event ring %.1f ms; "
                 "%ld search launches, %ld searches\n", la->st.put_sync * 1e3, la->st.n_put, la->st.cost_sync * 1e3, la->st.n_cost,
                 la->st.weight_sync * 1e3, la->st.n_weight, la->st.ev_sync * 1e3, la->st.n_batch, la->st.n_jobs );
    for( auto &s : la->slots )
    {
        cudaFree( s.plane_buf ); cudaFree( s.dev.mvs ); cudaFree( s.dev.mv_costs ); cudaFree( s.dev.costs );
        cudaFree( s.dev.intra ); cudaFree( s.dev.qscale ); cudaFree( s.dev.row_satds ); cudaFree( s.dev.recs ); cudaFree( s.d_stats );
        cudaFree( s.dev.propagate ); cudaFree( s.dev.qp_offset ); cudaFree( s.dev.qp_offset_aq );
    }
    cudaFree( la->d_chroma ); cudaFree( la->d_aq_q4 );
    cudaFree( la->d_cost_mv ); cudaFree( la->d_luma ); cudaFree( la->d_record ); cudaFree( la->d_weight_plane ); cudaFree( la->d_tickets );
    cudaFreeHost( la->h_stats );
    {
        auto &ax = la->ctx->aux_streams;
        for( size_t i = 0; i < ax.size(); )
            if( ax[i] == la->up_stream || ax[i] == la->xch_stream || ax[i] == la->mt_stream || ax[i] == la->spec_stream || ax[i] == la->search_streams[0] || ax[i] == la->search_streams[1] ) ax.erase( ax.begin() + i ); else i++;
    }
    if( la->up_stream ) { cudaStreamSynchronize( la->up_stream ); cudaStreamDestroy( la->up_stream ); }
    if( la->xch_stream ) { cudaStreamSynchronize( la->xch_stream ); cudaStreamDestroy( la->xch_stream ); }
    if( la->mt_stream ) { cudaStreamSynchronize( la->mt_stream ); cudaStreamDestroy( la->mt_stream ); }
    if( la->spec_stream ) { cudaStreamSynchronize( la->spec_stream ); cudaStreamDestroy( la->spec_stream ); }
    cudaFreeHost( la->h_spec_rec ); cudaFree( la->d_spec_rec );
    for( int i = 0; i < 2; i++ )
    {
        cudaFreeHost( la->h_spec_args2[i] ); cudaFree( la->d_spec_args2[i] );
        if( la->spec_args_ev2[i] ) cudaEventDestroy( la->spec_args_ev2[i] );
    }
    if( la->spec_done_ev ) cudaEventDestroy( la->spec_done_ev );
    for( int i = 0; i < LA_SPEC_RING; i++ ) if( la->spec_ev[i] ) cudaEventDestroy( la->spec_ev[i] );
    if( la->ev_mt_dep ) cudaEventDestroy( la->ev_mt_dep );
    for( int i = 0; i < 32; i++ ) if( la->ev_mt_ckpt[i] ) cudaEventDestroy( la->ev_mt_ckpt[i] );
    for( int i = 0; i < 2; i++ ) { cudaFreeHost( la->h_luma[i] ); if( la->h_luma_ev[i] ) cudaEventDestroy( la->h_luma_ev[i] ); }
    if( la->ev_up_guard ) cudaEventDestroy( la->ev_up_guard );
    for( int i = 0; i < LA_ZC_DEPTH; i++ ) if( la->ev_zero_copy[i] ) cudaEventDestroy( la->ev_zero_copy[i] );
    for( auto &s : la->slots ) if( s.ev_ready ) cudaEventDestroy( s.ev_ready );
    cudaFreeHost( la->h_record ); cudaFreeHost( la->h_qscale );
    cudaFree( la->d_qscale_flat );
    for( int i = 0; i < 4; i++ ) if( la->qs_ev[i] ) cudaEventDestroy( la->qs_ev[i] );
    for( int i = 0; i < 2; i++ )
        if( la->search_streams[i] ) { cudaStreamSynchronize( la->search_streams[i] ); cudaStreamDestroy( la->search_streams[i] ); }
    for( int i = 0; i < la->n_ev; i++ ) cudaEventDestroy( la->ev[i] );
    for( int i = 0; i < 64; i++ )
        for( int k = 0; k < 2; k++ ) if( la->tm_ev[i][k] ) cudaEventDestroy( la->tm_ev[i][k] );
    if( la->ev_main ) cudaEventDestroy( la->ev_main );
    delete la;
}

int x264cu_lookahead_open( x264cu_ctx_t *ctx, const x264cu_lookahead_params_t *p, x264cu_lookahead_t **out )
{
    X264CU_ENTER( ctx );
    if( !ctx || !p || !out ) return -1;
    *out = nullptr;
    if( p->width < 16 || p->height < 16 ) return x264cu_fail( ctx, "lookahead_open: picture %dx%d too small", p->width, p->height );
    if( p->bframes < 0 || p->bframes > LA_MAX_B ) return x264cu_fail( ctx, "lookahead_open: bframes %d out of range", p->bframes );
    if( p->me_method < X264CU_ME_DIA || p->me_method > 4 )          // esa / tesa included: the lookahead never goes beyond hex (slicetype.c:50)
        return x264cu_fail( ctx, "lookahead_open: me_method %d out of range", p->me_method );
    if( p->n_slots < 2 ) return x264cu_fail( ctx, "lookahead_open: need at least 2 frame slots" );
    if( p->mv_range < 32 || p->mv_range > 8192 ) return x264cu_fail( ctx, "lookahead_open: mv_range %d out of range", p->mv_range );
    x264cu_lookahead *la = new x264cu_lookahead;
    la->ctx = ctx; la->p = *p;
#ifdef X264CU_TUNING
    la->stats_on = getenv( "X264CU_STATS" ) != nullptr;
#endif
    LaDims &d = la->d;
    d.mb_w = ( p->width + 15 ) >> 4; d.mb_h = ( p->height + 15 ) >> 4; d.mb_count = d.mb_w * d.mb_h;
    la->wl = d.mb_w * 8; la->ll = d.mb_h * 8;
    d.stride = la_stride_lowres( la->wl );
    d.B = p->bframes;
    d.subme = p->subpel_refine;
    if( p->subpel_refine > 1 ) { d.me_method = p->me_method < X264CU_ME_HEX ? p->me_method : X264CU_ME_HEX; d.subpel = 4; }
    else { d.me_method = X264CU_ME_DIA; d.subpel = 2; }
    d.me_range = p->me_range;
    d.mv_range2 = 2 * p->mv_range;
    d.bipred_weighted = p->weighted_bipred; d.aq = p->aq_mode != 0;
    d.do_edges = p->mb_tree || p->vbv || d.mb_w <= 2 || d.mb_h <= 2;
    d.cost_len = 2 * 4 * p->mv_range;
    la->plane_bytes = (size_t)d.stride * ( la->ll + 2 * X264CU_PAD ) + 256;

    cudaSetDevice( ctx->device );
    bool ok = true;
    auto alloc = [&]( void **ptr, size_t n ) { if( cudaMalloc( ptr, n ) != cudaSuccess ) ok = false; else cudaMemsetAsync( *ptr, 0, n, ctx->stream ); };
    la->slots.resize( p->n_slots );
    const int B1 = d.B + 1, B2 = ( d.B + 2 ) * ( d.B + 2 );
    for( auto &s : la->slots )
    {
        memset( &s.dev, 0, sizeof( s.dev ) );
        s.plane_buf = nullptr;
        alloc( (void **)&s.plane_buf, 4 * la->plane_bytes );
        if( !ok ) break;
        for( int i = 0; i < 4; i++ )
            s.dev.planes[i] = s.plane_buf + i * la->plane_bytes + (size_t)X264CU_PAD * d.stride + X264CU_PAD + 32;
        alloc( (void **)&s.dev.mvs, (size_t)2 * B1 * d.mb_count * 4 + 64 );
        alloc( (void **)&s.dev.mv_costs, (size_t)2 * B1 * d.mb_count * 4 );
        alloc( (void **)&s.dev.costs, (size_t)B2 * d.mb_count * 2 );
        alloc( (void **)&s.dev.intra, (size_t)d.mb_count * 4 );
        alloc( (void **)&s.dev.qscale, (size_t)d.mb_count * 2 );
        alloc( (void **)&s.dev.row_satds, (size_t)B2 * d.mb_h * 4 );
        alloc( (void **)&s.dev.recs, (size_t)2 * B1 * d.mb_count * 8 );
        alloc( (void **)&s.dev.propagate, (size_t)d.mb_count * 4 );
        alloc( (void **)&s.dev.qp_offset, (size_t)d.mb_count * 4 );
        alloc( (void **)&s.dev.qp_offset_aq, (size_t)d.mb_count * 4 );
        s.d_stats = nullptr;
        alloc( (void **)&s.d_stats, 16 );
        if( ok && cudaEventCreateWithFlags( &s.ev_ready, cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    }
    la->luma_bytes = (size_t)( ( p->width + 63 ) & ~63 ) * p->height + 64;
    alloc( (void **)&la->d_luma, la->luma_bytes );
    {   // two chroma planes at a stride padded to 64 bytes (la_put_host)
        const size_t cw = ( p->width + 1 ) >> 1, ch = ( p->height + 1 ) >> 1;
        alloc( (void **)&la->d_chroma, 2 * ( ( cw + 63 ) & ~(size_t)63 ) * ch + 256 );
    }
    alloc( (void **)&la->d_aq_q4, (size_t)( d.mb_count + 2 ) * 4 );
    alloc( (void **)&la->d_cost_mv, ( 2 * d.cost_len + 1 ) * 2 + 16 );
    alloc( (void **)&la->d_record, 64 );
    alloc( (void **)&la->d_weight_plane, la->plane_bytes );
    alloc( (void **)&la->d_tickets, 64 * 4 );
    if( ok && cudaMallocHost( (void **)&la->h_stats, 16 * (size_t)p->n_slots ) != cudaSuccess ) ok = false;   // {sum, sqr} per slot
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange( &prio_lo, &prio_hi );
    for( int i = 0; i < 2; i++ )        // background work: lowest priority (the context's stream has the highest)
        if( cudaStreamCreateWithPriority( &la->search_streams[i], cudaStreamNonBlocking, prio_lo ) != cudaSuccess ) ok = false;
    la->search_stream = la->search_streams[0];
    for( int i = 0; ok && i < 64; i++ )
    {
        if( cudaEventCreateWithFlags( &la->ev[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false; else la->n_ev++;
    }
    if( ok && cudaEventCreateWithFlags( &la->ev_main, cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    for( int i = 0; i < 2; i++ )
    {
        if( ok && cudaMallocHost( (void **)&la->h_luma[i], la->luma_bytes ) != cudaSuccess ) ok = false;
        if( ok && cudaEventCreateWithFlags( &la->h_luma_ev[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    }
    if( ok && cudaStreamCreateWithPriority( &la->up_stream, cudaStreamNonBlocking, prio_hi ) != cudaSuccess ) ok = false;
    if( ok && cudaStreamCreateWithPriority( &la->xch_stream, cudaStreamNonBlocking, prio_hi ) != cudaSuccess ) ok = false;
    if( ok && cudaStreamCreateWithPriority( &la->mt_stream, cudaStreamNonBlocking, prio_hi ) != cudaSuccess ) ok = false;
    if( ok && cudaStreamCreateWithPriority( &la->spec_stream, cudaStreamNonBlocking, prio_lo ) != cudaSuccess ) ok = false;
    for( int i = 0; i < 2; i++ )
    {
        if( ok && cudaMallocHost( (void **)&la->h_spec_args2[i], sizeof( LaFinalizeArgs ) * LA_SPEC_MAX ) != cudaSuccess ) ok = false;
        alloc( (void **)&la->d_spec_args2[i], sizeof( LaFinalizeArgs ) * LA_SPEC_MAX );
        if( ok && cudaEventCreateWithFlags( &la->spec_args_ev2[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    }
    if( ok && cudaMallocHost( (void **)&la->h_spec_rec, (size_t)LA_SPEC_RING * LA_SPEC_MAX * 32 ) != cudaSuccess ) ok = false;
    alloc( (void **)&la->d_spec_rec, (size_t)LA_SPEC_RING * LA_SPEC_MAX * 32 );
    if( ok && cudaEventCreateWithFlags( &la->spec_done_ev, cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    for( int i = 0; i < LA_SPEC_RING; i++ )
        if( ok && cudaEventCreateWithFlags( &la->spec_ev[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    if( ok && cudaEventCreateWithFlags( &la->ev_mt_dep, cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    for( int i = 0; i < 32; i++ )
        if( ok && cudaEventCreateWithFlags( &la->ev_mt_ckpt[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    if( ok && cudaEventCreateWithFlags( &la->ev_up_guard, cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    for( int i = 0; i < LA_ZC_DEPTH; i++ )
        if( ok && cudaEventCreateWithFlags( &la->ev_zero_copy[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    if( ok && cudaMallocHost( (void **)&la->h_record, 64 ) != cudaSuccess ) ok = false;
    if( ok && cudaMallocHost( (void **)&la->h_qscale, (size_t)4 * d.mb_count * 2 ) != cudaSuccess ) ok = false;
    for( int i = 0; i < 4 && ok; i++ )
        if( cudaEventCreateWithFlags( &la->qs_ev[i], cudaEventDisableTiming ) != cudaSuccess ) ok = false;
    alloc( (void **)&la->d_qscale_flat, (size_t)d.mb_count * 2 );
    if( ok )
    {
        std::vector<uint16_t> flat( d.mb_count, 256 );
        // same stream as alloc()'s clearing memset, so it lands after it
        if( cudaMemcpyAsync( la->d_qscale_flat, flat.data(), (size_t)d.mb_count * 2, cudaMemcpyHostToDevice, ctx->stream ) != cudaSuccess ||
            cudaStreamSynchronize( ctx->stream ) != cudaSuccess ) ok = false;
    }
    if( !ok )
    {
        x264cu_fail( ctx, "lookahead_open: out of device / pinned memory" );
        x264cu_lookahead_close( la );
        return -1;
    }
    // cost_mv[X264_LOOKAHEAD_QP]: lambda = x264_lambda_tab[12] = 1 (analyse.c:143-157, :179-188, tables.c:99).
    // Built on the host with the same float expressions as the reference and uploaded.
    std::vector<uint16_t> tab( 2 * d.cost_len + 1 );
    for( int i = 0; i <= d.cost_len; i++ )
    {
        float l = i ? log2f( (float)( i + 1 ) ) * 2.0f + 1.718f : 0.718f;
        int c = (int)( 1 * l + .5f );
        if( c > 65535 ) c = 65535;
        tab[d.cost_len + i] = tab[d.cost_len - i] = (uint16_t)c;
    }
    if( cudaMemcpyAsync( la->d_cost_mv, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice, ctx->stream ) != cudaSuccess ||
        cudaStreamSynchronize( ctx->stream ) != cudaSuccess )
    {
        x264cu_fail( ctx, "lookahead_open: cost table upload failed" );
        x264cu_lookahead_close( la );
        return -1;
    }
    {   // x264_log2_lut (common/tables.c:66-85): log2( 1 + i/128 ) printed with five decimals
        float lut[128];
        for( int i = 0; i < 128; i++ )
        {
            char buf[32];
            snprintf( buf, sizeof( buf ), "%.5f", log2( 1.0 + i / 128.0 ) );
            lut[i] = strtof( buf, nullptr );
        }
        if( cudaMemcpyToSymbol( c_log2_lut, lut, sizeof( lut ) ) != cudaSuccess )
        {
            x264cu_fail( ctx, "lookahead_open: log2 table upload failed" );
            x264cu_lookahead_close( la );
            return -1;
        }
    }
    {   // x264_exp2_lut (common/tables.c:58-64): round( 256 * (2^(i/64) - 1) )
        uint8_t e2[64];
        for( int i = 0; i < 64; i++ ) e2[i] = (uint8_t)( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) + 0.5 );
        if( cudaMemcpyToSymbol( c_exp2_lut, e2, sizeof( e2 ) ) != cudaSuccess )
        {
            x264cu_fail( ctx, "lookahead_open: exp2 table upload failed" );
            x264cu_lookahead_close( la );
            return -1;
        }
    }
    ctx->aux_streams.push_back( la->up_stream );
    ctx->aux_streams.push_back( la->xch_stream );
    ctx->aux_streams.push_back( la->mt_stream );
    ctx->aux_streams.push_back( la->spec_stream );
    for( int i = 0; i < 2; i++ ) ctx->aux_streams.push_back( la->search_streams[i] );
    *out = la;
    return 0;
}

static int la_reset_slot( x264cu_lookahead *la, int slot, const uint16_t *h_inv_qscale, cudaStream_t stream )
{
    x264cu_ctx *ctx = la->ctx;
    LaSlotHost &s = la->slots[slot];
    const LaDims &d = la->d;
    // x264_frame_init_lowres's memo reset (mc.c:472-481) + zeroed vectors (frame.c:287-293)
    memset( s.cost_est, -1, sizeof( s.cost_est ) );
    memset( s.cost_est_aq, -1, sizeof( s.cost_est_aq ) );
    memset( s.intra_mbs, 0, sizeof( s.intra_mbs ) );
    memset( s.row_satds_valid, 0, sizeof( s.row_satds_valid ) );
    memset( s.searched, 0, sizeof( s.searched ) );
    memset( s.requested, 0, sizeof( s.requested ) );
    memset( s.pending, -1, sizeof( s.pending ) );
    memset( s.spec, 0, sizeof( s.spec ) );
    s.b_intra_calculated = 0;
    s.gen++;
    s.intra_on_device = false;
    s.in_use = true;
    s.stats_ready = false;
    s.weight.enabled = 0; s.weight.scale = 1; s.weight.denom = 0; s.weight.offset = 0;
    memset( s.weighted_cost_delta, 0, sizeof( s.weighted_cost_delta ) );            // frame.c:798
    {   // without AQ both offset arrays are zero (x264_adaptive_quant_frame, ratecontrol.c:308-330); see frame_set_qp_offset_aq
        CU_CHECK( ctx, cudaMemsetAsync( s.dev.qp_offset, 0, (size_t)d.mb_count * 4, stream ) );
        CU_CHECK( ctx, cudaMemsetAsync( s.dev.qp_offset_aq, 0, (size_t)d.mb_count * 4, stream ) );
    }
    CU_CHECK( ctx, cudaMemsetAsync( s.dev.mvs, 0, (size_t)2 * ( d.B + 1 ) * d.mb_count * 4, stream ) );
    if( h_inv_qscale )
    {   // staged through a ring of pinned buffers: the caller's array may be reused as soon as this returns, and the
        // calling thread never waits for the stream
        const int k = la->qs_next++ & 3;
        uint16_t *stage = la->h_qscale + (size_t)k * d.mb_count;
        LA_TIMED( la->st.put_sync, la->st.n_put, CU_CHECK( ctx, cudaEventSynchronize( la->qs_ev[k] ) ) );
        memcpy( stage, h_inv_qscale, d.mb_count * 2 );
        CU_CHECK( ctx, cudaMemcpyAsync( s.dev.qscale, stage, d.mb_count * 2, cudaMemcpyHostToDevice, stream ) );
        CU_CHECK( ctx, cudaEventRecord( la->qs_ev[k], stream ) );
    }
    else        // no AQ: every factor is 256 (x264_adaptive_quant_frame with aq off, ratecontrol.c:308-330)
        CU_CHECK( ctx, cudaMemcpyAsync( s.dev.qscale, la->d_qscale_flat, d.mb_count * 2, cudaMemcpyDeviceToDevice, stream ) );
    return 0;
}

// Order the upload stream after everything that may still read the picture leaving `slot`: the work queued so far on the
// context's stream (cost requests, on-demand searches) and the last prefetch launch that involved the slot -- not the
// launches in flight for OTHER pictures: an upload must not wait for the searches of the previous group.
static int la_put_begin( x264cu_lookahead *la, int slot )
{
    x264cu_ctx *ctx = la->ctx;
    LaSlotHost &s = la->slots[slot];
    CU_CHECK( ctx, cudaEventRecord( la->ev_up_guard, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamWaitEvent( la->up_stream, la->ev_up_guard, 0 ) );
    if( s.mt_last )
    {   // MB-tree work on the slot's previous picture
        int best = -1;
        for( int i = 0; i < 32; i++ )
            if( la->mt_ckpt_seq[i] >= s.mt_last && ( best < 0 || la->mt_ckpt_seq[i] < la->mt_ckpt_seq[best] ) ) best = i;
        if( best < 0 )
        {
            best = la->mt_ckpt_next++ & 31;
            CU_CHECK( ctx, cudaEventRecord( la->ev_mt_ckpt[best], la->mt_stream ) );
            la->mt_ckpt_seq[best] = la->mt_seq;
        }
        CU_CHECK( ctx, cudaStreamWaitEvent( la->up_stream, la->ev_mt_ckpt[best], 0 ) );
        s.mt_last = 0;
    }
    // a ring entry recorded again since belongs to a launch that had finished by then (search_batch waits before reuse)
    for( int k = 0; k < 4; k++ )
    {
        if( s.last_search_ev[k] >= 0 && la->ev_seq[s.last_search_ev[k]] == s.last_search_seq[k] )
            CU_CHECK( ctx, cudaStreamWaitEvent( la->up_stream, la->ev[s.last_search_ev[k]], 0 ) );
        s.last_search_ev[k] = -1;
    }
    return 0;
}

// lowres planes, luma statistics and memo reset of `slot` from a picture in HBM, all on the upload stream; the slot's
// ev_ready is what its later readers (searches, cost requests, read-backs) are ordered after
struct LaAq { const uint8_t *d_cb, *d_cr; intptr_t cstride; int mode; float strength; };

static int la_put_finish( x264cu_lookahead *la, int slot, const uint8_t *d_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                          const LaAq *aq = nullptr )
{
    x264cu_ctx *ctx = la->ctx;
    LaSlotHost &s = la->slots[slot];
    if( x264cu_frame_init_lowres_on( ctx, la->up_stream, d_luma, luma_stride, la->p.width, la->p.height, s.dev.planes, la->d.stride ) ) return -1;
    if( la->p.weighted_pred )
    {
        CU_CHECK( ctx, cudaMemsetAsync( s.d_stats, 0, 16, la->up_stream ) );
        luma_stats_kernel<<<ctx->sm_count * 2, 256, 0, la->up_stream>>>( d_luma, luma_stride, la->p.width, la->p.height,
                                                                        la->d.mb_w * 16, la->d.mb_h * 16, s.d_stats );
        CU_LAUNCH_CHECK( ctx );
        // read back with the upload: whoever needs the statistics waits for this slot's ev_ready, not for a stream
        CU_CHECK( ctx, cudaMemcpyAsync( la->h_stats + 2 * slot, s.d_stats, 16, cudaMemcpyDeviceToHost, la->up_stream ) );
    }
    if( la_reset_slot( la, slot, h_inv_qscale, la->up_stream ) ) return -1;
    if( aq )
    {   // x264_adaptive_quant_frame (encoder.c:3417) on the device: i_inv_qscale_factor and f_qp_offset(_aq) straight into the slot
        if( x264cu_adaptive_quant_frame_on( ctx, la->up_stream, d_luma, luma_stride, aq->d_cb, aq->d_cr, aq->cstride, la->p.width, la->p.height,
                                            aq->mode, aq->strength, s.dev.qp_offset_aq, s.dev.qscale, la->d_aq_q4, nullptr ) )
            return -1;
        CU_CHECK( ctx, cudaMemcpyAsync( s.dev.qp_offset, s.dev.qp_offset_aq, (size_t)la->d.mb_count * 4, cudaMemcpyDeviceToDevice, la->up_stream ) );
    }
    // the intra costs are a function of the picture alone (slicetype.c:714-757): computed here, off the request path
    intra_kernel<<<( la->d.mb_count + 63 ) / 64, 256, 0, la->up_stream>>>( la->d, s.dev.planes[0], s.dev.intra );
    CU_LAUNCH_CHECK( ctx );
    s.intra_on_device = true;
    CU_CHECK( ctx, cudaEventRecord( s.ev_ready, la->up_stream ) );
    s.main_waited = false;
    return 0;
}

// the context's stream reads `slot` next: order it after the slot's upload (once per picture)
static int la_slot_ready( x264cu_lookahead *la, int slot )
{
    LaSlotHost &s = la->slots[slot];
    if( !s.main_waited )
    {
        CU_CHECK( la->ctx, cudaStreamWaitEvent( la->ctx->stream, s.ev_ready, 0 ) );
        s.main_waited = true;
    }
    return 0;
}

int x264cu_lookahead_frame_put_device( x264cu_lookahead_t *la, int slot, const uint8_t *d_luma, intptr_t luma_stride,
                                       const uint16_t *h_inv_qscale )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    if( slot < 0 || slot >= (int)la->slots.size() ) return x264cu_fail( la->ctx, "frame_put: slot %d out of range", slot );
    // d_luma is complete in the context's stream order (la_put_begin orders the upload stream after that stream)
    if( la_put_begin( la, slot ) ) return -1;
    return la_put_finish( la, slot, d_luma, luma_stride, h_inv_qscale );
}

void x264cu_lookahead_set_async_upload( x264cu_lookahead_t *la, int on )
{
    X264CU_ENTER_LA( la );
    if( la ) la->async_upload = on <= 0 ? 0 : on == 1 ? 4 : on > LA_ZC_DEPTH ? LA_ZC_DEPTH : on;
}

static int la_put_host( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                        const uint8_t *h_cb, const uint8_t *h_cr, intptr_t chroma_stride, int aq_mode, float aq_strength )
{
    if( !la || !h_luma ) return -1;
    x264cu_ctx *ctx = la->ctx;
    if( slot < 0 || slot >= (int)la->slots.size() ) return x264cu_fail( ctx, "frame_put: slot %d out of range", slot );
    const int w = la->p.width, h = la->p.height;
    const intptr_t st = ( w + 63 ) & ~63;
    const int zc = la->zc_next++ % ( la->async_upload ? la->async_upload : 1 );
    if( la->zero_copy_live[zc] )
    {   // the picture queued async_upload calls ago was read in place: that copy ended long ago in steady state
        LA_TIMED( la->st.put_sync, la->st.n_put, CU_CHECK( ctx, cudaEventSynchronize( la->ev_zero_copy[zc] ) ) );
        la->zero_copy_live[zc] = false;
    }
    if( la_put_begin( la, slot ) ) return -1;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes( &at, h_luma ) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if( !pinned ) cudaGetLastError();
    LaAq aq;
    if( h_cb )
    {   // the chroma planes go first: pageable ones are consumed before cudaMemcpy2DAsync returns, page-locked ones fall under
        // the same rule as the luma
        const int cw = ( w + 1 ) >> 1, ch = ( h + 1 ) >> 1;
        const intptr_t cst = ( cw + 63 ) & ~63;
        aq.d_cb = la->d_chroma; aq.d_cr = la->d_chroma + (size_t)cst * ch; aq.cstride = cst; aq.mode = aq_mode; aq.strength = aq_strength;
        CU_CHECK( ctx, cudaMemcpy2DAsync( (void *)aq.d_cb, cst, h_cb, chroma_stride, cw, ch, cudaMemcpyHostToDevice, la->up_stream ) );
        CU_CHECK( ctx, cudaMemcpy2DAsync( (void *)aq.d_cr, cst, h_cr, chroma_stride, cw, ch, cudaMemcpyHostToDevice, la->up_stream ) );
    }
    if( pinned )
    {   // page-locked source (x264cu_malloc_host, like the reference's pinned page-locked staging, opencl.h:718): the DMA
        // engine reads it in place, the calling thread neither copies nor waits
        if( luma_stride == st && w == st )
            CU_CHECK( ctx, cudaMemcpyAsync( la->d_luma, h_luma, (size_t)st * h, cudaMemcpyHostToDevice, la->up_stream ) );
        else
            CU_CHECK( ctx, cudaMemcpy2DAsync( la->d_luma, st, h_luma, luma_stride, w, h, cudaMemcpyHostToDevice, la->up_stream ) );
        CU_CHECK( ctx, cudaEventRecord( la->ev_zero_copy[zc], la->up_stream ) );
        if( la->async_upload ) la->zero_copy_live[zc] = true;
        else LA_TIMED( la->st.put_sync, la->st.n_put, CU_CHECK( ctx, cudaEventSynchronize( la->ev_zero_copy[zc] ) ) );
    }
    else
    {   // pageable source: staged through a ring of two pinned buffers, so that the caller's buffer is free on return
        const int k = la->h_luma_next++ & 1;
        LA_TIMED( la->st.put_sync, la->st.n_put, CU_CHECK( ctx, cudaEventSynchronize( la->h_luma_ev[k] ) ) );
        for( int y = 0; y < h; y++ )
            memcpy( la->h_luma[k] + y * st, h_luma + y * luma_stride, w );
        CU_CHECK( ctx, cudaMemcpyAsync( la->d_luma, la->h_luma[k], (size_t)st * h, cudaMemcpyHostToDevice, la->up_stream ) );
        CU_CHECK( ctx, cudaEventRecord( la->h_luma_ev[k], la->up_stream ) );
        if( h_cb )      // chroma planes in page-locked memory beside a pageable luma: their copies must be over before the caller goes on
            LA_TIMED( la->st.put_sync, la->st.n_put, CU_CHECK( ctx, cudaEventSynchronize( la->h_luma_ev[k] ) ) );
    }
    return la_put_finish( la, slot, la->d_luma, st, h_inv_qscale, h_cb ? &aq : nullptr );
}

int x264cu_lookahead_frame_put( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride,
                                const uint16_t *h_inv_qscale )
{
    X264CU_ENTER_LA( la );
    return la_put_host( la, slot, h_luma, luma_stride, h_inv_qscale, nullptr, nullptr, 0, 0, 0.f );
}

int x264cu_lookahead_frame_put_i420( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride,
                                     const uint8_t *h_cb, const uint8_t *h_cr, intptr_t chroma_stride, int aq_mode, float aq_strength )
{
    X264CU_ENTER_LA( la );
    if( !la || !h_cb || !h_cr ) return -1;
    if( aq_mode < 0 || aq_mode > 3 ) return x264cu_fail( la->ctx, "frame_put_i420: aq-mode %d out of range", aq_mode );
    return la_put_host( la, slot, h_luma, luma_stride, nullptr, h_cb, h_cr, chroma_stride, aq_mode, aq_strength );
}

// enqueue the n searches assembled in la->pack as one launch on `stream` (no host synchronisation)
static int la_launch_searches( x264cu_lookahead *la, int n, cudaStream_t stream )
{
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    if( n <= 0 ) return 0;
    constexpr int NW = LA_NW;
    const int rows = d.mb_h - ( d.do_edges ? 0 : 2 );
    if( rows <= 0 ) return 0;
    const int groups = ( rows + NW - 1 ) / NW;
    size_t smem = (size_t)NW * LA_WIN_BYTES;
    if( !la->search_attr_set )                       // per device: one lookahead object belongs to one device
    {
        CU_CHECK( ctx, cudaFuncSetAttribute( search_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
        CU_CHECK( ctx, cudaFuncSetAttribute( search_kernel<NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared ) );
        la->search_attr_set = true;
    }
    // ticket counters: a ring, one per launch, re-zeroed in stream order before use
    unsigned int *t = la->d_tickets + ( la->ticket_next++ & 63 );
    CU_CHECK( ctx, cudaMemsetAsync( t, 0, 4, stream ) );
    la->st.n_batch++; la->st.n_jobs += n;
    const int ti = ( la->ticket_next - 1 ) & 63;
    {
        if( !la->tm_ev[ti][0] ) { cudaEventCreate( &la->tm_ev[ti][0] ); cudaEventCreate( &la->tm_ev[ti][1] ); }
        if( la->tm_live[ti] )
        {
            float ms = 0;
            cudaEventSynchronize( la->tm_ev[ti][1] );
            cudaEventElapsedTime( &ms, la->tm_ev[ti][0], la->tm_ev[ti][1] );
            la->search_busy_ms += ms;
        }
        cudaEventRecord( la->tm_ev[ti][0], stream );
    }
    search_kernel<NW><<<groups * n, NW * 32, smem, stream>>>( d, la->pack, la->d_cost_mv, n, t );
    CU_LAUNCH_CHECK( ctx );
    cudaEventRecord( la->tm_ev[ti][1], stream ); la->tm_live[ti] = true;
    return 0;
}

static void la_fill_job( x264cu_lookahead *la, LaSearchJob &j, int fenc_slot, int ref_slot, int list, int dist )
{
    const LaDims &d = la->d;
    LaSlotHost &f = la->slots[fenc_slot], &r = la->slots[ref_slot];
    const size_t idx = (size_t)list * ( d.B + 1 ) + ( dist - 1 );
    j.fenc = f.dev.planes[0];
    for( int i = 0; i < 4; i++ ) j.ref[i] = r.dev.planes[i];
    j.mvs = f.dev.mvs + idx * d.mb_count * 2;
    j.mv_costs = f.dev.mv_costs + idx * d.mb_count;
    j.recs = f.dev.recs + idx * d.mb_count;
    j.gen = f.gen;
    j.ref_w = nullptr;
    j.w_enabled = 0; j.w_scale = 1; j.w_denom = 0; j.w_offset = 0;
}

// make the main stream wait for a prefetched search that may still be running on the search stream
static int la_wait_pending( x264cu_lookahead *la, LaSlotHost &s, int list, int dm1 )
{
    int e = s.pending[list][dm1];
    if( e >= 0 )
    {
        CU_CHECK( la->ctx, cudaStreamWaitEvent( la->ctx->stream, la->ev[e], 0 ) );
        s.pending[list][dm1] = -1;
    }
    return 0;
}

int x264cu_lookahead_search_batch( x264cu_lookahead_t *la, int n_jobs, const int *fenc, const int *ref, const int *list, const int *dist )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    x264cu_ctx *ctx = la->ctx;
    // everything queued so far on the main stream (lowres planes, vector resets) must be visible to the searches
    CU_CHECK( ctx, cudaEventRecord( la->ev_main, ctx->stream ) );
    const int si = la->batch_no++ & 1;
    la->search_stream = la->search_streams[si];
    CU_CHECK( ctx, cudaStreamWaitEvent( la->search_stream, la->ev_main, 0 ) );
    int n = 0, launched = 0;
    struct Mark { int slot, list, dm1; };
    std::vector<Mark> marks;
    {   // the pictures of these jobs may still be on their way in on the upload stream
        std::vector<char> seen( la->slots.size(), 0 );
        for( int i = 0; i < n_jobs; i++ )
            for( int sl : { fenc[i], ref[i] } )
                if( sl >= 0 && sl < (int)la->slots.size() && !seen[sl] )
                {
                    seen[sl] = 1;
                    CU_CHECK( ctx, cudaStreamWaitEvent( la->search_stream, la->slots[sl].ev_ready, 0 ) );
                }
    }
    for( int i = 0; i < n_jobs; i++ )
    {
        if( fenc[i] < 0 || fenc[i] >= (int)la->slots.size() || ref[i] < 0 || ref[i] >= (int)la->slots.size() ||
            list[i] < 0 || list[i] > 1 || dist[i] < 1 || dist[i] > la->d.B + 1 )
            return x264cu_fail( ctx, "search_batch: bad job %d", i );
        if( list[i] == 1 && la->d.B == 0 ) return x264cu_fail( ctx, "search_batch: list 1 without b-frames" );
        LaSlotHost &f = la->slots[fenc[i]];
        if( !f.in_use || !la->slots[ref[i]].in_use ) return x264cu_fail( ctx, "search_batch: empty slot in job %d", i );
        if( f.searched[list[i]][dist[i] - 1] ) continue;
        f.searched[list[i]][dist[i] - 1] = true;
        if( n == LA_PACK )
        {
            if( la_launch_searches( la, n, la->search_stream ) ) return -1;
            n = 0; launched = 1;
        }
        la_fill_job( la, la->pack.j[n], fenc[i], ref[i], list[i], dist[i] );
        marks.push_back( Mark{ fenc[i], list[i], dist[i] - 1 } );
        n++;
    }
    if( la_launch_searches( la, n, la->search_stream ) ) return -1;
    if( n || launched )
    {
        const int e = la->ev_next;
        la->ev_next = ( la->ev_next + 1 ) % la->n_ev;
        { long dummy = 0; LA_TIMED( la->st.ev_sync, dummy, CU_CHECK( ctx, cudaEventSynchronize( la->ev[e] ) ) ); }   // ring slot reuse: its previous recording is long done
        CU_CHECK( ctx, cudaEventRecord( la->ev[e], la->search_stream ) );
        la->ev_seq[e] = la->ev_seq_next++;
        for( auto &m : marks ) la->slots[m.slot].pending[m.list][m.dm1] = e;
        for( int i = 0; i < n_jobs; i++ )
            for( int sl : { fenc[i], ref[i] } ) { la->slots[sl].last_search_ev[si] = e; la->slots[sl].last_search_seq[si] = la->ev_seq[e]; }
        la->last_prefetch_ev = e;
        la->last_ev_of[si] = e;
    }
    return 0;
}

int x264cu_lookahead_search_stats( x264cu_lookahead_t *la, double *busy_ms, long *launches, long *searches )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    for( int i = 0; i < 64; i++ )
        if( la->tm_live[i] )
        {   // waits for the launches still in flight
            float ms = 0;
            CU_CHECK( la->ctx, cudaEventSynchronize( la->tm_ev[i][1] ) );
            CU_CHECK( la->ctx, cudaEventElapsedTime( &ms, la->tm_ev[i][0], la->tm_ev[i][1] ) );
            la->search_busy_ms += ms;
            la->tm_live[i] = false;
        }
    if( busy_ms ) *busy_ms = la->search_busy_ms;
    if( launches ) *launches = la->st.n_batch;
    if( searches ) *searches = la->st.n_jobs;
    return 0;
}

int x264cu_lookahead_join( x264cu_lookahead_t *la )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    for( int i = 0; i < 2; i++ )
        if( la->last_ev_of[i] >= 0 )
            CU_CHECK( la->ctx, cudaStreamWaitEvent( la->ctx->stream, la->ev[la->last_ev_of[i]], 0 ) );
    return 0;
}

static int ue_bits( unsigned v ) { int n = 0; v++; while( v >> ( n + 1 ) ) n++; return 2*n + 1; }       /* bs_size_ue */
static int se_bits( int v ) { int t = 1 - 2*v; if( t < 0 ) t = 2*v; int n = 0; while( t >> ( n + 1 ) ) n++; return 2*n + 1; }  /* bs_size_se */

static int la_fetch_stats( x264cu_lookahead *la, LaSlotHost &s )
{
    if( s.stats_ready ) return 0;
    x264cu_ctx *ctx = la->ctx;
    const size_t slot = &s - la->slots.data();
    LA_TIMED( la->st.weight_sync, la->st.n_weight, CU_CHECK( ctx, cudaEventSynchronize( s.ev_ready ) ) );
    const unsigned long long sum = la->h_stats[2 * slot], sqr = la->h_stats[2 * slot + 1];
    const unsigned long long N = (unsigned long long)( la->d.mb_w * 16 ) * ( la->d.mb_h * 16 );
    s.pixel_sum = sum;
    s.pixel_ssd = sqr - ( sum * sum + N / 2 ) / N;                   /* ratecontrol.c:405-414 */
    s.stats_ready = true;
    return 0;
}

static int la_weight_cost( x264cu_lookahead *la, LaSlotHost &fenc, LaSlotHost &ref, const LaWeight &w, unsigned *out )
{
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    unsigned int *d_acc = (unsigned int *)( la->d_record + 8 );
    CU_CHECK( ctx, cudaMemsetAsync( d_acc, 0, 4, ctx->stream ) );
    weight_cost_kernel<<<( d.mb_count + 63 ) / 64, 256, 0, ctx->stream>>>( d, fenc.dev.planes[0], ref.dev.planes[0], fenc.dev.intra, w, d_acc );
    CU_LAUNCH_CHECK( ctx );
    CU_CHECK( ctx, cudaMemcpyAsync( la->h_record + 8, d_acc, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    unsigned c = (unsigned)la->h_record[8];
    if( w.enabled )                /* weight_slice_header_cost, slicetype.c:170-189: one slice, lambda 1 */
        c += 10 + ue_bits( w.denom ) * 2 + 2 * ( se_bits( w.scale ) + se_bits( w.offset ) );
    *out = c;
    return 0;
}

/* x264_weights_analyse( h, fenc, ref, b_lookahead = 1 ), slicetype.c:284-501: luma only; the float expressions are the
 * reference's.  On success fenc.weight holds the weight (enabled = 0: none) and la->d_weight_plane the weighted plane. */
static int la_weights_analyse( x264cu_lookahead *la, int fenc_slot, int ref_slot, int delta_index )
{
    x264cu_ctx *ctx = la->ctx;
    LaSlotHost &fenc = la->slots[fenc_slot], &ref = la->slots[ref_slot];
    const LaDims &d = la->d;
    fenc.weight.enabled = 0; fenc.weight.scale = 1; fenc.weight.denom = 0; fenc.weight.offset = 0;
    if( la_fetch_stats( la, fenc ) || la_fetch_stats( la, ref ) ) return -1;
    const float epsilon = 1.f / 128.f;
    const int zero_bias = !ref.pixel_ssd;
    const float fenc_var = fenc.pixel_ssd + zero_bias, ref_var = ref.pixel_ssd + zero_bias;
    const float guess_scale = sqrtf( fenc_var / ref_var );
    const int npix = ( d.mb_h * 16 ) * ( d.mb_w * 16 );
    const float fenc_mean = (float)( fenc.pixel_sum + zero_bias ) / npix;
    const float ref_mean = (float)( ref.pixel_sum + zero_bias ) / npix;
    if( fabsf( ref_mean - fenc_mean ) < 0.5f && fabsf( 1.f - guess_scale ) < epsilon )
        return 0;
    int denom = 7, scale = (int)round( guess_scale * 128 );               /* weight_get_h264, slicetype.c:64-75 */
    while( denom > 0 && scale > 127 ) { denom--; scale >>= 1; }
    if( scale > 127 ) scale = 127;
    int mindenom = denom, minscale = scale, minoff = 0, found = 0;
    if( !fenc.b_intra_calculated )
    {   /* slicetype.c:364-369: an intra-only request on fenc first */
        int one[1] = { fenc_slot }, sc;
        if( x264cu_lookahead_frame_cost( la, one, 0, 0, 0, &sc ) ) return -1;
    }
    LaWeight none; none.enabled = 0; none.scale = 1; none.denom = 0; none.offset = 0;
    unsigned origscore, minscore;
    if( la_weight_cost( la, fenc, ref, none, &origscore ) ) return -1;
    minscore = origscore;
    if( !minscore )
        return 0;
    {   /* scale_dist = offset_dist = 0 in the lookahead: one (scale, offset) pair */
        int cur_scale = minscale;
        int cur_offset = fenc_mean - ref_mean * cur_scale / ( 1 << mindenom ) + 0.5f;
        if( cur_offset < -128 || cur_offset > 127 )
        {
            cur_offset = cur_offset < -128 ? -128 : 127;
            double v = ( 1 << mindenom ) * ( fenc_mean - cur_offset ) / ref_mean + 0.5f;
            cur_scale = (int)( v < 0 ? 0 : v > 127 ? 127 : v );
        }
        LaWeight w; w.enabled = 1; w.scale = cur_scale; w.denom = mindenom; w.offset = cur_offset;
        unsigned s;
        if( la_weight_cost( la, fenc, ref, w, &s ) ) return -1;
        if( s < minscore ) { minscore = s; minscale = cur_scale; minoff = cur_offset; found = 1; }
    }
    while( mindenom > 0 && !( minscale & 1 ) ) { mindenom--; minscale >>= 1; }
    if( !found || ( minscale == 1 << mindenom && minoff == 0 ) || (float)minscore / origscore > 0.998f )
        return 0;
    fenc.weight.enabled = 1; fenc.weight.scale = minscale; fenc.weight.denom = mindenom; fenc.weight.offset = minoff;
    if( la->p.weighted_pred < 0 )                                   /* X264_WEIGHTP_FAKE, slicetype.c:462-463 */
        fenc.weighted_cost_delta[delta_index] = (float)minscore / origscore;
    const size_t words = ( (size_t)d.stride * ( la->ll + 2 * X264CU_PAD ) + 128 ) / 4;
    weight_plane_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>( (const uint32_t *)( ref.plane_buf ), (uint32_t *)la->d_weight_plane, words, fenc.weight );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

static int la_check_slot( x264cu_lookahead *la, int slot );

// MB-tree runs on its own stream, off the chain of work the calling thread waits for in a cost request: it only has to come after
// what is queued on the context's stream so far (the finalize of the triple it propagates, on-demand searches, the slots' uploads)
// after an MB-tree operation has been enqueued: note which slots it touched, drop a checkpoint every 8 operations
static int la_mt_end( x264cu_lookahead *la, std::initializer_list<int> slots )
{
    la->mt_seq++;
    for( int sl : slots ) la->slots[sl].mt_last = la->mt_seq;
    if( !( la->mt_seq & 7 ) )
    {
        const int i = la->mt_ckpt_next++ & 31;
        CU_CHECK( la->ctx, cudaEventRecord( la->ev_mt_ckpt[i], la->mt_stream ) );
        la->mt_ckpt_seq[i] = la->mt_seq;
    }
    return 0;
}

static int la_mt_begin( x264cu_lookahead *la )
{
    CU_CHECK( la->ctx, cudaEventRecord( la->ev_mt_dep, la->ctx->stream ) );
    CU_CHECK( la->ctx, cudaStreamWaitEvent( la->mt_stream, la->ev_mt_dep, 0 ) );
    return 0;
}

/* ---- MB-tree (slicetype.c:1029-1184) ---- */
int x264cu_lookahead_frame_set_qp_offset_aq( x264cu_lookahead_t *la, int slot, const float *h_aq )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) || la_mt_begin( la ) ) return -1;
    x264cu_ctx *ctx = la->ctx;
    LaSlotHost &s = la->slots[slot];
    const size_t n = (size_t)la->d.mb_count * 4;
    if( h_aq )
    {
        CU_CHECK( ctx, cudaMemcpyAsync( s.dev.qp_offset_aq, h_aq, n, cudaMemcpyHostToDevice, la->mt_stream ) );
        CU_CHECK( ctx, cudaStreamSynchronize( la->mt_stream ) );             // the caller's array is free on return
    }
    else
        CU_CHECK( ctx, cudaMemsetAsync( s.dev.qp_offset_aq, 0, n, la->mt_stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( s.dev.qp_offset, s.dev.qp_offset_aq, n, cudaMemcpyDeviceToDevice, la->mt_stream ) );
    return la_mt_end( la, { slot } );
}

int x264cu_lookahead_mbtree_reset( x264cu_lookahead_t *la, int slot )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use ) return x264cu_fail( la->ctx, "mbtree_reset: empty slot %d", slot );
    if( la_slot_ready( la, slot ) || la_mt_begin( la ) ) return -1;
    CU_CHECK( la->ctx, cudaMemsetAsync( la->slots[slot].dev.propagate, 0, (size_t)la->d.mb_count * 4, la->mt_stream ) );
    return la_mt_end( la, { slot } );
}

int x264cu_lookahead_mbtree_swap( x264cu_lookahead_t *la, int slot_a, int slot_b )
{   // XCHG( uint16_t*, a->i_propagate_cost, b->i_propagate_cost ), slicetype.c:1125, :1177; work already queued keeps its pointers
    if( !la ) return -1;
    for( int s : { slot_a, slot_b } )
        if( s < 0 || s >= (int)la->slots.size() || !la->slots[s].in_use ) return x264cu_fail( la->ctx, "mbtree_swap: empty slot %d", s );
    if( la_slot_ready( la, slot_a ) || la_slot_ready( la, slot_b ) || la_mt_begin( la ) ) return -1;
    std::swap( la->slots[slot_a].dev.propagate, la->slots[slot_b].dev.propagate );
    return la_mt_end( la, { slot_a, slot_b } );
}

int x264cu_lookahead_mbtree_propagate( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int referenced, float fps_factor )
{
    X264CU_ENTER_LA( la );
    if( !la || !frames ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    if( !( p0 < p1 && p0 < b && b <= p1 ) || b - p0 > d.B + 1 || p1 - b > d.B + 1 )
        return x264cu_fail( ctx, "mbtree_propagate: bad frame triple (%d,%d,%d)", p0, p1, b );
    const int sb = frames[b], s0 = frames[p0], s1 = frames[p1];
    for( int s : { sb, s0, s1 } )
        if( s < 0 || s >= (int)la->slots.size() || !la->slots[s].in_use ) return x264cu_fail( ctx, "mbtree_propagate: empty slot %d", s );
    LaSlotHost &fb = la->slots[sb];
    const int i0 = b - p0, i1 = p1 - b;
    if( fb.cost_est[i0][i1] < 0 ) return x264cu_fail( ctx, "mbtree_propagate: the cost (%d,%d,%d) has not been requested", p0, p1, b );
    for( int s : { sb, s0, s1 } )
        if( la_slot_ready( la, s ) ) return -1;
    if( la_mt_begin( la ) ) return -1;
    const int B1 = d.B + 1;
    LaMbtreeArgs A;
    memset( &A, 0, sizeof( A ) );
    const int dist_scale_factor = ( ( i0 << 8 ) + ( ( p1 - p0 ) >> 1 ) ) / ( p1 - p0 );
    const int bw = d.bipred_weighted ? 64 - ( dist_scale_factor >> 2 ) : 32;
    A.bipred_weight[0] = bw; A.bipred_weight[1] = 64 - bw;
    A.ref_costs[0] = la->slots[s0].dev.propagate;
    A.mvs[0] = fb.dev.mvs + (size_t)( 0 * B1 + i0 - 1 ) * d.mb_count * 2;
    if( b != p1 )
    {
        A.ref_costs[1] = la->slots[s1].dev.propagate;
        A.mvs[1] = fb.dev.mvs + (size_t)( 1 * B1 + i1 - 1 ) * d.mb_count * 2;
    }
    A.intra = fb.dev.intra; A.qscale = fb.dev.qscale;
    A.lowres_costs = fb.dev.costs + (size_t)( i0 * ( d.B + 2 ) + i1 ) * d.mb_count;
    A.fps_factor = fps_factor;
    if( referenced ) A.prop_in = fb.dev.propagate;
    else            // slicetype.c:1066-1067: the first row of the frame's own array is cleared and re-used as the all-zero input
        CU_CHECK( ctx, cudaMemsetAsync( fb.dev.propagate, 0, (size_t)d.mb_w * 4, la->mt_stream ) );
    mbtree_propagate_kernel<<<( d.mb_count + 255 ) / 256, 256, 0, la->mt_stream>>>( d, A );
    CU_LAUNCH_CHECK( ctx );
    return la_mt_end( la, { sb, s0, s1 } );
}

int x264cu_lookahead_mbtree_finish( x264cu_lookahead_t *la, int slot, int fps_factor, int ref0_distance, float strength )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    x264cu_ctx *ctx = la->ctx;
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use ) return x264cu_fail( ctx, "mbtree_finish: empty slot %d", slot );
    if( ref0_distance < 0 || ref0_distance > la->d.B + 1 ) return x264cu_fail( ctx, "mbtree_finish: bad distance %d", ref0_distance );
    if( la_slot_ready( la, slot ) || la_mt_begin( la ) ) return -1;
    LaSlotHost &f = la->slots[slot];
    if( !fps_factor )
    {   // lookahead-less intra case (slicetype.c:1121-1123): f_qp_offset = f_qp_offset_aq
        CU_CHECK( ctx, cudaMemcpyAsync( f.dev.qp_offset, f.dev.qp_offset_aq, (size_t)la->d.mb_count * 4, cudaMemcpyDeviceToDevice, la->mt_stream ) );
        return la_mt_end( la, { slot } );
    }
    float weightdelta = 0.0f;
    if( ref0_distance && f.weighted_cost_delta[ref0_distance - 1] > 0 )
        weightdelta = ( 1.0 - f.weighted_cost_delta[ref0_distance - 1] );
    mbtree_finish_kernel<<<( la->d.mb_count + 255 ) / 256, 256, 0, la->mt_stream>>>( la->d.mb_count, f.dev.intra, f.dev.qscale, f.dev.propagate,
                                                                                  f.dev.qp_offset_aq, f.dev.qp_offset, fps_factor, weightdelta, strength );
    CU_LAUNCH_CHECK( ctx );
    return la_mt_end( la, { slot } );
}

int x264cu_lookahead_frame_cost_recalculate( x264cu_lookahead_t *la, int slot, int dist0, int dist1, int b_type, int *h_score, int32_t *h_row_satd )
{
    X264CU_ENTER_LA( la );
    if( !la || !h_score ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use ) return x264cu_fail( ctx, "frame_cost_recalculate: empty slot %d", slot );
    if( dist0 < 0 || dist1 < 0 || dist0 > d.B + 1 || dist1 > d.B + 1 ) return x264cu_fail( ctx, "frame_cost_recalculate: bad distances" );
    LaSlotHost &f = la->slots[slot];
    if( f.cost_est[dist0][dist1] < 0 )
        return x264cu_fail( ctx, "frame_cost_recalculate: the cost (%d,%d) of slot %d has not been requested", dist0, dist1, slot );
    if( la_slot_ready( la, slot ) || la_mt_begin( la ) ) return -1;
    int *d_score = (int *)x264cu_scratch( ctx, 10, 64 );
    if( !d_score ) return -1;
    int32_t *rows = f.dev.row_satds + (size_t)( dist0 * ( d.B + 2 ) + dist1 ) * d.mb_h;
    CU_CHECK( ctx, cudaMemsetAsync( d_score, 0, sizeof(int), la->mt_stream ) );
    recalculate_kernel<<<d.mb_h, 128, 0, la->mt_stream>>>( d.mb_w, d.mb_h, f.dev.costs + (size_t)( dist0 * ( d.B + 2 ) + dist1 ) * d.mb_count,
                                                          !dist0 && !dist1 ? f.dev.intra : nullptr, b_type ? f.dev.qp_offset_aq : f.dev.qp_offset, rows, d_score );
    CU_LAUNCH_CHECK( ctx );
    f.row_satds_valid[dist0][dist1] = true;
    CU_CHECK( ctx, cudaMemcpyAsync( h_score, d_score, sizeof(int), cudaMemcpyDeviceToHost, la->mt_stream ) );
    if( h_row_satd )
        CU_CHECK( ctx, cudaMemcpyAsync( h_row_satd, rows, (size_t)d.mb_h * 4, cudaMemcpyDeviceToHost, la->mt_stream ) );
    if( la_mt_end( la, { slot } ) ) return -1;
    CU_CHECK( ctx, cudaStreamSynchronize( la->mt_stream ) );
    return 0;
}

int x264cu_lookahead_get_qp_offset( x264cu_lookahead_t *la, int slot, float *h_qp_offset )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) || la_mt_begin( la ) ) return -1;
    CU_CHECK( la->ctx, cudaMemcpyAsync( h_qp_offset, la->slots[slot].dev.qp_offset, (size_t)la->d.mb_count * 4, cudaMemcpyDeviceToHost, la->mt_stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->mt_stream ) );
    return 0;
}

int x264cu_lookahead_get_propagate_cost( x264cu_lookahead_t *la, int slot, uint16_t *h_out )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) || la_mt_begin( la ) ) return -1;
    std::vector<unsigned int> tmp( la->d.mb_count );
    CU_CHECK( la->ctx, cudaMemcpyAsync( tmp.data(), la->slots[slot].dev.propagate, (size_t)la->d.mb_count * 4, cudaMemcpyDeviceToHost, la->mt_stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->mt_stream ) );
    for( int i = 0; i < la->d.mb_count; i++ ) h_out[i] = (uint16_t)( tmp[i] > LA_PROPAGATE_MAX ? LA_PROPAGATE_MAX : tmp[i] );
    return 0;
}

float x264cu_lookahead_get_weighted_cost_delta( x264cu_lookahead_t *la, int slot, int dist_minus1 )
{
    X264CU_ENTER_LA( la );
    if( !la || slot < 0 || slot >= (int)la->slots.size() || dist_minus1 < 0 || dist_minus1 > la->d.B ) return -1.0f;
    return la->slots[slot].weighted_cost_delta[dist_minus1];
}

/* ---- one picture stream sharded over several GPUs: the results of a search (lowres_mvs + lowres_mv_costs of one
 * (picture, list, distance)) leave the GPU that ran it / enter the GPUs that did not ---- */
size_t x264cu_lookahead_search_bytes( x264cu_lookahead_t *la ) { return la ? (size_t)la->d.mb_count * 8 : 0; }
void *x264cu_lookahead_exchange_stream( x264cu_lookahead_t *la ) { return la ? (void *)la->xch_stream : nullptr; }

static int la_xch_args( x264cu_lookahead *la, int slot, int list, int dist, const char *who )
{
    X264CU_ENTER_LA( la );
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use || list < 0 || list > 1 || dist < 1 || dist > la->d.B + 1 ||
        ( list == 1 && !la->d.B ) )
        return x264cu_fail( la->ctx, "%s: bad search (slot %d, list %d, distance %d)", who, slot, list, dist );
    return 0;
}

int x264cu_lookahead_export_search( x264cu_lookahead_t *la, int slot, int list, int dist, void *d_dst )
{
    X264CU_ENTER_LA( la );
    if( !la || !d_dst ) return -1;
    if( la_xch_args( la, slot, list, dist, "export_search" ) ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    LaSlotHost &s = la->slots[slot];
    if( !s.searched[list][dist - 1] ) return x264cu_fail( ctx, "export_search: (slot %d, list %d, distance %d) has not been searched here", slot, list, dist );
    const int e = s.pending[list][dist - 1];
    if( e >= 0 ) CU_CHECK( ctx, cudaStreamWaitEvent( la->xch_stream, la->ev[e], 0 ) );       // the search itself, if still in flight
    const size_t idx = (size_t)list * ( d.B + 1 ) + ( dist - 1 );
    CU_CHECK( ctx, cudaMemcpyAsync( d_dst, s.dev.mvs + idx * d.mb_count * 2, (size_t)d.mb_count * 4, cudaMemcpyDeviceToDevice, la->xch_stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( (uint8_t *)d_dst + (size_t)d.mb_count * 4, s.dev.mv_costs + idx * d.mb_count, (size_t)d.mb_count * 4,
                                    cudaMemcpyDeviceToDevice, la->xch_stream ) );
    return 0;
}

int x264cu_lookahead_import_search( x264cu_lookahead_t *la, int slot, int list, int dist, const void *d_src )
{
    X264CU_ENTER_LA( la );
    if( !la || !d_src ) return -1;
    if( la_xch_args( la, slot, list, dist, "import_search" ) ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    LaSlotHost &s = la->slots[slot];
    // the slot's upload (its vector arrays are cleared there) comes first
    CU_CHECK( ctx, cudaStreamWaitEvent( la->xch_stream, s.ev_ready, 0 ) );
    const size_t idx = (size_t)list * ( d.B + 1 ) + ( dist - 1 );
    CU_CHECK( ctx, cudaMemcpyAsync( s.dev.mvs + idx * d.mb_count * 2, d_src, (size_t)d.mb_count * 4, cudaMemcpyDeviceToDevice, la->xch_stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( s.dev.mv_costs + idx * d.mb_count, (const uint8_t *)d_src + (size_t)d.mb_count * 4, (size_t)d.mb_count * 4,
                                    cudaMemcpyDeviceToDevice, la->xch_stream ) );
    s.searched[list][dist - 1] = true;
    s.xch_dirty = true;
    la->xch_marks.push_back( { slot, list, dist - 1 } );
    return 0;
}

int x264cu_lookahead_import_done( x264cu_lookahead_t *la )
{   // one event for everything imported since the last call: cost requests (and the slots' next uploads) wait for it
    if( !la ) return -1;
    x264cu_ctx *ctx = la->ctx;
    if( la->xch_marks.empty() ) return 0;
    const int e = la->ev_next;
    la->ev_next = ( la->ev_next + 1 ) % la->n_ev;
    CU_CHECK( ctx, cudaEventSynchronize( la->ev[e] ) );                  // ring entry reuse
    CU_CHECK( ctx, cudaEventRecord( la->ev[e], la->xch_stream ) );
    la->ev_seq[e] = la->ev_seq_next++;
    for( auto &m : la->xch_marks )
    {
        LaSlotHost &s = la->slots[m.slot];
        s.pending[m.list][m.dm1] = e;
        s.last_search_ev[2] = e; s.last_search_seq[2] = la->ev_seq[e];
    }
    la->xch_marks.clear();
    return 0;
}

int x264cu_lookahead_weight_trivial( x264cu_lookahead_t *la, int fenc_slot, int ref_slot )
{
    X264CU_ENTER_LA( la );
    if( !la ) return -1;
    if( !la->p.weighted_pred ) return 1;
    for( int s : { fenc_slot, ref_slot } )
        if( s < 0 || s >= (int)la->slots.size() || !la->slots[s].in_use ) return x264cu_fail( la->ctx, "weight_trivial: empty slot %d", s );
    LaSlotHost &fenc = la->slots[fenc_slot], &ref = la->slots[ref_slot];
    if( la_fetch_stats( la, fenc ) || la_fetch_stats( la, ref ) ) return -1;
    // the early exit of x264_weights_analyse (slicetype.c:316-330), same float expressions as la_weights_analyse
    const float epsilon = 1.f / 128.f;
    const int zero_bias = !ref.pixel_ssd;
    const float fenc_var = fenc.pixel_ssd + zero_bias, ref_var = ref.pixel_ssd + zero_bias;
    const float guess_scale = sqrtf( fenc_var / ref_var );
    const int npix = ( la->d.mb_h * 16 ) * ( la->d.mb_w * 16 );
    const float fenc_mean = (float)( fenc.pixel_sum + zero_bias ) / npix;
    const float ref_mean = (float)( ref.pixel_sum + zero_bias ) / npix;
    return fabsf( ref_mean - fenc_mean ) < 0.5f && fabsf( 1.f - guess_scale ) < epsilon;
}

/* Cost requests answered ahead of time.  finalize (the bidir candidates, the list / intra choice, the sums) of a triple is a pure
 * function of three pictures' planes and vectors, given which of two variants applies: with or without the temporal-direct
 * vectors of the later reference (slicetype.c:629-642 -- whether that reference's P search had been ASKED FOR by then, which only
 * the host's request order knows).  The decision nearly always asks in an order that has them (path cost: the P cost first), so
 * that variant is computed for every triple a window can ask about as soon as its searches are queued: n triples in ONE launch,
 * their records read back in ONE copy.  x264cu_lookahead_frame_cost then finds the record instead of launching, copying back and
 * waiting per request; a request that needs the other variant, a weighted search or row sums (VBV) takes the on-demand path. */
static int la_finalize_batch( x264cu_lookahead_t *la, int n, const int *b_slot, const int *p0_slot, const int *p1_slot,
                              const int *d0, const int *d1, const int *owner, int rank, int world, x264cu_exchange_fn exchange, void *user )
{
    X264CU_ENTER_LA( la );
    if( !la || ( n > 0 && ( !b_slot || !p0_slot || !p1_slot || !d0 || !d1 ) ) ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    if( n <= 0 || la->p.vbv ) return 0;
    if( n > LA_SPEC_MAX ) n = LA_SPEC_MAX;
    const bool sharded = world > 1;
    // sharded: the triples are split between the ranks (owner[i]); every rank walks the same list and filters it by the same
    // (replicated) state, so all agree on which triples exist, whose they are and where their results sit in the exchange:
    // rank r's block = [most x 32 B records][most x cost_bytes lowres_costs], its k-th triple at index k of both parts
    std::vector<int> rank_count( sharded ? world : 1, 0 );
    struct Kept { int slot, i0, i1, owner, k; };
    std::vector<Kept> kept;
    const int B1 = d.B + 1, B2w = d.B + 2;
    const int ring = (int)( la->spec_next % LA_SPEC_RING );
    // ring entry reuse (its batch is hundreds of pictures old) and the descriptor staging buffer (its last upload)
    CU_CHECK( ctx, cudaEventSynchronize( la->spec_ev[ring] ) );
    const int stg = (int)( la->spec_next & 1 );
    LaFinalizeArgs *h_args = la->h_spec_args2[stg], *d_args = la->d_spec_args2[stg];
    CU_CHECK( ctx, cudaEventSynchronize( la->spec_args_ev2[stg] ) );         // the batch before last: long through
    int32_t *d_rec = la->d_spec_rec + (size_t)ring * LA_SPEC_MAX * 8;
    std::vector<char> ev_seen( la->n_ev, 0 ), slot_seen( la->slots.size(), 0 );
    auto wait_ev = [&]( int e ) -> int {
        if( e >= 0 && !ev_seen[e] ) { ev_seen[e] = 1; CU_CHECK( ctx, cudaStreamWaitEvent( la->spec_stream, la->ev[e], 0 ) ); }
        return 0;
    };
    int m = 0;
    for( int i = 0; i < n; i++ )
    {
        const int sb = b_slot[i], s0 = p0_slot[i], s1 = p1_slot[i], i0 = d0[i], i1 = d1[i];
        for( int sl : { sb, s0, s1 } )
            if( sl < 0 || sl >= (int)la->slots.size() || !la->slots[sl].in_use ) return x264cu_fail( ctx, "finalize_batch: empty slot %d", sl );
        if( i0 < 1 || i1 < 0 || i0 > B1 || i1 > d.B || i0 + i1 > B1 ) return x264cu_fail( ctx, "finalize_batch: bad distances (%d,%d)", i0, i1 );
        LaSlotHost &fenc = la->slots[sb], &f1 = la->slots[s1];
        // only triples whose searches exist (unweighted, launched ahead of time) and that have not been answered yet
        if( fenc.cost_est[i0][i1] >= 0 || fenc.spec[i0][i1].batch || !fenc.searched[0][i0 - 1] ) continue;
        if( i1 && ( !fenc.searched[1][i1 - 1] || !f1.searched[0][i0 + i1 - 1] ) ) continue;
        for( int sl : { sb, s0, s1 } )
            if( !slot_seen[sl] ) { slot_seen[sl] = 1; CU_CHECK( ctx, cudaStreamWaitEvent( la->spec_stream, la->slots[sl].ev_ready, 0 ) ); }
        if( wait_ev( fenc.pending[0][i0 - 1] ) ) return -1;
        if( i1 && ( wait_ev( fenc.pending[1][i1 - 1] ) || wait_ev( f1.pending[0][i0 + i1 - 1] ) ) ) return -1;
        const int own = sharded ? owner[i] : 0;
        if( sharded )
        {
            if( own < 0 || own >= world ) return x264cu_fail( ctx, "finalize_batch: owner %d of %d ranks", own, world );
            kept.push_back( Kept{ sb, i0, i1, own, rank_count[own]++ } );
            fenc.spec[i0][i1].batch = la->spec_next;               // index set once the blocks' size is known
            if( own != rank ) continue;
        }
        LaFinalizeArgs &A = h_args[m];
        memset( &A, 0, sizeof( A ) );
        A.fenc = fenc.dev.planes[0];
        for( int k = 0; k < 4; k++ ) { A.ref0[k] = la->slots[s0].dev.planes[k]; A.ref1[k] = f1.dev.planes[k]; }
        A.b_inter = 1; A.b_bidir = i1 != 0;
        A.mvs0 = fenc.dev.mvs + (size_t)( 0 * B1 + i0 - 1 ) * d.mb_count * 2;
        A.cost0 = fenc.dev.mv_costs + (size_t)( 0 * B1 + i0 - 1 ) * d.mb_count;
        if( i1 )
        {
            A.mvs1 = fenc.dev.mvs + (size_t)( 1 * B1 + i1 - 1 ) * d.mb_count * 2;
            A.cost1 = fenc.dev.mv_costs + (size_t)( 1 * B1 + i1 - 1 ) * d.mb_count;
            A.mvr = f1.dev.mvs + (size_t)( 0 * B1 + i0 + i1 - 1 ) * d.mb_count * 2;
        }
        A.intra = fenc.dev.intra; A.qscale = fenc.dev.qscale;
        A.costs = fenc.dev.costs + (size_t)( i0 * B2w + i1 ) * d.mb_count;
        A.row_inter = nullptr; A.row_intra = nullptr;               // row sums are a VBV matter: on demand
        A.record = d_rec + m * 8;
        A.dist_scale_factor = i1 ? ( ( i0 << 8 ) + ( ( i0 + i1 ) >> 1 ) ) / ( i0 + i1 ) : 128;
        A.bipred_weight = d.bipred_weighted ? 64 - ( A.dist_scale_factor >> 2 ) : 32;
        fenc.spec[i0][i1].batch = la->spec_next; fenc.spec[i0][i1].index = m;
        m++;
    }
    if( !m && kept.empty() ) return 0;
    if( m )
    {
        CU_CHECK( ctx, cudaMemcpyAsync( d_args, h_args, sizeof( LaFinalizeArgs ) * m, cudaMemcpyHostToDevice, la->spec_stream ) );
        CU_CHECK( ctx, cudaMemsetAsync( d_rec, 0, (size_t)m * 32, la->spec_stream ) );
        const dim3 grid( ( d.mb_count + LA_FIN_THREADS / 4 - 1 ) / ( LA_FIN_THREADS / 4 ), m );
        finalize_batch_kernel<<<grid, LA_FIN_THREADS, 0, la->spec_stream>>>( d, d_args );
        CU_LAUNCH_CHECK( ctx );
        CU_CHECK( ctx, cudaEventRecord( la->spec_args_ev2[stg], la->spec_stream ) );
    }
    int32_t *h_rec = la->h_spec_rec + (size_t)ring * LA_SPEC_MAX * 8;
    if( !sharded )
    {
        CU_CHECK( ctx, cudaMemcpyAsync( h_rec, d_rec, (size_t)m * 32, cudaMemcpyDeviceToHost, la->spec_stream ) );
        CU_CHECK( ctx, cudaEventRecord( la->spec_ev[ring], la->spec_stream ) );
    }
    else
    {   // ONE all-gather of every rank's records and lowres_costs (the second exchange of a window; phases 2 / 3 of the callback)
        int most = 0;
        for( int r = 0; r < world; r++ ) most = rank_count[r] > most ? rank_count[r] : most;
        if( world * most > LA_SPEC_MAX ) return x264cu_fail( ctx, "finalize_batch: %d triples on %d ranks exceed the batch size", most, world );
        const size_t cost_bytes = ( (size_t)d.mb_count * 2 + 15 ) & ~(size_t)15, per_rank = (size_t)most * ( 32 + cost_bytes );
        void *d_send = nullptr, *d_recv = nullptr;
        if( exchange( user, 2, per_rank, &d_send, &d_recv, (void *)la->xch_stream ) || !d_send || !d_recv )
            return x264cu_fail( ctx, "finalize_batch: the exchange callback gave no buffers" );
        // my results into my block, behind the kernel -- and behind the previous batch's exchange, which reads the same buffers
        if( la->spec_next > 1 )
            CU_CHECK( ctx, cudaStreamWaitEvent( la->spec_stream, la->spec_ev[( la->spec_next - 1 ) % LA_SPEC_RING], 0 ) );
        CU_CHECK( ctx, cudaMemcpyAsync( d_send, d_rec, (size_t)m * 32, cudaMemcpyDeviceToDevice, la->spec_stream ) );
        for( const Kept &t : kept )
            if( t.owner == rank )
                CU_CHECK( ctx, cudaMemcpyAsync( (uint8_t *)d_send + (size_t)most * 32 + (size_t)t.k * cost_bytes,
                                                la->slots[t.slot].dev.costs + (size_t)( t.i0 * B2w + t.i1 ) * d.mb_count, (size_t)d.mb_count * 2,
                                                cudaMemcpyDeviceToDevice, la->spec_stream ) );
        CU_CHECK( ctx, cudaEventRecord( la->spec_done_ev, la->spec_stream ) );
        CU_CHECK( ctx, cudaStreamWaitEvent( la->xch_stream, la->spec_done_ev, 0 ) );
        if( exchange( user, 3, per_rank, &d_send, &d_recv, (void *)la->xch_stream ) )
            return x264cu_fail( ctx, "finalize_batch: the exchange failed" );
        for( int r = 0; r < world; r++ )
            if( rank_count[r] )
                CU_CHECK( ctx, cudaMemcpyAsync( h_rec + (size_t)r * most * 8, (uint8_t *)d_recv + (size_t)r * per_rank, (size_t)rank_count[r] * 32,
                                                cudaMemcpyDeviceToHost, la->xch_stream ) );
        for( const Kept &t : kept )
        {
            la->slots[t.slot].spec[t.i0][t.i1].index = t.owner * most + t.k;
            if( t.owner != rank )
                CU_CHECK( ctx, cudaMemcpyAsync( la->slots[t.slot].dev.costs + (size_t)( t.i0 * B2w + t.i1 ) * d.mb_count,
                                                (uint8_t *)d_recv + (size_t)t.owner * per_rank + (size_t)most * 32 + (size_t)t.k * cost_bytes,
                                                (size_t)d.mb_count * 2, cudaMemcpyDeviceToDevice, la->xch_stream ) );
        }
        CU_CHECK( ctx, cudaEventRecord( la->spec_ev[ring], la->xch_stream ) );
    }
    la->spec_seq[ring] = la->spec_next++;
    la->spec_launched += m;
    {   // the slots' next uploads wait for this batch (it reads their planes and vectors)
        const int e = la->ev_next;
        la->ev_next = ( la->ev_next + 1 ) % la->n_ev;
        CU_CHECK( ctx, cudaEventSynchronize( la->ev[e] ) );
        CU_CHECK( ctx, cudaEventRecord( la->ev[e], sharded ? la->xch_stream : la->spec_stream ) );
        la->ev_seq[e] = la->ev_seq_next++;
        for( const Kept &t : kept ) slot_seen[t.slot] = 1;                 // imported lowres_costs are written into those
        for( size_t sl = 0; sl < la->slots.size(); sl++ )
            if( slot_seen[sl] ) { la->slots[sl].last_search_ev[3] = e; la->slots[sl].last_search_seq[3] = la->ev_seq[e]; }
    }
    return 0;
}

int x264cu_lookahead_finalize_batch( x264cu_lookahead_t *la, int n, const int *b_slot, const int *p0_slot, const int *p1_slot,
                                     const int *d0, const int *d1 )
{
    return la_finalize_batch( la, n, b_slot, p0_slot, p1_slot, d0, d1, nullptr, 0, 1, nullptr, nullptr );
}

int x264cu_lookahead_finalize_batch_sharded( x264cu_lookahead_t *la, int n, const int *b_slot, const int *p0_slot, const int *p1_slot,
                                             const int *d0, const int *d1, const int *owner, int rank, int world,
                                             x264cu_exchange_fn exchange, void *user )
{
    if( la && ( world < 1 || rank < 0 || rank >= world || ( world > 1 && ( !owner || !exchange ) ) ) )
        return x264cu_fail( la->ctx, "finalize_batch_sharded: bad rank %d of %d", rank, world );
    return la_finalize_batch( la, n, b_slot, p0_slot, p1_slot, d0, d1, owner, rank, world, exchange, user );
}

long x264cu_lookahead_speculation_stats( x264cu_lookahead_t *la, long *hits, long *misses )
{
    if( !la ) return -1;
    if( hits ) *hits = la->spec_hits;
    if( misses ) *misses = la->spec_misses;
    return la->spec_launched;
}

int x264cu_lookahead_frame_cost( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int *score )
{
    X264CU_ENTER_LA( la );
    if( !la || !frames || !score ) return -1;
    x264cu_ctx *ctx = la->ctx;
    const LaDims &d = la->d;
    if( !( p0 <= b && b <= p1 ) || b - p0 > d.B + 1 || p1 - b > d.B + 1 )
        return x264cu_fail( ctx, "frame_cost: bad frame triple (%d,%d,%d)", p0, p1, b );
    const int sb = frames[b], s0 = frames[p0], s1 = frames[p1];
    for( int s : { sb, s0, s1 } )
        if( s < 0 || s >= (int)la->slots.size() || !la->slots[s].in_use ) return x264cu_fail( ctx, "frame_cost: empty slot %d", s );
    LaSlotHost &fenc = la->slots[sb];
    const int i0 = b - p0, i1 = p1 - b;
    for( int s : { sb, s0, s1 } )
        if( la_slot_ready( la, s ) ) return -1;
    // memo check, slicetype.c:848-849
    if( fenc.cost_est[i0][i1] >= 0 && ( !la->p.vbv || fenc.row_satds_valid[i0][i1] ) )
    {
        *score = fenc.cost_est[i0][i1];
        return 0;
    }
    // do_search / sentinels, slicetype.c:855-866.  `requested` is the reference's state (has a cost request asked for this
    // pair yet), `searched` ours (has its search been launched, possibly ahead of time by x264cu_lookahead_search_batch)
    int n = 0;
    if( b != p0 && !fenc.requested[0][i0 - 1] )
    {
        fenc.requested[0][i0 - 1] = true;
        if( la->p.weighted_pred && b == p1 )
        {   /* slicetype.c:857-864: weights are analysed only when this first search of the pair is P-type */
            if( fenc.searched[0][i0 - 1] )
            {   // searched ahead of time: only done for pairs whose analysis ends at its early exit (x264cu_lookahead_weight_trivial)
                fenc.weight.enabled = 0; fenc.weight.scale = 1; fenc.weight.denom = 0; fenc.weight.offset = 0;
            }
            /* The analysis may itself issue an intra-only request, which reuses la->pack: run it before the job list is assembled */
            else if( la_weights_analyse( la, sb, s0, i0 - 1 ) ) return -1;
        }
        if( !fenc.searched[0][i0 - 1] )
        {
            fenc.searched[0][i0 - 1] = true;
            la_fill_job( la, la->pack.j[n], sb, s0, 0, i0 );
            if( la->p.weighted_pred && b == p1 && fenc.weight.enabled )
            {
                LaSearchJob &j = la->pack.j[n];
                j.ref_w = la->d_weight_plane + ( la->slots[s0].dev.planes[0] - la->slots[s0].plane_buf );
                j.w_enabled = 1; j.w_scale = fenc.weight.scale; j.w_denom = fenc.weight.denom; j.w_offset = fenc.weight.offset;
            }
            n++;
        }
    }
    if( b != p1 && !fenc.requested[1][i1 - 1] )
    {
        fenc.requested[1][i1 - 1] = true;
        if( !fenc.searched[1][i1 - 1] )
        {
            fenc.searched[1][i1 - 1] = true;
            la_fill_job( la, la->pack.j[n], sb, s1, 1, i1 );
            n++;
        }
    }
    // accumulator hand-over in the reference's order, slicetype.c:946-989; r = {cost_est, cost_est_aq, intra_mbs, intra cost_est, intra cost_est_aq}
    auto account = [&]( const int32_t *r ) {
        if( b == p1 ) fenc.intra_mbs[i0] = r[2];
        if( !fenc.b_intra_calculated ) { fenc.cost_est[0][0] = 0; fenc.cost_est_aq[0][0] = 0; }
        fenc.cost_est[i0][i1] = 0; fenc.cost_est_aq[i0][i1] = 0;
        if( !fenc.b_intra_calculated ) { fenc.cost_est[0][0] += r[3]; fenc.cost_est_aq[0][0] += r[4]; }
        fenc.cost_est[i0][i1] += r[0]; fenc.cost_est_aq[i0][i1] += r[1];
        if( la->p.vbv )
        {
            fenc.row_satds_valid[i0][i1] = true;
            if( !fenc.b_intra_calculated ) fenc.row_satds_valid[0][0] = true;
        }
        int sc = fenc.cost_est[i0][i1];
        if( b != p1 ) sc = (int)( (uint64_t)sc * 100 / ( 120 + la->p.bframe_bias ) );
        else fenc.b_intra_calculated = 1;
        fenc.cost_est[i0][i1] = sc;
        *score = sc;
    };
    {   // answered ahead of time?  (x264cu_lookahead_finalize_batch: the variant WITH the later reference's vectors)
        const LaSlotHost::Spec &sp = fenc.spec[i0][i1];
        const bool mvr_wanted = b < p1 && p1 - p0 - 1 <= d.B && la->slots[s1].requested[0][p1 - p0 - 1];
        if( sp.batch && !n && b != p0 && !la->p.vbv && fenc.intra_on_device && ( b == p1 || mvr_wanted ) &&
            la->spec_seq[sp.batch % LA_SPEC_RING] == sp.batch )
        {
            const int ring = (int)( sp.batch % LA_SPEC_RING );
            LA_TIMED( la->st.cost_sync, la->st.n_cost, CU_CHECK( ctx, cudaEventSynchronize( la->spec_ev[ring] ) ) );
            account( la->h_spec_rec + ( (size_t)ring * LA_SPEC_MAX + sp.index ) * 8 );
            la->spec_hits++;
            return 0;
        }
        if( sp.batch )
        {   // the on-demand path rewrites lowres_costs of this triple: not before the batch that also writes them is through
            const int ring = (int)( sp.batch % LA_SPEC_RING );
            if( la->spec_seq[ring] == sp.batch ) CU_CHECK( ctx, cudaStreamWaitEvent( ctx->stream, la->spec_ev[ring], 0 ) );
        }
        la->spec_misses++;
    }
    if( !fenc.intra_on_device )
    {
        intra_kernel<<<( d.mb_count + 63 ) / 64, 256, 0, ctx->stream>>>( d, fenc.dev.planes[0], fenc.dev.intra );
        CU_LAUNCH_CHECK( ctx );
        fenc.intra_on_device = true;
    }
    if( la_launch_searches( la, n, ctx->stream ) ) return -1;
    // prefetched searches this request reads (its own two lists and the temporal-direct vectors of the later reference)
    if( b != p0 && la_wait_pending( la, fenc, 0, i0 - 1 ) ) return -1;
    if( b != p1 && la_wait_pending( la, fenc, 1, i1 - 1 ) ) return -1;
    if( b < p1 && p1 - p0 - 1 <= d.B && la_wait_pending( la, la->slots[s1], 0, p1 - p0 - 1 ) ) return -1;

    int dist_scale_factor = 128;
    if( p1 != p0 ) dist_scale_factor = ( ( ( b - p0 ) << 8 ) + ( ( p1 - p0 ) >> 1 ) ) / ( p1 - p0 );
    LaFinalizeArgs A;
    memset( &A, 0, sizeof( A ) );
    const int B1 = d.B + 1, B2w = d.B + 2;
    A.fenc = fenc.dev.planes[0];
    for( int i = 0; i < 4; i++ ) { A.ref0[i] = la->slots[s0].dev.planes[i]; A.ref1[i] = la->slots[s1].dev.planes[i]; }
    A.b_inter = p0 != p1;
    A.b_bidir = b < p1;
    if( b != p0 )
    {
        A.mvs0 = fenc.dev.mvs + (size_t)( 0 * B1 + i0 - 1 ) * d.mb_count * 2;
        A.cost0 = fenc.dev.mv_costs + (size_t)( 0 * B1 + i0 - 1 ) * d.mb_count;
    }
    if( b != p1 )
    {
        A.mvs1 = fenc.dev.mvs + (size_t)( 1 * B1 + i1 - 1 ) * d.mb_count * 2;
        A.cost1 = fenc.dev.mv_costs + (size_t)( 1 * B1 + i1 - 1 ) * d.mb_count;
    }
    // temporal direct only if the reference itself would have those vectors by now (slicetype.c:629-635), however early we searched
    if( A.b_bidir && p1 - p0 - 1 <= d.B && la->slots[s1].requested[0][p1 - p0 - 1] )
        A.mvr = la->slots[s1].dev.mvs + (size_t)( 0 * B1 + p1 - p0 - 1 ) * d.mb_count * 2;
    A.intra = fenc.dev.intra;
    A.qscale = fenc.dev.qscale;
    A.costs = fenc.dev.costs + (size_t)( i0 * B2w + i1 ) * d.mb_count;
    // row sums (slicetype.c:968-977): the inter rows of this request go to row_satds[b-p0][p1-b]; the intra rows
    // go to row_satds[0][0] only while b_intra_calculated is unset.  Anything else lands in scratch rows.
    int32_t *scratch_rows = (int32_t *)x264cu_scratch( ctx, 4, (size_t)2 * d.mb_h * 4 );
    if( !scratch_rows ) return -1;
    A.row_inter = A.b_inter ? fenc.dev.row_satds + (size_t)( i0 * B2w + i1 ) * d.mb_h : scratch_rows;
    A.row_intra = !fenc.b_intra_calculated ? fenc.dev.row_satds : scratch_rows + d.mb_h;
    A.record = la->d_record;
    A.dist_scale_factor = dist_scale_factor;
    A.bipred_weight = d.bipred_weighted ? 64 - ( dist_scale_factor >> 2 ) : 32;
    CU_CHECK( ctx, cudaMemsetAsync( la->d_record, 0, 64, ctx->stream ) );
    CU_CHECK( ctx, cudaMemsetAsync( A.row_inter, 0, d.mb_h * 4, ctx->stream ) );
    CU_CHECK( ctx, cudaMemsetAsync( A.row_intra, 0, d.mb_h * 4, ctx->stream ) );
    finalize_kernel<<<( d.mb_count + LA_FIN_THREADS / 4 - 1 ) / ( LA_FIN_THREADS / 4 ), LA_FIN_THREADS, 0, ctx->stream>>>( d, A );
    CU_LAUNCH_CHECK( ctx );
    CU_CHECK( ctx, cudaMemcpyAsync( la->h_record, la->d_record, 32, cudaMemcpyDeviceToHost, ctx->stream ) );
    LA_TIMED( la->st.cost_sync, la->st.n_cost, CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) ) );

    account( la->h_record );
    return 0;
}

static int la_check_slot( x264cu_lookahead *la, int slot )
{
    if( !la ) return -1;
    for( int i = 0; i < 2; i++ ) cudaStreamSynchronize( la->search_streams[i] );          // read-backs see prefetched searches too
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use )
        return x264cu_fail( la->ctx, "lookahead: slot %d is empty / out of range", slot );
    return la_slot_ready( la, slot );
}

int x264cu_lookahead_get_mvs( x264cu_lookahead_t *la, int slot, int list, int dist_minus1, int16_t *h_mvs, int32_t *h_mv_costs )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) ) return -1;
    const LaDims &d = la->d;
    if( list < 0 || list > 1 || dist_minus1 < 0 || dist_minus1 > d.B ) return x264cu_fail( la->ctx, "get_mvs: bad index" );
    const size_t idx = (size_t)list * ( d.B + 1 ) + dist_minus1;
    if( h_mvs ) CU_CHECK( la->ctx, cudaMemcpyAsync( h_mvs, la->slots[slot].dev.mvs + idx * d.mb_count * 2, d.mb_count * 4, cudaMemcpyDeviceToHost, la->ctx->stream ) );
    if( h_mv_costs ) CU_CHECK( la->ctx, cudaMemcpyAsync( h_mv_costs, la->slots[slot].dev.mv_costs + idx * d.mb_count, d.mb_count * 4, cudaMemcpyDeviceToHost, la->ctx->stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->ctx->stream ) );
    if( h_mvs && !la->slots[slot].searched[list][dist_minus1] ) h_mvs[0] = 0x7FFF;      // the reference's sentinel (mc.c:478-480)
    return 0;
}

int x264cu_lookahead_get_weight( x264cu_lookahead_t *la, int slot, int *out4 )
{
    if( !la || !out4 ) return -1;
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use )
        return x264cu_fail( la->ctx, "lookahead: slot %d is empty / out of range", slot );
    const LaWeight &w = la->slots[slot].weight;
    out4[0] = w.enabled; out4[1] = w.scale; out4[2] = w.denom; out4[3] = w.offset;
    return 0;
}

int x264cu_lookahead_get_costs( x264cu_lookahead_t *la, int slot, int i0, int i1, uint16_t *h_out )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) ) return -1;
    const LaDims &d = la->d;
    if( i0 < 0 || i1 < 0 || i0 > d.B + 1 || i1 > d.B + 1 ) return x264cu_fail( la->ctx, "get_costs: bad index" );
    CU_CHECK( la->ctx, cudaMemcpyAsync( h_out, la->slots[slot].dev.costs + (size_t)( i0 * ( d.B + 2 ) + i1 ) * d.mb_count, d.mb_count * 2, cudaMemcpyDeviceToHost, la->ctx->stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->ctx->stream ) );
    return 0;
}

int x264cu_lookahead_get_intra( x264cu_lookahead_t *la, int slot, int32_t *h_out )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) ) return -1;
    CU_CHECK( la->ctx, cudaMemcpyAsync( h_out, la->slots[slot].dev.intra, la->d.mb_count * 4, cudaMemcpyDeviceToHost, la->ctx->stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->ctx->stream ) );
    return 0;
}

int x264cu_lookahead_get_row_satds( x264cu_lookahead_t *la, int slot, int i0, int i1, int32_t *h_rows )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) ) return -1;
    const LaDims &d = la->d;
    if( i0 < 0 || i1 < 0 || i0 > d.B + 1 || i1 > d.B + 1 ) return x264cu_fail( la->ctx, "get_row_satds: bad index" );
    CU_CHECK( la->ctx, cudaMemcpyAsync( h_rows, la->slots[slot].dev.row_satds + (size_t)( i0 * ( d.B + 2 ) + i1 ) * d.mb_h, d.mb_h * 4, cudaMemcpyDeviceToHost, la->ctx->stream ) );
    CU_CHECK( la->ctx, cudaStreamSynchronize( la->ctx->stream ) );
    return 0;
}

int x264cu_lookahead_get_cost_est( x264cu_lookahead_t *la, int slot, int i0, int i1, int *cost_est, int *cost_est_aq, int *intra_mbs )
{
    // the memo lives on the host (x264cu_lookahead_frame_cost has already synchronised on what it returns): no device wait
    if( !la ) return -1;
    if( slot < 0 || slot >= (int)la->slots.size() || !la->slots[slot].in_use )
        return x264cu_fail( la->ctx, "lookahead: slot %d is empty / out of range", slot );
    const LaDims &d = la->d;
    if( i0 < 0 || i1 < 0 || i0 > d.B + 1 || i1 > d.B + 1 ) return x264cu_fail( la->ctx, "get_cost_est: bad index" );
    LaSlotHost &s = la->slots[slot];
    if( cost_est ) *cost_est = s.cost_est[i0][i1];
    if( cost_est_aq ) *cost_est_aq = s.cost_est_aq[i0][i1];
    if( intra_mbs ) *intra_mbs = s.intra_mbs[i0];
    return 0;
}

int x264cu_lookahead_get_lowres_plane( x264cu_lookahead_t *la, int slot, int plane, uint8_t *h_out, intptr_t *stride )
{
    X264CU_ENTER_LA( la );
    if( la_check_slot( la, slot ) ) return -1;
    if( plane < 0 || plane > 3 ) return x264cu_fail( la->ctx, "get_lowres_plane: bad plane" );
    const LaDims &d = la->d;
    if( stride ) *stride = d.stride;
    if( h_out )
    {   // padded plane starting at (-PAD,-PAD), stride bytes per row, ll + 2*PAD rows
        const uint8_t *src = la->slots[slot].dev.planes[plane] - (size_t)X264CU_PAD * d.stride - X264CU_PAD;
        CU_CHECK( la->ctx, cudaMemcpyAsync( h_out, src, (size_t)d.stride * ( la->ll + 2 * X264CU_PAD ), cudaMemcpyDeviceToHost, la->ctx->stream ) );
        CU_CHECK( la->ctx, cudaStreamSynchronize( la->ctx->stream ) );
    }
    return 0;
}

} // extern "C"
