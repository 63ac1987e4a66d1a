// x264cu_me_search_batch: warp-per-search replay of x264_me_search_ref (encoder/me.c:182-992); x264cu_me_refine_bidir_batch: of
// x264_me_refine_bidir_satd (me.c:1027-1183).  Device code: me_dev.cuh.
#include "ctx.h"
#include "me_dev.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace x264cu;

struct MeJob                                    // == x264cu_me_job_t
{
    int32_t i_pixel; uint32_t fenc_off, ref_off; int16_t mvp[2]; int16_t mvc[9][2]; int32_t i_mvc;
    int16_t mv_min_spel[2], mv_max_spel[2]; int32_t halfpel_thresh;
};
struct MeResult { int16_t mv[2]; int32_t cost, cost_mv, halfpel_thresh; };
struct MeFrameJob { MeJob job; int16_t i_ref, i_lambda; };          // == x264cu_me_frame_job_t
static_assert( sizeof( MeJob ) == sizeof( x264cu_me_job_t ) && sizeof( MeResult ) == sizeof( x264cu_me_result_t ) &&
               sizeof( MeFrameJob ) == sizeof( x264cu_me_frame_job_t ), "ABI structs" );
// one reference picture of a frame-level batch, as the kernel reads it
struct MeRefDev { const uint8_t *fref[4], *fref_w, *fref_uv; LaWeight w[3]; };
struct MeFrameDev { const MeRefDev *refs; const uint16_t *tabs; int tab_stride; int job_bytes; };   // refs == NULL: plain MeJob array

template <int BW, int BH, bool EXH>
__device__ __noinline__ void run_job( const MeShared &g, const MeJob &j, int lane, MeResult &r, uint2 *tesa_list )
{
    int mvx, mvy, cost, cost_mv, thresh = j.halfpel_thresh;
    int16_t lim[4] = { j.mv_min_spel[0], j.mv_min_spel[1], j.mv_max_spel[0], j.mv_max_spel[1] };
    me_search_generic<BW, BH, EXH>( g, j.i_pixel, j.fenc_off, j.ref_off, j.mvp[0], j.mvp[1], &j.mvc[0][0], j.i_mvc, lim, thresh, lane,
                               mvx, mvy, cost, cost_mv, tesa_list );
    r.mv[0] = (int16_t)mvx; r.mv[1] = (int16_t)mvy; r.cost = cost; r.cost_mv = cost_mv; r.halfpel_thresh = thresh;
}

// ---- the order the jobs are walked in: by partition size ------------------------------------------------------------------
// The search of each partition size is its own instantiation of the whole algorithm (run_job<BW,BH>: ~12 000 SASS instructions
// each, 1.3 MB for the seven).  In the caller's order (macroblock by macroblock: 16x16, 8x8, 16x8 ... interleaved) the warps
// resident on an SM run seven different instruction streams at once and stall on instruction fetch (ncu, round 2: 46 of the 51
// cycles between two issues of a warp were "no instruction").  A stable counting sort by i_pixel in chunks of ME_ORDER_CHUNK jobs
// makes the CTAs in flight share one instantiation; spatial order survives within a size, so the reference windows still meet in L2.
constexpr int ME_ORDER_CHUNK = 1024, ME_CLASSES = 8;

__global__ void __launch_bounds__( 256 )
me_order_count_kernel( const char *__restrict__ jobs, int job_bytes, int n, int *__restrict__ counts )
{
    __shared__ int hist[ME_CLASSES];
    if( threadIdx.x < ME_CLASSES ) hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * ME_ORDER_CHUNK;
    for( int i = base + threadIdx.x; i < min( base + ME_ORDER_CHUNK, n ); i += blockDim.x )
        atomicAdd( &hist[min( (unsigned)*(const int32_t *)( jobs + (size_t)i * job_bytes ), (unsigned)ME_CLASSES - 1 )], 1 );
    __syncthreads();
    if( threadIdx.x < ME_CLASSES ) counts[blockIdx.x * ME_CLASSES + threadIdx.x] = hist[threadIdx.x];
}

// counts[chunk][class] -> where that chunk's jobs of that class start in the order (classes one after the other)
__global__ void __launch_bounds__( 32 )
me_order_scan_kernel( int *__restrict__ counts, int chunks )
{
    __shared__ int total[ME_CLASSES];
    const int c = threadIdx.x;
    if( c < ME_CLASSES )
    {
        int sum = 0;
        for( int b = 0; b < chunks; b++ ) { const int v = counts[b * ME_CLASSES + c]; counts[b * ME_CLASSES + c] = sum; sum += v; }
        total[c] = sum;
    }
    __syncthreads();
    if( c < ME_CLASSES )
    {
        int base = 0;
        for( int k = 0; k < c; k++ ) base += total[k];
        for( int b = 0; b < chunks; b++ ) counts[b * ME_CLASSES + c] += base;
    }
}

__global__ void __launch_bounds__( 256 )
me_order_scatter_kernel( const char *__restrict__ jobs, int job_bytes, int n, const int *__restrict__ starts, int *__restrict__ order )
{
    __shared__ int cursor[ME_CLASSES];
    if( threadIdx.x < ME_CLASSES ) cursor[threadIdx.x] = starts[blockIdx.x * ME_CLASSES + threadIdx.x];
    __syncthreads();
    const int base = blockIdx.x * ME_ORDER_CHUNK;
    for( int i = base + threadIdx.x; i < min( base + ME_ORDER_CHUNK, n ); i += blockDim.x )
        order[atomicAdd( &cursor[min( (unsigned)*(const int32_t *)( jobs + (size_t)i * job_bytes ), (unsigned)ME_CLASSES - 1 )], 1 )] = i;
}

// 64 registers, 8 CTAs (32 warps) per SM: measured on the 4K preset-slower stream (profiles/README.md) 36.6 M searches/s with the
// compiler's own choice (168 registers), 39.7 / 41.4 / 42.5 / 43.6 M at 4 / 5 / 6 / 8 CTAs per SM; no spills at 64
#ifndef ME_MIN_CTAS
#define ME_MIN_CTAS 8
#endif
template <bool EXH>
__global__ void __launch_bounds__( 128, ME_MIN_CTAS )
me_search_kernel( MeShared g0, MeFrameDev fd, const char *__restrict__ jobs, int n, MeResult *__restrict__ results, const int *__restrict__ order )
{
    const int lane = threadIdx.x & 31;
    const int w0 = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5, nw = ( gridDim.x * blockDim.x ) >> 5;
    uint2 *tesa_list = g0.tesa_list ? g0.tesa_list + (size_t)w0 * g0.tesa_cap : nullptr;
    for( int wi = w0; wi < n; wi += nw )         // one job per warp, except TESA whose warps (and candidate lists) are bounded
    {
        const int w = order ? order[wi] : wi;
        const char *jp = jobs + (size_t)w * fd.job_bytes;
        MeJob j = *(const MeJob *)jp;            // every lane holds the (uniform) job
        MeShared g = g0;
        if( fd.refs )
        {   // frame-level batch: the job names its reference picture and its lambda
            const int i_ref = ( (const MeFrameJob *)jp )->i_ref, i_lambda = ( (const MeFrameJob *)jp )->i_lambda;
            const MeRefDev &rf = fd.refs[i_ref];
#pragma unroll
            for( int i = 0; i < 4; i++ ) g.fref[i] = rf.fref[i];
            g.fref_w = rf.fref_w; g.fref_uv = rf.fref_uv;
            g.w = rf.w[0]; g.wc[0] = rf.w[1]; g.wc[1] = rf.w[2];
            g.cost_mv = fd.tabs + (size_t)i_lambda * fd.tab_stride;
        }
        MeResult r;
        switch( j.i_pixel )
        {
            case X264CU_PIXEL_16x16: run_job<16, 16, EXH>( g, j, lane, r, tesa_list ); break;
            case X264CU_PIXEL_16x8:  run_job<16, 8, EXH>( g, j, lane, r, tesa_list ); break;
            case X264CU_PIXEL_8x16:  run_job<8, 16, EXH>( g, j, lane, r, tesa_list ); break;
            case X264CU_PIXEL_8x8:   run_job<8, 8, EXH>( g, j, lane, r, tesa_list ); break;
            case X264CU_PIXEL_8x4:   run_job<8, 4, EXH>( g, j, lane, r, tesa_list ); break;
            case X264CU_PIXEL_4x8:   run_job<4, 8, EXH>( g, j, lane, r, tesa_list ); break;
            default:                 run_job<4, 4, EXH>( g, j, lane, r, tesa_list ); break;
        }
        if( lane == 0 ) results[w] = r;
    }
}

// cost_mv[lambda] (analyse.c:143-157, :179-188), same float expressions as the reference
static void me_fill_cost_table( uint16_t *tab, int len, int lambda )
{
    for( int i = 0; i <= len; i++ )
    {
        float l = i ? log2f( (float)( i + 1 ) ) * 2.0f + 1.718f : 0.718f;
        int c = (int)( lambda * l + .5f );
        if( c > 65535 ) c = 65535;
        tab[len + i] = tab[len - i] = (uint16_t)c;
    }
}

// kept in this context's scratch slot 5 (nobody else's) and rebuilt when (lambda, range) change.  Returns the table's centre.
static const uint16_t *me_cost_table( x264cu_ctx_t *ctx, int lambda, int mv_range )
{
    const int len = 2 * 4 * mv_range;
    const void *before = ctx->scratch[5];
    uint16_t *d_tab = (uint16_t *)x264cu_scratch( ctx, 5, ( 2 * len + 1 ) * 2 + 64 );
    if( !d_tab ) return nullptr;
    if( ctx->me_tab_lambda != lambda || ctx->me_tab_range != mv_range || before != (const void *)d_tab )
    {
        std::vector<uint16_t> tab( 2 * len + 1 );
        me_fill_cost_table( tab.data(), len, lambda );
        if( cudaMemcpyAsync( d_tab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice, ctx->stream ) != cudaSuccess ||
            cudaStreamSynchronize( ctx->stream ) != cudaSuccess )
        {
            x264cu_fail( ctx, "me cost table: %s", cudaGetErrorString( cudaGetLastError() ) );
            return nullptr;
        }
        ctx->me_tab_lambda = lambda; ctx->me_tab_range = mv_range;
    }
    return d_tab + len;
}

static int me_launch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, MeShared &g, const MeFrameDev &fd, const void *d_jobs, int n,
                      x264cu_me_result_t *d_results );

extern "C" int x264cu_me_search_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, const uint8_t *d_fenc, intptr_t fenc_stride,
                                       const uint8_t *const d_fref[4], const uint8_t *d_fref_w, intptr_t ref_stride,
                                       const x264cu_me_job_t *d_jobs, int n, x264cu_me_result_t *d_results )
{
    X264CU_ENTER( ctx );
    if( !ctx || !p ) return -1;
    if( n <= 0 ) return 0;
    if( !d_fenc || !d_fref || !d_fref[0] || !d_fref[1] || !d_fref[2] || !d_fref[3] || !d_jobs || !d_results )
        return x264cu_fail( ctx, "me_search_batch: null argument" );
    if( p->lambda < 1 || p->mv_range < 32 || p->mv_range > 8192 )
        return x264cu_fail( ctx, "me_search_batch: bad parameters" );
    const uint16_t *d_tab = me_cost_table( ctx, p->lambda, p->mv_range );
    if( !d_tab ) return -1;
    MeShared g;
    memset( &g, 0, sizeof( g ) );
    g.fenc = d_fenc; g.fenc_stride = (int)fenc_stride;
    for( int i = 0; i < 4; i++ ) g.fref[i] = d_fref[i];
    g.fref_w = d_fref_w ? d_fref_w : d_fref[0];
    g.stride = (int)ref_stride;
    g.cost_mv = d_tab;
    g.w.enabled = p->weight_enabled; g.w.scale = p->weight_scale; g.w.denom = p->weight_denom; g.w.offset = p->weight_offset;
    MeFrameDev fd = { nullptr, nullptr, 0, (int)sizeof( MeJob ) };
    return me_launch( ctx, p, g, fd, d_jobs, n, d_results );
}

// x264cu_me_search_frame: the searches of one coded picture in one launch -- every job names its reference picture (multi-ref,
// both lists, weighted duplicates) and its lambda (per-macroblock quantisers: AQ, MB-tree, VBV)
extern "C" int x264cu_me_search_frame( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, const x264cu_me_frame_t *f,
                                       const x264cu_me_frame_job_t *d_jobs, int n, x264cu_me_result_t *d_results )
{
    X264CU_ENTER( ctx );
    if( !ctx || !p || !f ) return -1;
    if( n <= 0 ) return 0;
    if( !f->d_fenc || !f->refs || f->n_refs < 1 || f->n_refs > 64 || !f->lambdas || f->n_lambdas < 1 || f->n_lambdas > 128 || !d_jobs || !d_results )
        return x264cu_fail( ctx, "me_search_frame: null / out-of-range argument" );
    if( p->mv_range < 32 || p->mv_range > 8192 )
        return x264cu_fail( ctx, "me_search_frame: bad parameters" );
    if( f->chroma_me && ( !f->d_fenc_uv || f->fenc_uv_stride <= 0 || f->ref_uv_stride <= 0 ) )
        return x264cu_fail( ctx, "me_search_frame: chroma ME without chroma planes" );
    const int len = 2 * 4 * p->mv_range, tab_stride = 2 * len + 2;
    std::vector<MeRefDev> refs( f->n_refs );
    for( int i = 0; i < f->n_refs; i++ )
    {
        const x264cu_me_ref_t &r = f->refs[i];
        for( int k = 0; k < 4; k++ )
        {
            if( !r.d_fref[k] ) return x264cu_fail( ctx, "me_search_frame: reference %d lacks plane %d", i, k );
            refs[i].fref[k] = r.d_fref[k];
        }
        refs[i].fref_w = r.d_fref_w ? r.d_fref_w : r.d_fref[0];
        refs[i].fref_uv = r.d_fref_uv;
        if( f->chroma_me && !r.d_fref_uv ) return x264cu_fail( ctx, "me_search_frame: chroma ME, reference %d has no chroma plane", i );
        for( int k = 0; k < 3; k++ )
        {
            refs[i].w[k].enabled = r.weight[k][0]; refs[i].w[k].scale = r.weight[k][1];
            refs[i].w[k].denom = r.weight[k][2]; refs[i].w[k].offset = r.weight[k][3];
        }
    }
    for( int i = 0; i < f->n_lambdas; i++ )
        if( f->lambdas[i] < 1 ) return x264cu_fail( ctx, "me_search_frame: lambda %d", f->lambdas[i] );
    // the cost tables of the lambdas in use, kept until the list changes
    const void *before = ctx->scratch[11];
    uint16_t *d_tabs = (uint16_t *)x264cu_scratch( ctx, 11, (size_t)f->n_lambdas * tab_stride * 2 + 64 );
    if( !d_tabs ) return -1;
    std::vector<int> want( f->lambdas, f->lambdas + f->n_lambdas );
    if( before != (const void *)d_tabs || ctx->me_tabs_range != p->mv_range || ctx->me_tabs_lambdas != want )
    {
        std::vector<uint16_t> tabs( (size_t)f->n_lambdas * tab_stride );
        for( int i = 0; i < f->n_lambdas; i++ )
            me_fill_cost_table( tabs.data() + (size_t)i * tab_stride, len, f->lambdas[i] );
        CU_CHECK( ctx, cudaMemcpyAsync( d_tabs, tabs.data(), tabs.size() * 2, cudaMemcpyHostToDevice, ctx->stream ) );
        CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
        ctx->me_tabs_lambdas = want; ctx->me_tabs_range = p->mv_range;
    }
    MeRefDev *d_refs = (MeRefDev *)x264cu_scratch( ctx, 12, 64 * sizeof( MeRefDev ) );
    if( !d_refs ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( d_refs, refs.data(), refs.size() * sizeof( MeRefDev ), cudaMemcpyHostToDevice, ctx->stream ) );
    MeShared g;
    memset( &g, 0, sizeof( g ) );
    g.fenc = f->d_fenc; g.fenc_stride = (int)f->fenc_stride;
    g.stride = (int)f->ref_stride;
    g.chroma = f->chroma_me != 0;
    g.fenc_uv = f->d_fenc_uv; g.fenc_uv_stride = (int)f->fenc_uv_stride; g.ref_uv_stride = (int)f->ref_uv_stride;
    MeFrameDev fd = { d_refs, d_tabs + len, tab_stride, (int)sizeof( MeFrameJob ) };
    return me_launch( ctx, p, g, fd, d_jobs, n, d_results );
}

static int me_launch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, MeShared &g, const MeFrameDev &fd, const void *d_jobs, int n,
                      x264cu_me_result_t *d_results )
{
    if( p->me_method < X264CU_ME_DIA || p->me_method > X264CU_ME_TESA )
        return x264cu_fail( ctx, "me_search_batch: unknown method %d", p->me_method );
    if( p->me_method == X264CU_ME_ESA && p->me_range > 120 )
        return x264cu_fail( ctx, "me_search_batch: esa me_range %d > 120", p->me_range );
    if( p->me_method == X264CU_ME_TESA && p->me_range > 64 )
        return x264cu_fail( ctx, "me_search_batch: tesa me_range %d > 64", p->me_range );
    if( p->subpel_refine < 0 || p->subpel_refine > 11 || p->me_range < 1 )
        return x264cu_fail( ctx, "me_search_batch: bad parameters" );
    if( p->fpel_border < 0 || p->fpel_border > 16 ) return x264cu_fail( ctx, "me_search_batch: fpel_border %d", p->fpel_border );
    g.me_method = p->me_method; g.subpel_refine = p->subpel_refine; g.me_range = p->me_range; g.satd = p->mbcmp_satd;
    g.fpel_border = p->fpel_border;
    g.fpel_satd = p->mbcmp_satd && p->me_method == X264CU_ME_TESA;                   // encoder.c:1409-1427
    g.tesa_list = nullptr; g.tesa_cap = 0;
    const int warps_per_block = 4;
    int blocks = ( n + warps_per_block - 1 ) / warps_per_block;
    if( p->me_method == X264CU_ME_TESA )
    {   // a warp's candidate list can hold every position of the window: (2r+1) rows of up to 2r+4 (me.c:627-630); the warps
        // in flight are bounded so that the lists stay within 256 MB and each warp walks several jobs
        g.tesa_cap = ( 2 * p->me_range + 1 ) * ( 2 * p->me_range + 4 );
        const size_t per_block = (size_t)warps_per_block * g.tesa_cap * sizeof(uint2);
        int max_blocks = (int)( ( (size_t)256 << 20 ) / per_block );
        max_blocks = max_blocks < 1 ? 1 : max_blocks > ctx->sm_count * 8 ? ctx->sm_count * 8 : max_blocks;
        if( blocks > max_blocks ) blocks = max_blocks;
        g.tesa_list = (uint2 *)x264cu_scratch( ctx, 9, (size_t)blocks * per_block );
        if( !g.tesa_list ) return -1;
    }
    int *d_order = nullptr;
    bool by_size = n >= 4 * ME_ORDER_CHUNK;
#ifdef X264CU_TUNING
    if( getenv( "X264CU_ME_CALLER_ORDER" ) ) by_size = false;
#endif
    if( by_size )
    {   // walk the jobs by partition size (see me_order_count_kernel)
        const int chunks = ( n + ME_ORDER_CHUNK - 1 ) / ME_ORDER_CHUNK;
        int *d_counts = (int *)x264cu_scratch( ctx, 15, ( (size_t)chunks * ME_CLASSES + n ) * sizeof(int) );
        if( !d_counts ) return -1;
        d_order = d_counts + (size_t)chunks * ME_CLASSES;
        me_order_count_kernel<<<chunks, 256, 0, ctx->stream>>>( (const char *)d_jobs, fd.job_bytes, n, d_counts );
        me_order_scan_kernel<<<1, 32, 0, ctx->stream>>>( d_counts, chunks );
        me_order_scatter_kernel<<<chunks, 256, 0, ctx->stream>>>( (const char *)d_jobs, fd.job_bytes, n, d_counts, d_order );
    }
    if( p->me_method >= X264CU_ME_ESA )
        me_search_kernel<true><<<blocks, warps_per_block * 32, 0, ctx->stream>>>( g, fd, (const char *)d_jobs, n, (MeResult *)d_results, d_order );
    else
        me_search_kernel<false><<<blocks, warps_per_block * 32, 0, ctx->stream>>>( g, fd, (const char *)d_jobs, n, (MeResult *)d_results, d_order );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

// ---- x264cu_me_refine_bidir_batch ------------------------------------------------------------------------------------------
struct BidirJob                                 // == x264cu_bidir_job_t
{
    int32_t i_pixel; uint32_t fenc_off, ref0_off, ref1_off; int16_t mv[4], mvp[4]; int16_t mv_min_spel[2], mv_max_spel[2]; int32_t i_weight;
};
struct BidirResult { int16_t mv[4]; int32_t cost; };
static_assert( sizeof( BidirJob ) == sizeof( x264cu_bidir_job_t ) && sizeof( BidirResult ) == sizeof( x264cu_bidir_result_t ), "ABI structs" );

template <int BW, int BH>
__device__ __noinline__ void run_bidir( const BidirShared &g, const BidirJob &j, int lane, uint32_t *visited, BidirResult &r )
{
    int bm[4], cost;
    int16_t lim[4] = { j.mv_min_spel[0], j.mv_min_spel[1], j.mv_max_spel[0], j.mv_max_spel[1] };
    me_refine_bidir<BW, BH>( g, j.fenc_off, j.ref0_off, j.ref1_off, j.mv, j.mvp, lim, j.i_weight, lane, visited, bm, cost );
    for( int k = 0; k < 4; k++ ) r.mv[k] = (int16_t)bm[k];
    r.cost = cost;
}

__global__ void __launch_bounds__( 128 )
me_bidir_kernel( BidirShared g, const BidirJob *__restrict__ jobs, int n, BidirResult *__restrict__ results )
{
    __shared__ uint32_t visited[4][128];
    const int lane = threadIdx.x & 31, wb = threadIdx.x >> 5;
    const int w = blockIdx.x * 4 + wb;
    if( w >= n ) return;
    BidirJob j = jobs[w];
    BidirResult r;
    switch( j.i_pixel )
    {
        case X264CU_PIXEL_16x16: run_bidir<16, 16>( g, j, lane, visited[wb], r ); break;
        case X264CU_PIXEL_16x8:  run_bidir<16, 8>( g, j, lane, visited[wb], r ); break;
        case X264CU_PIXEL_8x16:  run_bidir<8, 16>( g, j, lane, visited[wb], r ); break;
        case X264CU_PIXEL_8x8:   run_bidir<8, 8>( g, j, lane, visited[wb], r ); break;
        case X264CU_PIXEL_8x4:   run_bidir<8, 4>( g, j, lane, visited[wb], r ); break;
        case X264CU_PIXEL_4x8:   run_bidir<4, 8>( g, j, lane, visited[wb], r ); break;
        default:                 run_bidir<4, 4>( g, j, lane, visited[wb], r ); break;
    }
    if( lane == 0 ) results[w] = r;
}

extern "C" int x264cu_me_refine_bidir_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, const uint8_t *d_fenc, intptr_t fenc_stride,
                                             const uint8_t *const d_fref0[4], const uint8_t *const d_fref1[4], intptr_t ref_stride,
                                             const x264cu_bidir_job_t *d_jobs, int n, x264cu_bidir_result_t *d_results )
{
    X264CU_ENTER( ctx );
    if( !ctx || !p ) return -1;
    if( n <= 0 ) return 0;
    if( !d_fenc || !d_fref0 || !d_fref1 || !d_jobs || !d_results )
        return x264cu_fail( ctx, "me_refine_bidir_batch: null argument" );
    if( p->lambda < 1 || p->mv_range < 32 || p->mv_range > 8192 )
        return x264cu_fail( ctx, "me_refine_bidir_batch: bad parameters" );
    const uint16_t *d_tab = me_cost_table( ctx, p->lambda, p->mv_range );
    if( !d_tab ) return -1;
    BidirShared g;
    g.fenc = d_fenc; g.fenc_stride = (int)fenc_stride;
    for( int i = 0; i < 4; i++ ) { g.fref0[i] = d_fref0[i]; g.fref1[i] = d_fref1[i]; }
    g.stride = (int)ref_stride;
    g.cost_mv = d_tab;
    g.satd = p->mbcmp_satd;
    me_bidir_kernel<<<( n + 3 ) / 4, 128, 0, ctx->stream>>>( g, (const BidirJob *)d_jobs, n, (BidirResult *)d_results );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

// ---- x264cu_me_refine_qpel_batch: x264_me_refine_qpel / x264_me_refine_qpel_refdupe, me.c:800-814 ---------------------------
struct RefineJob                                // == x264cu_me_refine_job_t
{
    int32_t i_pixel; uint32_t fenc_off, ref_off; int16_t mvp[2], mv[2]; int32_t cost, i_ref_cost; int16_t mv_min_spel[2], mv_max_spel[2];
    int32_t halfpel_thresh;
};
static_assert( sizeof( RefineJob ) == sizeof( x264cu_me_refine_job_t ), "ABI structs" );

template <int BW, int BH>
__device__ __noinline__ void run_refine( const MeShared &g, const RefineJob &j, int mode, int lane, MeResult &r )
{
    MeWarp<BW, BH> m;
    int16_t lim[4] = { j.mv_min_spel[0], j.mv_min_spel[1], j.mv_max_spel[0], j.mv_max_spel[1] };
    me_warp_setup<BW, BH>( m, g, j.fenc_off, j.ref_off, j.mvp[0], j.mvp[1], lim, lane );
    m.bmx = m.bmy = 0; m.bcost = LA_COST_MAX;
    const int subpel = g.subpel_refine;
    int qx = j.mv[0], qy = j.mv[1], qcost = j.cost, thresh = -1;
    if( mode == 0 )
    {   // subpel_iterations[subme][0..1]: refine_hpel, refine_qpel (me.c:38-50)
        const int hpel = subpel == 1 ? 1 : 0, qpel = subpel == 0 ? 0 : subpel <= 2 ? 1 : subpel <= 5 ? 2 : 0;
        if( j.i_pixel <= X264CU_PIXEL_8x8 ) qcost -= j.i_ref_cost;
        me_refine_subpel<BW, BH>( m, subpel, hpel, qpel, true, thresh, qx, qy, qcost );
    }
    else
    {   // at most two quarter-pel rounds of me_qpel = subpel_iterations[subme][3]
        const int me_qpel = subpel < 4 ? 0 : subpel == 4 ? 1 : subpel < 8 ? 2 : 10;
        thresh = j.halfpel_thresh;
        me_refine_subpel<BW, BH>( m, subpel, 0, min( 2, me_qpel ), false, thresh, qx, qy, qcost );
    }
    r.mv[0] = (int16_t)qx; r.mv[1] = (int16_t)qy; r.cost = qcost; r.halfpel_thresh = thresh;
    r.cost_mv = __ldg( g.cost_mv + ( qx - j.mvp[0] ) ) + __ldg( g.cost_mv + ( qy - j.mvp[1] ) );
}

__global__ void __launch_bounds__( 128 )
me_refine_kernel( MeShared g, int mode, const RefineJob *__restrict__ jobs, int n, MeResult *__restrict__ results )
{
    const int lane = threadIdx.x & 31;
    const int w = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    if( w >= n ) return;
    RefineJob j = jobs[w];
    MeResult r;
    switch( j.i_pixel )
    {
        case X264CU_PIXEL_16x16: run_refine<16, 16>( g, j, mode, lane, r ); break;
        case X264CU_PIXEL_16x8:  run_refine<16, 8>( g, j, mode, lane, r ); break;
        case X264CU_PIXEL_8x16:  run_refine<8, 16>( g, j, mode, lane, r ); break;
        case X264CU_PIXEL_8x8:   run_refine<8, 8>( g, j, mode, lane, r ); break;
        case X264CU_PIXEL_8x4:   run_refine<8, 4>( g, j, mode, lane, r ); break;
        case X264CU_PIXEL_4x8:   run_refine<4, 8>( g, j, mode, lane, r ); break;
        default:                 run_refine<4, 4>( g, j, mode, lane, r ); break;
    }
    if( lane == 0 ) results[w] = r;
}

extern "C" int x264cu_me_refine_qpel_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, int refdupe, const uint8_t *d_fenc, intptr_t fenc_stride,
                                            const uint8_t *const d_fref[4], intptr_t ref_stride,
                                            const x264cu_me_refine_job_t *d_jobs, int n, x264cu_me_result_t *d_results )
{
    X264CU_ENTER( ctx );
    if( !ctx || !p ) return -1;
    if( n <= 0 ) return 0;
    if( !d_fenc || !d_fref || !d_jobs || !d_results ) return x264cu_fail( ctx, "me_refine_qpel_batch: null argument" );
    if( p->subpel_refine < 0 || p->subpel_refine > 11 || p->lambda < 1 || p->mv_range < 32 || p->mv_range > 8192 )
        return x264cu_fail( ctx, "me_refine_qpel_batch: bad parameters" );
    const uint16_t *d_tab = me_cost_table( ctx, p->lambda, p->mv_range );
    if( !d_tab ) return -1;
    MeShared g;
    memset( &g, 0, sizeof( g ) );
    g.fenc = d_fenc; g.fenc_stride = (int)fenc_stride;
    for( int i = 0; i < 4; i++ ) g.fref[i] = d_fref[i];
    g.fref_w = d_fref[0];
    g.stride = (int)ref_stride;
    g.cost_mv = d_tab;
    g.me_method = p->me_method; g.subpel_refine = p->subpel_refine; g.me_range = p->me_range; g.satd = p->mbcmp_satd;
    g.fpel_satd = p->mbcmp_satd && p->me_method == X264CU_ME_TESA;
    g.tesa_list = nullptr; g.tesa_cap = 0;
    g.w.enabled = p->weight_enabled; g.w.scale = p->weight_scale; g.w.denom = p->weight_denom; g.w.offset = p->weight_offset;
    me_refine_kernel<<<( n + 3 ) / 4, 128, 0, ctx->stream>>>( g, !!refdupe, (const RefineJob *)d_jobs, n, (MeResult *)d_results );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}
