// x264cu_me_search_batch: warp-per-search replay of x264_me_search_ref (encoder/me.c:182-992).  Device code: me_dev.cuh.
#include "ctx.h"
#include "me_dev.cuh"
#include <math.h>
#include <vector>

using namespace x264cu;

struct MeJob                                    // == x264cu_me_job_t
{
    int32_t i_pixel; uint32_t fenc_off, ref_off; int16_t mvp[2]; int16_t mvc[8][2]; int32_t i_mvc;
    int16_t mv_min_spel[2], mv_max_spel[2]; int32_t halfpel_thresh;
};
struct MeResult { int16_t mv[2]; int32_t cost, cost_mv, halfpel_thresh; };
static_assert( sizeof( MeJob ) == sizeof( x264cu_me_job_t ) && sizeof( MeResult ) == sizeof( x264cu_me_result_t ), "ABI structs" );

template <int BW, int BH>
__device__ __noinline__ void run_job( const MeShared &g, const MeJob &j, int lane, MeResult &r )
{
    int mvx, mvy, cost, cost_mv, thresh = j.halfpel_thresh;
    int16_t lim[4] = { j.mv_min_spel[0], j.mv_min_spel[1], j.mv_max_spel[0], j.mv_max_spel[1] };
    me_search_generic<BW, BH>( g, j.i_pixel, j.fenc_off, j.ref_off, j.mvp[0], j.mvp[1], &j.mvc[0][0], j.i_mvc, lim, thresh, lane,
                               mvx, mvy, cost, cost_mv );
    r.mv[0] = (int16_t)mvx; r.mv[1] = (int16_t)mvy; r.cost = cost; r.cost_mv = cost_mv; r.halfpel_thresh = thresh;
}

__global__ void __launch_bounds__( 128 )
me_search_kernel( MeShared g, const MeJob *__restrict__ jobs, int n, MeResult *__restrict__ results )
{
    const int lane = threadIdx.x & 31;
    const int w = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    if( w >= n ) return;
    MeJob j = jobs[w];                           // every lane holds the (uniform) job
    MeResult r;
    switch( j.i_pixel )
    {
        case X264CU_PIXEL_16x16: run_job<16, 16>( g, j, lane, r ); break;
        case X264CU_PIXEL_16x8:  run_job<16, 8>( g, j, lane, r ); break;
        case X264CU_PIXEL_8x16:  run_job<8, 16>( g, j, lane, r ); break;
        case X264CU_PIXEL_8x8:   run_job<8, 8>( g, j, lane, r ); break;
        case X264CU_PIXEL_8x4:   run_job<8, 4>( g, j, lane, r ); break;
        case X264CU_PIXEL_4x8:   run_job<4, 8>( g, j, lane, r ); break;
        default:                 run_job<4, 4>( g, j, lane, r ); break;
    }
    if( lane == 0 ) results[w] = r;
}

extern "C" int x264cu_me_search_batch( x264cu_ctx_t *ctx, const x264cu_me_params_t *p, const uint8_t *d_fenc, intptr_t fenc_stride,
                                       const uint8_t *const d_fref[4], const uint8_t *d_fref_w, intptr_t ref_stride,
                                       const x264cu_me_job_t *d_jobs, int n, x264cu_me_result_t *d_results )
{
    if( !ctx || !p ) return -1;
    if( n <= 0 ) return 0;
    if( p->me_method < X264CU_ME_DIA || p->me_method > X264CU_ME_ESA )
        return x264cu_fail( ctx, "me_search_batch: method %d not supported (tesa is outside this backend)", p->me_method );
    if( p->me_method == X264CU_ME_ESA && p->me_range > 120 )
        return x264cu_fail( ctx, "me_search_batch: esa me_range %d > 120", p->me_range );
    if( p->subpel_refine < 0 || p->subpel_refine > 11 || p->lambda < 1 || p->mv_range < 32 || p->mv_range > 4096 )
        return x264cu_fail( ctx, "me_search_batch: bad parameters" );
    // cost_mv[lambda] (analyse.c:143-157, :179-188), same float expressions as the reference; kept in this context's scratch
    // slot 5 (nobody else's) and rebuilt when (lambda, range) change
    const int len = 2 * 4 * p->mv_range;
    const void *before = ctx->scratch[5];
    uint16_t *d_tab = (uint16_t *)x264cu_scratch( ctx, 5, ( 2 * len + 1 ) * 2 + 64 );
    if( !d_tab ) return -1;
    if( ctx->me_tab_lambda != p->lambda || ctx->me_tab_range != p->mv_range || before != (const void *)d_tab )
    {
        std::vector<uint16_t> tab( 2 * len + 1 );
        for( int i = 0; i <= len; i++ )
        {
            float l = i ? log2f( (float)( i + 1 ) ) * 2.0f + 1.718f : 0.718f;
            int c = (int)( p->lambda * l + .5f );
            if( c > 65535 ) c = 65535;
            tab[len + i] = tab[len - i] = (uint16_t)c;
        }
        CU_CHECK( ctx, cudaMemcpyAsync( d_tab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice, ctx->stream ) );
        CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
        ctx->me_tab_lambda = p->lambda; ctx->me_tab_range = p->mv_range;
    }
    MeShared g;
    g.fenc = d_fenc; g.fenc_stride = (int)fenc_stride;
    for( int i = 0; i < 4; i++ ) g.fref[i] = d_fref[i];
    g.fref_w = d_fref_w ? d_fref_w : d_fref[0];
    g.stride = (int)ref_stride;
    g.cost_mv = d_tab + len;
    g.me_method = p->me_method; g.subpel_refine = p->subpel_refine; g.me_range = p->me_range; g.satd = p->mbcmp_satd;
    g.w.enabled = p->weight_enabled; g.w.scale = p->weight_scale; g.w.denom = p->weight_denom; g.w.offset = p->weight_offset;
    const int warps_per_block = 4;
    me_search_kernel<<<( n + warps_per_block - 1 ) / warps_per_block, warps_per_block * 32, 0, ctx->stream>>>(
        g, (const MeJob *)d_jobs, n, (MeResult *)d_results );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}
