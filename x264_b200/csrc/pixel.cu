// Batched twins of the x264_pixel_function_t table (common/pixel.h:78-144) for sm_100a.
//
//  * cmp_batch_kernel  -- arbitrary (fenc_off, ref_off) candidate lists, the literal batched form of
//                         int cmp(pixel*,intptr_t,pixel*,intptr_t) (common/pixel.h:33) and of the x3/x4
//                         entries (pixel.h:34-35).  Reads straight from HBM/L2.
//  * mvfield_kernel    -- "every block of the frame against its displaced reference block", the shape in
//                         which me.c evaluates candidates.  Persistent CTAs; fenc tile and reference tile
//                         (+/-R halo) are staged into shared memory by TMA (cp.async.bulk.tensor.3d) through a
//                         2-stage mbarrier pipeline, so both planes cross HBM once in full 128-byte lines and
//                         the unaligned 4-byte gathers happen in shared memory.
//
// Lane mapping and metric arithmetic: pixel_dev.cuh.  No tensor cores (there is no contraction here); the
// bound is HBM bandwidth: 2*W*H + 4 algorithmic bytes per candidate.
#include "ctx.h"
#include "pixel_dev.cuh"
#include <stdlib.h>

using namespace x264cu;

static const int k_pixel_w[X264CU_PIXEL_NB] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static const int k_pixel_h[X264CU_PIXEL_NB] = { 16, 8, 16, 8, 4, 8, 4, 16 };

// ---------------------------------------------------------------------------------------------------
// generic candidate-list kernel
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_unaligned4( const uint8_t *p )
{
    uintptr_t u = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)( u & ~(uintptr_t)3 );
    uint32_t sh = ( (uint32_t)u & 3u ) * 8u;
    uint32_t lo = __ldg( q );
    uint32_t hi = sh ? __ldg( q + 1 ) : 0u;       // never touch the next word when aligned (may be past the plane)
    return funnel( lo, hi, sh );
}

template <int METRIC, int BW, int BH, int NREFS>
__global__ void __launch_bounds__( 256 )
cmp_batch_kernel( const uint8_t *__restrict__ fenc, intptr_t fenc_stride,
                  const uint8_t *__restrict__ ref, intptr_t ref_stride,
                  const uint32_t *__restrict__ cand, int n_out, int32_t *__restrict__ out )
{
    using G = BlockGeom<BW, BH>;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = ( gridDim.x * blockDim.x ) >> 5;
    const int n_tasks = ( n_out + G::CPT - 1 ) / G::CPT;
    const int sx = G::sub_x( lane ), sy = G::sub_y( lane );
    for( int task = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5; task < n_tasks; task += warps_per_grid )
    {
        int idx = task * G::CPT + G::cand_in_task( lane );
        bool valid = idx < n_out;
        int cidx = valid ? idx : n_out - 1;
        uint32_t fo, ro;
        if( NREFS == 1 )
        {
            uint2 c = __ldg( (const uint2 *)cand + cidx );
            fo = c.x; ro = c.y;
        }
        else
        {   // x264cu_cand_x4_t: { fenc_off, ref_off[4] }, output index = entry*4 + j
            int e = cidx >> 2, j = cidx & 3;
            fo = __ldg( cand + e * 5 );
            ro = __ldg( cand + e * 5 + 1 + ( j < NREFS ? j : 0 ) );
        }
        const uint8_t *pa = fenc + fo + (intptr_t)sy * fenc_stride + sx;
        const uint8_t *pb = ref + ro + (intptr_t)sy * ref_stride + sx;
        uint32_t a[4], b[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            a[r] = ldg_unaligned4( pa + r * fenc_stride );
            b[r] = ldg_unaligned4( pb + r * ref_stride );
        }
        int v = G::reduce( metric4x4<METRIC>( a, b, lane ) );
        if( valid && G::leader( lane ) )
            out[idx] = metric_finish<METRIC>( v );
    }
}

// ---------------------------------------------------------------------------------------------------
// TMA / mbarrier helpers (PTX; SASS: UTMALDG + SYNCS)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32( const void *p ) { return (uint32_t)__cvta_generic_to_shared( p ); }

__device__ __forceinline__ void mbar_init( uint64_t *bar, uint32_t count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( count ) );
}
__device__ __forceinline__ void mbar_expect_tx( uint64_t *bar, uint32_t bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_wait( uint64_t *bar, uint32_t parity )
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"( smem_u32( bar ) ), "r"( parity ) : "memory" );
}
__device__ __forceinline__ void tma_load_3d( void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2 )
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"( smem_u32( dst ) ), "l"( tm ), "r"( smem_u32( bar ) ), "r"( c0 ), "r"( c1 ), "r"( c2 ) : "memory" );
}

// ---------------------------------------------------------------------------------------------------
// MV-field kernel
// ---------------------------------------------------------------------------------------------------
struct MvFieldParams
{
    int width, height, n_planes;
    int blocks_x, blocks_y;
    int tiles_x, tiles_y, n_tiles;
    int k_cands;
    const int16_t *mv;
    int32_t *out;
    // fall-back path for vectors outside the staged halo
    const uint8_t *ref_origin;
    intptr_t ref_stride, ref_plane_pitch;
};

template <int TW, int TH, int R> struct TileCfg
{
    static constexpr int FP = TW + 32;                 // fenc smem pitch (TMA box width; 16-byte multiple)
    static constexpr int RP = TW + 2 * R;              // ref smem pitch
    static constexpr int RH = TH + 2 * R;
    static constexpr int FENC_BYTES = FP * TH;
    static constexpr int REF_BYTES = RP * RH;
    static constexpr int STAGE_BYTES = FENC_BYTES + REF_BYTES;
    static_assert( FP == RP, "fenc and ref tiles share one row pitch" );
    static_assert( FP <= 256 && RP <= 256 && RH <= 256, "TMA box dimension limit" );
    static_assert( STAGE_BYTES % 128 == 0 && FENC_BYTES % 128 == 0, "TMA smem alignment" );
};

// shared-memory loads with compile-time immediates (one LDS each, no address arithmetic in the hot loop)
template <int OFF> __device__ __forceinline__ uint32_t lds_off( uint32_t addr )
{
    uint32_t v;
    asm volatile( "ld.shared.b32 %0, [%1+%2];" : "=r"( v ) : "r"( addr ), "n"( OFF ) );
    return v;
}

// One 4-row strip of a task: the lane owns the 4x4 block at tile column 4*lane, rows S*4 .. S*4+3 of the task.
template <int METRIC, int PITCH, int S>
__device__ __forceinline__ void strip_load( uint32_t fenc_addr, uint32_t ref_addr, uint32_t sh, uint32_t a[4], uint32_t b[4] )
{
    a[0] = lds_off<( S * 4 + 0 ) * PITCH>( fenc_addr );
    a[1] = lds_off<( S * 4 + 1 ) * PITCH>( fenc_addr );
    a[2] = lds_off<( S * 4 + 2 ) * PITCH>( fenc_addr );
    a[3] = lds_off<( S * 4 + 3 ) * PITCH>( fenc_addr );
    b[0] = funnel( lds_off<( S * 4 + 0 ) * PITCH>( ref_addr ), lds_off<( S * 4 + 0 ) * PITCH + 4>( ref_addr ), sh );
    b[1] = funnel( lds_off<( S * 4 + 1 ) * PITCH>( ref_addr ), lds_off<( S * 4 + 1 ) * PITCH + 4>( ref_addr ), sh );
    b[2] = funnel( lds_off<( S * 4 + 2 ) * PITCH>( ref_addr ), lds_off<( S * 4 + 2 ) * PITCH + 4>( ref_addr ), sh );
    b[3] = funnel( lds_off<( S * 4 + 3 ) * PITCH>( ref_addr ), lds_off<( S * 4 + 3 ) * PITCH + 4>( ref_addr ), sh );
}

template <int METRIC, int PITCH, int NY, int S>
__device__ __forceinline__ int strips_metric( uint32_t fenc_addr, uint32_t ref_addr, uint32_t sh, int lane )
{
    if constexpr( S >= NY )
        return 0;
    else if constexpr( METRIC == M_SA8D )
    {   // 8x8 Hadamard = two vertically adjacent 4x4 coefficient sets of this lane + the horizontal neighbour lane
        uint32_t a[4], b[4];
        int c0[16], c1[16];
        strip_load<METRIC, PITCH, S>( fenc_addr, ref_addr, sh, a, b );
        had4x4( a, b, c0 );
        strip_load<METRIC, PITCH, S + 1>( fenc_addr, ref_addr, sh, a, b );
        had4x4( a, b, c1 );
        int acc = 0;
#pragma unroll
        for( int i = 0; i < 16; i++ )
        {
            int u = c0[i] + c1[i], d = c0[i] - c1[i];
            int tu = __shfl_xor_sync( 0xffffffffu, u, 1 ), td = __shfl_xor_sync( 0xffffffffu, d, 1 );
            acc += abs( ( lane & 1 ) ? tu - u : u + tu ) + abs( ( lane & 1 ) ? td - d : d + td );
        }
        return acc + strips_metric<METRIC, PITCH, NY, S + 2>( fenc_addr, ref_addr, sh, lane );
    }
    else
    {
        uint32_t a[4], b[4];
        strip_load<METRIC, PITCH, S>( fenc_addr, ref_addr, sh, a, b );
        int v = METRIC == M_SAD ? sad4x4( a, b ) : METRIC == M_SSD ? ssd4x4( a, b ) : satd4x4( a, b );
        return v + strips_metric<METRIC, PITCH, NY, S + 1>( fenc_addr, ref_addr, sh, lane );
    }
}

// Same strip walk with the reference rows fetched from global memory (vectors outside the staged halo).
template <int METRIC, int PITCH, int NY>
__device__ __noinline__ int strips_metric_global( uint32_t fenc_addr, const uint8_t *g, intptr_t ref_stride, int lane )
{
    int acc = 0;
    int cprev[16];
    for( int s = 0; s < NY; s++ )
    {
        uint32_t a[4], b[4];
#pragma unroll
        for( int j = 0; j < 4; j++ )
        {
            asm volatile( "ld.shared.b32 %0, [%1];" : "=r"( a[j] ) : "r"( fenc_addr + ( s * 4 + j ) * PITCH ) );
            b[j] = ldg_unaligned4( g + (intptr_t)( s * 4 + j ) * ref_stride );
        }
        if( METRIC == M_SA8D )
        {
            int c[16];
            had4x4( a, b, c );
            if( s & 1 )
            {
#pragma unroll
                for( int i = 0; i < 16; i++ )
                {
                    int u = cprev[i] + c[i], d = cprev[i] - c[i];
                    int tu = __shfl_xor_sync( 0xffffffffu, u, 1 ), td = __shfl_xor_sync( 0xffffffffu, d, 1 );
                    acc += abs( ( lane & 1 ) ? tu - u : u + tu ) + abs( ( lane & 1 ) ? td - d : d + td );
                }
            }
            else
            {
#pragma unroll
                for( int i = 0; i < 16; i++ ) cprev[i] = c[i];
            }
        }
        else
            acc += METRIC == M_SAD ? sad4x4( a, b ) : METRIC == M_SSD ? ssd4x4( a, b ) : satd4x4( a, b );
    }
    return acc;
}

// Work decomposition: a warp task is one row of blocks across the 128-pixel tile: lane l owns the 4-pixel
// column 4*l and walks the block's BH/4 four-row strips in registers; the BW/4 lanes of a block are summed with
// xor-shuffles.  Every shared-memory load of the fenc tile reads 128 contiguous bytes of one row.
template <int METRIC, int BW, int BH, int TW, int TH, int R, int NSTAGE, int NWARPS>
__global__ void __launch_bounds__( NWARPS * 32 )
mvfield_kernel( const __grid_constant__ CUtensorMap tm_fenc, const __grid_constant__ CUtensorMap tm_ref,
                const MvFieldParams p )
{
    using T = TileCfg<TW, TH, R>;
    static_assert( TW == 128, "one lane per 4-pixel column of the tile" );
    constexpr int PITCH = T::FP;
    constexpr int LX = BW / 4, NY = BH / 4;
    constexpr int N_TASKS = TH / BH;
    static_assert( N_TASKS % NWARPS == 0, "block rows must divide evenly over the warps" );
    constexpr int TASKS_PER_WARP = N_TASKS / NWARPS;

    extern __shared__ __align__( 1024 ) uint8_t smem[];
    uint64_t *full = (uint64_t *)( smem + NSTAGE * T::STAGE_BYTES + 64 );
    const uint32_t smem_base = smem_u32( smem );

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if( threadIdx.x == 0 )
    {
        for( int s = 0; s < NSTAGE; s++ ) mbar_init( &full[s], 1 );
        asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    __syncthreads();

    // tile cursor (tx,ty,f) advanced incrementally by gridDim.x tiles: no divisions in the loop
    const int tiles_x = p.tiles_x, tiles_y = p.tiles_y;
    const int tiles_per_plane = tiles_x * tiles_y;
    int f = blockIdx.x / tiles_per_plane;
    int t2 = blockIdx.x - f * tiles_per_plane;
    int ty = t2 / tiles_x, tx = t2 - ty * tiles_x;
    const int step_f = gridDim.x / tiles_per_plane;
    const int step_r = gridDim.x - step_f * tiles_per_plane;
    const int step_y = step_r / tiles_x, step_x = step_r - step_y * tiles_x;
    auto advance = [&]( int &ax, int &ay, int &af ) {
        ax += step_x;
        if( ax >= tiles_x ) { ax -= tiles_x; ay++; }
        ay += step_y;
        if( ay >= tiles_y ) { ay -= tiles_y; af++; }
        af += step_f;
    };
    auto issue = [&]( int ax, int ay, int af, int stage ) {
        uint8_t *dst = smem + stage * T::STAGE_BYTES;
        mbar_expect_tx( &full[stage], T::STAGE_BYTES );
        // tensor coordinates are relative to (-PAD,-PAD) of the padded plane
        tma_load_3d( dst, &tm_fenc, &full[stage], X264CU_PAD + ax * TW, X264CU_PAD + ay * TH, af );
        tma_load_3d( dst + T::FENC_BYTES, &tm_ref, &full[stage], X264CU_PAD + ax * TW - R, X264CU_PAD + ay * TH - R, af );
    };

    // producer cursor runs NSTAGE tiles ahead of the consumer cursor (thread 0 only)
    int ptx = tx, pty = ty, pf = f, ptile = blockIdx.x;
    if( threadIdx.x == 0 )
        for( int s = 0; s < NSTAGE; s++ )
        {
            if( ptile < p.n_tiles ) issue( ptx, pty, pf, s );
            advance( ptx, pty, pf );
            ptile += gridDim.x;
        }

    const int width = p.width, height = p.height, blocks_x = p.blocks_x, blocks_y = p.blocks_y;
    const int n_tiles = p.n_tiles, k_cands = p.k_cands;
    const uint32_t *__restrict__ mvp = (const uint32_t *)p.mv;
    int32_t *__restrict__ outp = p.out;
    const int k_stride = p.n_planes * blocks_x * blocks_y;
    const int lane_blk = lane / LX;                      // block column of this lane inside the tile
    const bool lead = lane % LX == 0;
    const int lane_x = lane * 4;

    int it = 0;
    for( int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++ )
    {
        const int stage = it % NSTAGE;
        const uint32_t parity = ( it / NSTAGE ) & 1;
        const int x0 = tx * TW, y0 = ty * TH;
        const int tile_blk = ( f * blocks_y + y0 / BH ) * blocks_x + x0 / BW;
        const bool col_ok = x0 + lane_x < width;

        int oidx[TASKS_PER_WARP];
        uint32_t mvw[TASKS_PER_WARP];
#pragma unroll
        for( int t = 0; t < TASKS_PER_WARP; t++ )
        {
            const int row = warp + t * NWARPS;              // block row inside the tile
            const bool ok = col_ok && y0 + row * BH < height;
            oidx[t] = ok ? tile_blk + row * blocks_x + lane_blk : -1;
            mvw[t] = ok ? __ldg( mvp + oidx[t] ) : 0u;
        }

        mbar_wait( &full[stage], parity );
        const uint32_t sf = smem_base + stage * T::STAGE_BYTES + lane_x;           // fenc tile, this lane's column
        const uint32_t sr = sf + T::FENC_BYTES + R * PITCH + R;                    // ref tile at zero displacement

        for( int k = 0; k < k_cands; k++ )
        {
#pragma unroll
            for( int t = 0; t < TASKS_PER_WARP; t++ )
            {
                const int row = warp + t * NWARPS;
                const uint32_t mvcur = mvw[t];
                if( k + 1 < k_cands )
                    mvw[t] = oidx[t] >= 0 ? __ldg( mvp + ( k + 1 ) * k_stride + oidx[t] ) : 0u;
                const int mx = (int16_t)( mvcur & 0xffff ), my = (int16_t)( mvcur >> 16 );
                const uint32_t fa = sf + row * ( BH * PITCH );
                const bool inside = (unsigned)( mx + R ) <= 2u * R && (unsigned)( my + R ) <= 2u * R;
                int v;
                if( __all_sync( 0xffffffffu, inside ) )
                {
                    const int o = row * ( BH * PITCH ) + my * PITCH + mx;
                    const uint32_t ra = ( sr + o ) & ~3u;
                    const uint32_t sh = ( ( sr + o ) & 3u ) * 8u;
                    v = strips_metric<METRIC, PITCH, NY, 0>( fa, ra, sh, lane );
                }
                else
                {
                    const uint8_t *g = p.ref_origin + (intptr_t)f * p.ref_plane_pitch
                                     + (intptr_t)( y0 + row * BH + my ) * p.ref_stride + ( x0 + lane_x + mx );
                    v = oidx[t] >= 0 ? 1 : 0;
                    v = strips_metric_global<METRIC, PITCH, NY>( fa, oidx[t] >= 0 ? g : p.ref_origin, p.ref_stride, lane );
                }
                if( LX >= 2 ) v += __shfl_xor_sync( 0xffffffffu, v, 1 );
                if( LX >= 4 ) v += __shfl_xor_sync( 0xffffffffu, v, 2 );
                if( lead && oidx[t] >= 0 )
                    outp[k * k_stride + oidx[t]] = metric_finish<METRIC>( v );
            }
        }
        __syncthreads();                                        // everyone is done reading this stage
        if( threadIdx.x == 0 )
        {
            if( ptile < n_tiles ) issue( ptx, pty, pf, stage );
            advance( ptx, pty, pf );
            ptile += gridDim.x;
        }
        advance( tx, ty, f );
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int METRIC, int BW, int BH, int NREFS>
static int launch_batch( x264cu_ctx *ctx, const uint8_t *fenc, intptr_t fs, const uint8_t *ref, intptr_t rs,
                         const uint32_t *cand, int n_out, int32_t *out )
{
    using G = BlockGeom<BW, BH>;
    int n_tasks = ( n_out + G::CPT - 1 ) / G::CPT;
    int blocks = ( n_tasks + 7 ) / 8;
    int max_blocks = ctx->sm_count * 8;
    if( blocks > max_blocks ) blocks = max_blocks;
    if( blocks < 1 ) blocks = 1;
    cmp_batch_kernel<METRIC, BW, BH, NREFS><<<blocks, 256, 0, ctx->stream>>>( fenc, fs, ref, rs, cand, n_out, out );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

#define DISPATCH_SIZE( M, NREFS, ... )                                                              \
    switch( i_pixel )                                                                               \
    {                                                                                               \
        case X264CU_PIXEL_16x16: return launch_batch<M, 16, 16, NREFS>( __VA_ARGS__ );              \
        case X264CU_PIXEL_16x8:  return launch_batch<M, 16, 8, NREFS>( __VA_ARGS__ );               \
        case X264CU_PIXEL_8x16:  return launch_batch<M, 8, 16, NREFS>( __VA_ARGS__ );               \
        case X264CU_PIXEL_8x8:   return launch_batch<M, 8, 8, NREFS>( __VA_ARGS__ );                \
        case X264CU_PIXEL_8x4:   return launch_batch<M, 8, 4, NREFS>( __VA_ARGS__ );                \
        case X264CU_PIXEL_4x8:   return launch_batch<M, 4, 8, NREFS>( __VA_ARGS__ );                \
        case X264CU_PIXEL_4x4:   return launch_batch<M, 4, 4, NREFS>( __VA_ARGS__ );                \
        case X264CU_PIXEL_4x16:  return launch_batch<M, 4, 16, NREFS>( __VA_ARGS__ );               \
    }

template <int NREFS>
static int dispatch_batch( x264cu_ctx *ctx, int metric, int i_pixel, const uint8_t *fenc, intptr_t fs,
                           const uint8_t *ref, intptr_t rs, const uint32_t *cand, int n_out, int32_t *out )
{
    switch( metric )
    {
        case X264CU_SAD:  DISPATCH_SIZE( M_SAD, NREFS, ctx, fenc, fs, ref, rs, cand, n_out, out ) break;
        case X264CU_SSD:  DISPATCH_SIZE( M_SSD, NREFS, ctx, fenc, fs, ref, rs, cand, n_out, out ) break;
        case X264CU_SATD: DISPATCH_SIZE( M_SATD, NREFS, ctx, fenc, fs, ref, rs, cand, n_out, out ) break;
        case X264CU_SA8D:
            if( i_pixel == X264CU_PIXEL_16x16 ) return launch_batch<M_SA8D, 16, 16, NREFS>( ctx, fenc, fs, ref, rs, cand, n_out, out );
            if( i_pixel == X264CU_PIXEL_8x8 ) return launch_batch<M_SA8D, 8, 8, NREFS>( ctx, fenc, fs, ref, rs, cand, n_out, out );
            return x264cu_fail( ctx, "sa8d exists for 16x16 and 8x8 only (common/pixel.c:369-381)" );
    }
    return x264cu_fail( ctx, "bad metric %d / block size %d", metric, i_pixel );
}

static int make_tensor_map( x264cu_ctx *ctx, CUtensorMap *tm, const x264cu_planes_t *pl, int box_w, int box_h )
{
    const uint8_t *base = pl->d_origin - (intptr_t)X264CU_PAD * pl->stride - X264CU_PAD;
    if( ( (uintptr_t)base & 15 ) || ( pl->stride & 15 ) || ( pl->plane_pitch & 15 ) )
        return x264cu_fail( ctx, "planes must be 16-byte aligned (origin-pad %p, stride %ld, pitch %ld)", base,
                            (long)pl->stride, (long)pl->plane_pitch );
    cuuint64_t dims[3] = { (cuuint64_t)pl->stride, (cuuint64_t)( pl->height + 2 * X264CU_PAD ), (cuuint64_t)pl->n_planes };
    cuuint64_t strides[2] = { (cuuint64_t)pl->stride, (cuuint64_t)pl->plane_pitch };
    cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = ctx->encode_tiled( tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
    if( r != CUDA_SUCCESS )
        return x264cu_fail( ctx, "cuTensorMapEncodeTiled failed (%d)", (int)r );
    return 0;
}

template <int METRIC, int BW, int BH, int TW = 128, int TH = 64, int R = 16, int NSTAGE = 2, int NWARPS = 4>
static int launch_mvfield( x264cu_ctx *ctx, const x264cu_planes_t *fenc, const x264cu_planes_t *ref, int k_cands,
                           const int16_t *d_mv, int32_t *d_out )
{
    using T = TileCfg<TW, TH, R>;
    CUtensorMap tm_fenc, tm_ref;
    if( make_tensor_map( ctx, &tm_fenc, fenc, T::FP, TH ) || make_tensor_map( ctx, &tm_ref, ref, T::RP, T::RH ) )
        return -1;
    MvFieldParams p;
    p.width = fenc->width; p.height = fenc->height; p.n_planes = fenc->n_planes;
    p.blocks_x = fenc->width / BW; p.blocks_y = fenc->height / BH;
    p.tiles_x = ( fenc->width + TW - 1 ) / TW; p.tiles_y = ( fenc->height + TH - 1 ) / TH;
    p.n_tiles = p.tiles_x * p.tiles_y * fenc->n_planes;
    p.k_cands = k_cands; p.mv = d_mv; p.out = d_out;
    p.ref_origin = ref->d_origin; p.ref_stride = ref->stride; p.ref_plane_pitch = ref->plane_pitch;
    auto kern = mvfield_kernel<METRIC, BW, BH, TW, TH, R, NSTAGE, NWARPS>;
    size_t smem = NSTAGE * T::STAGE_BYTES + 128 + 64 + NSTAGE * 8 + 64;
    // the attribute is per device: remembered per context (one context = one device), not per process
    bool attr_set = false;
    for( const void *k : ctx->smem_attr_done ) attr_set |= k == (const void *)kern;
    if( !attr_set )
    {
        CU_CHECK( ctx, cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
        ctx->smem_attr_done.push_back( (const void *)kern );
    }
    int ctas_per_sm = (int)( ( 227 * 1024 ) / ( smem + 1024 ) );
    if( ctas_per_sm > 4 ) ctas_per_sm = 4;
    int grid = ctx->sm_count * ctas_per_sm;
    if( grid > p.n_tiles ) grid = p.n_tiles;
    kern<<<grid, NWARPS * 32, smem, ctx->stream>>>( tm_fenc, tm_ref, p );
    CU_LAUNCH_CHECK( ctx );
    return 0;
}

#define MVF_SIZE( M )                                                                               \
    switch( i_pixel )                                                                               \
    {                                                                                               \
        case X264CU_PIXEL_16x16: return launch_mvfield<M, 16, 16>( ctx, fenc, ref, k_cands, d_mv, d_out ); \
        case X264CU_PIXEL_16x8:  return launch_mvfield<M, 16, 8>( ctx, fenc, ref, k_cands, d_mv, d_out );  \
        case X264CU_PIXEL_8x16:  return launch_mvfield<M, 8, 16>( ctx, fenc, ref, k_cands, d_mv, d_out );  \
        case X264CU_PIXEL_8x8:   return launch_mvfield<M, 8, 8>( ctx, fenc, ref, k_cands, d_mv, d_out );   \
        case X264CU_PIXEL_8x4:   return launch_mvfield<M, 8, 4>( ctx, fenc, ref, k_cands, d_mv, d_out );   \
        case X264CU_PIXEL_4x8:   return launch_mvfield<M, 4, 8>( ctx, fenc, ref, k_cands, d_mv, d_out );   \
        case X264CU_PIXEL_4x4:   return launch_mvfield<M, 4, 4>( ctx, fenc, ref, k_cands, d_mv, d_out );   \
        case X264CU_PIXEL_4x16:  return launch_mvfield<M, 4, 16>( ctx, fenc, ref, k_cands, d_mv, d_out );  \
    }

extern "C" {

int x264cu_pixel_cmp_batch( x264cu_ctx_t *ctx, int metric, int i_pixel, const uint8_t *d_fenc, intptr_t fenc_stride,
                            const uint8_t *d_ref, intptr_t ref_stride, const x264cu_cand_t *d_cand, int n, int32_t *d_out )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( n <= 0 ) return 0;
    if( (unsigned)i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "bad block size index %d", i_pixel );
    return dispatch_batch<1>( ctx, metric, i_pixel, d_fenc, fenc_stride, d_ref, ref_stride, (const uint32_t *)d_cand, n, d_out );
}

int x264cu_pixel_cmp_x4_batch( x264cu_ctx_t *ctx, int metric, int i_pixel, int n_refs, const uint8_t *d_fenc,
                               intptr_t fenc_stride, const uint8_t *d_ref, intptr_t ref_stride,
                               const x264cu_cand_x4_t *d_cand, int n, int32_t *d_out )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( n <= 0 ) return 0;
    if( (unsigned)i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "bad block size index %d", i_pixel );
    if( metric != X264CU_SAD && metric != X264CU_SATD )
        return x264cu_fail( ctx, "x3/x4 entries exist for sad and satd only (common/pixel.h:100-107)" );
    if( n_refs == 4 )
        return dispatch_batch<4>( ctx, metric, i_pixel, d_fenc, fenc_stride, d_ref, ref_stride, (const uint32_t *)d_cand, n * 4, d_out );
    if( n_refs == 3 )
        return dispatch_batch<3>( ctx, metric, i_pixel, d_fenc, fenc_stride, d_ref, ref_stride, (const uint32_t *)d_cand, n * 4, d_out );
    return x264cu_fail( ctx, "n_refs must be 3 or 4" );
}

int x264cu_pixel_cmp_batch_host( x264cu_ctx_t *ctx, int metric, int i_pixel, const uint8_t *h_fenc, size_t fenc_bytes,
                                 intptr_t fenc_stride, const uint8_t *h_ref, size_t ref_bytes, intptr_t ref_stride,
                                 const x264cu_cand_t *h_cand, int n, int32_t *h_out )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( n <= 0 ) return 0;
    uint8_t *df = (uint8_t *)x264cu_scratch( ctx, 0, fenc_bytes + 16 );
    uint8_t *dr = (uint8_t *)x264cu_scratch( ctx, 1, ref_bytes + 16 );
    x264cu_cand_t *dc = (x264cu_cand_t *)x264cu_scratch( ctx, 2, (size_t)n * sizeof( x264cu_cand_t ) );
    int32_t *dout = (int32_t *)x264cu_scratch( ctx, 3, (size_t)n * sizeof( int32_t ) );
    if( !df || !dr || !dc || !dout ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( df, h_fenc, fenc_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( dr, h_ref, ref_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( dc, h_cand, (size_t)n * sizeof( x264cu_cand_t ), cudaMemcpyHostToDevice, ctx->stream ) );
    if( x264cu_pixel_cmp_batch( ctx, metric, i_pixel, df, fenc_stride, dr, ref_stride, dc, n, dout ) ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( h_out, dout, (size_t)n * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    return 0;
}

int x264cu_pixel_cmp_mvfield( x264cu_ctx_t *ctx, int metric, int i_pixel, const x264cu_planes_t *fenc,
                              const x264cu_planes_t *ref, int k_cands, const int16_t *d_mv, int32_t *d_out )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( !fenc || !ref || k_cands <= 0 ) return x264cu_fail( ctx, "mvfield: bad arguments" );
    if( (unsigned)i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "bad block size index %d", i_pixel );
    if( fenc->width != ref->width || fenc->height != ref->height || fenc->n_planes != ref->n_planes )
        return x264cu_fail( ctx, "mvfield: fenc and ref plane stacks differ in shape" );
    if( fenc->width % k_pixel_w[i_pixel] || fenc->height % k_pixel_h[i_pixel] )
        return x264cu_fail( ctx, "mvfield: %dx%d is not a multiple of the %dx%d block", fenc->width, fenc->height,
                            k_pixel_w[i_pixel], k_pixel_h[i_pixel] );
    if( fenc->n_planes <= 0 ) return 0;
#ifdef X264CU_TUNING        /* tuning build only (x264_b200/build.py --variant tuning -DX264CU_TUNING): not in the product library */
    if( metric == X264CU_SATD && i_pixel == X264CU_PIXEL_16x16 )
    {   // alternative tile / pipeline shapes for the headline kernel
        const char *e = getenv( "X264CU_MVF_CFG" );
        const int cfg = e ? atoi( e ) : 0;
        switch( cfg )
        {
            case 1: return launch_mvfield<M_SATD, 16, 16, 128, 128, 16, 2, 8>( ctx, fenc, ref, k_cands, d_mv, d_out );
            case 2: return launch_mvfield<M_SATD, 16, 16, 128, 128, 16, 2, 4>( ctx, fenc, ref, k_cands, d_mv, d_out );
            case 3: return launch_mvfield<M_SATD, 16, 16, 128, 64, 16, 2, 2>( ctx, fenc, ref, k_cands, d_mv, d_out );
            case 4: return launch_mvfield<M_SATD, 16, 16, 128, 64, 16, 3, 4>( ctx, fenc, ref, k_cands, d_mv, d_out );
            case 5: return launch_mvfield<M_SATD, 16, 16, 128, 64, 16, 1, 4>( ctx, fenc, ref, k_cands, d_mv, d_out );
            case 6: return launch_mvfield<M_SATD, 16, 16, 128, 32, 16, 2, 2>( ctx, fenc, ref, k_cands, d_mv, d_out );
            default: break;
        }
    }
#endif
    switch( metric )
    {
        case X264CU_SAD:  MVF_SIZE( M_SAD ) break;
        case X264CU_SSD:  MVF_SIZE( M_SSD ) break;
        case X264CU_SATD: MVF_SIZE( M_SATD ) break;
        case X264CU_SA8D:
            if( i_pixel == X264CU_PIXEL_16x16 ) return launch_mvfield<M_SA8D, 16, 16>( ctx, fenc, ref, k_cands, d_mv, d_out );
            if( i_pixel == X264CU_PIXEL_8x8 ) return launch_mvfield<M_SA8D, 8, 8>( ctx, fenc, ref, k_cands, d_mv, d_out );
            return x264cu_fail( ctx, "sa8d exists for 16x16 and 8x8 only (common/pixel.c:369-381)" );
    }
    return x264cu_fail( ctx, "bad metric %d", metric );
}

int x264cu_pixel_cmp_mvfield_host( x264cu_ctx_t *ctx, int metric, int i_pixel, const uint8_t *h_fenc_base,
                                   const uint8_t *h_ref_base, intptr_t stride, intptr_t plane_pitch, int width, int height,
                                   int n_planes, int k_cands, const int16_t *h_mv, int32_t *h_out )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( (unsigned)i_pixel >= X264CU_PIXEL_NB ) return x264cu_fail( ctx, "bad block size index %d", i_pixel );
    size_t bytes = (size_t)plane_pitch * n_planes;
    size_t n = (size_t)k_cands * n_planes * ( width / k_pixel_w[i_pixel] ) * ( height / k_pixel_h[i_pixel] );
    uint8_t *df = (uint8_t *)x264cu_scratch( ctx, 0, bytes + 256 );
    uint8_t *dr = (uint8_t *)x264cu_scratch( ctx, 1, bytes + 256 );
    int16_t *dmv = (int16_t *)x264cu_scratch( ctx, 2, n * 4 );
    int32_t *dout = (int32_t *)x264cu_scratch( ctx, 3, n * 4 );
    if( !df || !dr || !dmv || !dout ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( df, h_fenc_base, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( dr, h_ref_base, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    CU_CHECK( ctx, cudaMemcpyAsync( dmv, h_mv, n * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    intptr_t org = (intptr_t)X264CU_PAD * stride + X264CU_PAD;
    x264cu_planes_t pf = { df + org, stride, plane_pitch, width, height, n_planes };
    x264cu_planes_t pr = { dr + org, stride, plane_pitch, width, height, n_planes };
    if( x264cu_pixel_cmp_mvfield( ctx, metric, i_pixel, &pf, &pr, k_cands, dmv, dout ) ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( h_out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    return 0;
}

} // extern "C"
