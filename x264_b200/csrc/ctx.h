// Internal context shared by the translation units of libx264_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/x264_b200.h"

struct x264cu_ctx
{
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    std::string err;
    // driver entry point for TMA descriptors (no link-time libcuda dependency)
    CUresult (*encode_tiled)( CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill ) = nullptr;
    // scratch for the *_host entry points (grown on demand)
    void *scratch[16] = {};
    size_t scratch_bytes[16] = {};
    bool aq_tables = false;                      // x264cu_adaptive_quant_frame's constant tables uploaded
    int me_tab_lambda = -1, me_tab_range = -1;   // which cost_mv table scratch slot 5 currently holds (x264cu_me_search_batch)
    std::vector<int> me_tabs_lambdas; int me_tabs_range = -1;   // ... and the per-lambda tables of slot 11 (x264cu_me_search_frame)
    struct x264cu_lookahead *lookahead = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaStream_t> aux_streams;   // streams of live lookahead objects: x264cu_sync waits for them too
    std::vector<const void *> smem_attr_done; // kernels whose dynamic shared-memory limit has been raised on THIS context's device
};

int  x264cu_fail( x264cu_ctx *ctx, const char *fmt, ... );
void *x264cu_scratch( x264cu_ctx *ctx, int slot, size_t bytes );
int  x264cu_frame_init_lowres_on( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, int width, int height,
                                  uint8_t *const d_lowres[4], intptr_t lowres_stride );

int  x264cu_adaptive_quant_frame_on( x264cu_ctx *ctx, cudaStream_t stream, const uint8_t *d_luma, intptr_t luma_stride, const uint8_t *d_cb,
                                     const uint8_t *d_cr, intptr_t chroma_stride, int width, int height, int aq_mode, float aq_strength,
                                     float *d_qp_offset_aq, uint16_t *d_inv_qscale, float *d_q4, unsigned long long *d_stats );

// The current device is per-thread state: every public entry point selects its context's device first, so that a context may be
// used from a thread other than the one that opened it and contexts on different GPUs may live in one process.
#define X264CU_ENTER( ctx ) do { if( ctx ) cudaSetDevice( ( ctx )->device ); } while( 0 )
#define X264CU_ENTER_LA( la ) do { if( la ) cudaSetDevice( x264cu_lookahead_ctx( la )->device ); } while( 0 )
struct x264cu_lookahead;
x264cu_ctx *x264cu_lookahead_ctx( struct x264cu_lookahead *la );

#define CU_CHECK( ctx, call )                                                                      \
    do {                                                                                           \
        cudaError_t e_ = ( call );                                                                 \
        if( e_ != cudaSuccess )                                                                    \
            return x264cu_fail( ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString( e_ ) ); \
    } while( 0 )

#define CU_LAUNCH_CHECK( ctx )                                                                     \
    do {                                                                                           \
        ( ctx )->launches++;                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if( e_ != cudaSuccess )                                                                    \
            return x264cu_fail( ctx, "%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString( e_ ) ); \
    } while( 0 )
