// Context, error reporting and device-memory helpers of the C ABI (include/x264_b200.h).
// Takes the place of common/opencl.c's library loading / device selection / teardown
// (x264_opencl_load_library opencl.c:53, x264_opencl_lookahead_init :411, _delete :596): pick the
// device, create the single in-order stream all work is enqueued on, release everything at close.
#include "ctx.h"
#include <stdarg.h>
#include <string.h>

static thread_local std::string g_open_error;    // the reason of this thread's last failed x264cu_open

int x264cu_fail( x264cu_ctx *ctx, const char *fmt, ... )
{
    char buf[1024];
    va_list ap;
    va_start( ap, fmt );
    vsnprintf( buf, sizeof( buf ), fmt, ap );
    va_end( ap );
    if( ctx ) ctx->err = buf; else g_open_error = buf;
    return -1;
}

void *x264cu_scratch( x264cu_ctx *ctx, int slot, size_t bytes )
{
    if( ctx->scratch_bytes[slot] < bytes )
    {
        if( ctx->scratch[slot] ) cudaFree( ctx->scratch[slot] );
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        size_t want = bytes + ( bytes >> 3 ) + 4096;
        if( cudaMalloc( &ctx->scratch[slot], want ) != cudaSuccess )
        {
            x264cu_fail( ctx, "cudaMalloc(%zu) for scratch slot %d failed", want, slot );
            return nullptr;
        }
        ctx->scratch_bytes[slot] = want;
    }
    return ctx->scratch[slot];
}

extern "C" {

int x264cu_open( x264cu_ctx_t **out, int device )
{
    if( !out ) return x264cu_fail( nullptr, "x264cu_open: NULL out pointer" );
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount( &n );
    if( e != cudaSuccess || n <= 0 )
        return x264cu_fail( nullptr, "x264cu_open: no CUDA device (%s); this backend has no CPU fallback",
                            e != cudaSuccess ? cudaGetErrorString( e ) : "device count 0" );
    if( device < 0 || device >= n )
        return x264cu_fail( nullptr, "x264cu_open: device %d out of range (have %d)", device, n );
    cudaDeviceProp prop;
    if( ( e = cudaGetDeviceProperties( &prop, device ) ) != cudaSuccess )
        return x264cu_fail( nullptr, "x264cu_open: cudaGetDeviceProperties -> %s", cudaGetErrorString( e ) );
    if( prop.major < 10 )
        return x264cu_fail( nullptr, "x264cu_open: device %d is sm_%d%d; this library only carries sm_100a code",
                            device, prop.major, prop.minor );
    if( ( e = cudaSetDevice( device ) ) != cudaSuccess )
        return x264cu_fail( nullptr, "x264cu_open: cudaSetDevice -> %s", cudaGetErrorString( e ) );
    x264cu_ctx *ctx = new x264cu_ctx;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->hbm_bytes = prop.totalGlobalMem;
    // highest priority: what the calling thread waits for (cost requests) is placed ahead of the lookahead's background
    // streams (prefetched searches) whenever an SM has room
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange( &prio_lo, &prio_hi );
    if( ( e = cudaStreamCreateWithPriority( &ctx->stream, cudaStreamNonBlocking, prio_hi ) ) != cudaSuccess )
    {
        x264cu_fail( nullptr, "x264cu_open: cudaStreamCreate -> %s", cudaGetErrorString( e ) );
        delete ctx;
        return -1;
    }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres );
    if( e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess )
    {
        x264cu_fail( nullptr, "x264cu_open: cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString( e ) );
        cudaStreamDestroy( ctx->stream );
        delete ctx;
        return -1;
    }
    ctx->encode_tiled = reinterpret_cast<decltype( ctx->encode_tiled )>( fn );
    *out = ctx;
    return 0;
}

void x264cu_lookahead_close_internal( x264cu_ctx *ctx );

void x264cu_close( x264cu_ctx_t *ctx )
{
    if( !ctx ) return;
    cudaSetDevice( ctx->device );
    cudaStreamSynchronize( ctx->stream );
    x264cu_lookahead_close_internal( ctx );
    for( int i = 0; i < 16; i++ )
        if( ctx->scratch[i] ) cudaFree( ctx->scratch[i] );
    if( ctx->ev0 ) { cudaEventDestroy( ctx->ev0 ); cudaEventDestroy( ctx->ev1 ); }
    cudaStreamDestroy( ctx->stream );
    delete ctx;
}

const char *x264cu_strerror( x264cu_ctx_t *ctx ) { return ctx ? ctx->err.c_str() : g_open_error.c_str(); }

int x264cu_device_info( x264cu_ctx_t *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes )
{
    if( !ctx ) return -1;
    if( sm_count ) *sm_count = ctx->sm_count;
    if( cc_major ) *cc_major = ctx->cc_major;
    if( cc_minor ) *cc_minor = ctx->cc_minor;
    if( hbm_bytes ) *hbm_bytes = ctx->hbm_bytes;
    return 0;
}

void *x264cu_stream( x264cu_ctx_t *ctx ) { return ctx ? (void *)ctx->stream : nullptr; }

int x264cu_sync( x264cu_ctx_t *ctx )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    for( cudaStream_t st : ctx->aux_streams )          // uploads / prefetched searches of the lookahead (x264_opencl_flush)
        CU_CHECK( ctx, cudaStreamSynchronize( st ) );
    return 0;
}

int x264cu_timer_start( x264cu_ctx_t *ctx )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    if( !ctx->ev0 )
    {
        CU_CHECK( ctx, cudaEventCreate( &ctx->ev0 ) );
        CU_CHECK( ctx, cudaEventCreate( &ctx->ev1 ) );
    }
    CU_CHECK( ctx, cudaEventRecord( ctx->ev0, ctx->stream ) );
    return 0;
}

int x264cu_timer_stop( x264cu_ctx_t *ctx, float *elapsed_ms )
{
    X264CU_ENTER( ctx );
    if( !ctx || !ctx->ev0 ) return -1;
    CU_CHECK( ctx, cudaEventRecord( ctx->ev1, ctx->stream ) );
    CU_CHECK( ctx, cudaEventSynchronize( ctx->ev1 ) );
    float ms = 0.f;
    CU_CHECK( ctx, cudaEventElapsedTime( &ms, ctx->ev0, ctx->ev1 ) );
    if( elapsed_ms ) *elapsed_ms = ms;
    return 0;
}

uint64_t x264cu_launch_count( x264cu_ctx_t *ctx ) { return ctx ? ctx->launches : 0; }

void *x264cu_malloc( x264cu_ctx_t *ctx, size_t bytes )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return nullptr;
    void *p = nullptr;
    cudaSetDevice( ctx->device );
    if( cudaMalloc( &p, bytes ) != cudaSuccess )
    {
        x264cu_fail( ctx, "x264cu_malloc(%zu) failed", bytes );
        return nullptr;
    }
    return p;
}

void x264cu_free( x264cu_ctx_t *ctx, void *p )
{
    X264CU_ENTER( ctx );
    if( ctx && p ) { cudaSetDevice( ctx->device ); cudaFree( p ); }
}

void *x264cu_malloc_host( x264cu_ctx_t *ctx, size_t bytes )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return nullptr;
    void *p = nullptr;
    if( cudaMallocHost( &p, bytes ) != cudaSuccess )
    {
        x264cu_fail( ctx, "x264cu_malloc_host(%zu) failed", bytes );
        return nullptr;
    }
    return p;
}

void x264cu_free_host( x264cu_ctx_t *ctx, void *p ) { if( ctx && p ) cudaFreeHost( p ); }

int x264cu_memcpy_h2d( x264cu_ctx_t *ctx, void *d, const void *h, size_t bytes )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( d, h, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    return 0;
}

int x264cu_memcpy_d2h( x264cu_ctx_t *ctx, void *h, const void *d, size_t bytes )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    CU_CHECK( ctx, cudaMemcpyAsync( h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
    CU_CHECK( ctx, cudaStreamSynchronize( ctx->stream ) );
    return 0;
}

int x264cu_memset( x264cu_ctx_t *ctx, void *d, int value, size_t bytes )
{
    X264CU_ENTER( ctx );
    if( !ctx ) return -1;
    CU_CHECK( ctx, cudaMemsetAsync( d, value, bytes, ctx->stream ) );
    return 0;
}

} // extern "C"
