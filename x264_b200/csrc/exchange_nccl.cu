// x264cu_exchange_nccl: the all-gather of a sharded stream (x264cu_slicetype_set_shard) for hosts written in C --
// ncclAllGather over NVLink / NVSwitch on the lookahead's exchange stream.  libnccl.so.2 is loaded at run time (dlopen), so the
// library itself keeps depending on nothing but libc / libstdc++ / libm; a process that never shards never loads NCCL.
#include "ctx.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>

namespace {

// the few NCCL declarations used (nccl.h: ncclUniqueId is 128 opaque bytes, ncclChar = 0)
struct NcclId { char internal[128]; };
typedef void *NcclComm;
typedef int ( *GetUniqueIdFn )( NcclId * );
typedef int ( *CommInitRankFn )( NcclComm *, int, NcclId, int );
typedef int ( *AllGatherFn )( const void *, void *, size_t, int, NcclComm, cudaStream_t );
typedef int ( *CommDestroyFn )( NcclComm );
typedef const char *( *GetErrorStringFn )( int );

}

struct x264cu_nccl
{
    x264cu_ctx *ctx = nullptr;
    void *dl = nullptr;
    GetUniqueIdFn get_id = nullptr; CommInitRankFn init_rank = nullptr; AllGatherFn all_gather = nullptr;
    CommDestroyFn destroy = nullptr; GetErrorStringFn err = nullptr;
    NcclComm comm = nullptr;
    bool own_comm = false;
    int rank = 0, world = 1;
    void *send[2] = { nullptr, nullptr }, *recv[2] = { nullptr, nullptr };     // per channel: searches / cost requests
    size_t cap[2] = { 0, 0 };
    long calls = 0;
    unsigned long long bytes = 0;
};

static int nccl_load( x264cu_ctx *ctx, x264cu_nccl *nc )
{
    const char *names[] = { getenv( "X264CU_NCCL_LIB" ), "libnccl.so.2", "libnccl.so" };
    for( const char *n : names )
        if( n && ( nc->dl = dlopen( n, RTLD_NOW | RTLD_LOCAL ) ) )
            break;
    if( !nc->dl ) return x264cu_fail( ctx, "nccl: cannot load libnccl.so.2 (%s)", dlerror() );
    nc->get_id = (GetUniqueIdFn)dlsym( nc->dl, "ncclGetUniqueId" );
    nc->init_rank = (CommInitRankFn)dlsym( nc->dl, "ncclCommInitRank" );
    nc->all_gather = (AllGatherFn)dlsym( nc->dl, "ncclAllGather" );
    nc->destroy = (CommDestroyFn)dlsym( nc->dl, "ncclCommDestroy" );
    nc->err = (GetErrorStringFn)dlsym( nc->dl, "ncclGetErrorString" );
    if( !nc->get_id || !nc->init_rank || !nc->all_gather || !nc->destroy || !nc->err )
        return x264cu_fail( ctx, "nccl: libnccl lacks an entry point" );
    return 0;
}

extern "C" {

void x264cu_nccl_close( x264cu_nccl_t *nc )
{
    if( !nc ) return;
    if( nc->ctx ) cudaSetDevice( nc->ctx->device );
    for( int c = 0; c < 2; c++ ) { cudaFree( nc->send[c] ); cudaFree( nc->recv[c] ); }
    if( nc->comm && nc->own_comm && nc->destroy ) nc->destroy( nc->comm );
    if( nc->dl ) dlclose( nc->dl );
    delete nc;
}

int x264cu_nccl_wrap( x264cu_ctx_t *ctx, void *nccl_comm, int rank, int world, x264cu_nccl_t **out )
{
    X264CU_ENTER( ctx );
    if( !ctx || !out || !nccl_comm || world < 1 || rank < 0 || rank >= world ) return -1;
    *out = nullptr;
    x264cu_nccl *nc = new x264cu_nccl;
    nc->ctx = ctx; nc->rank = rank; nc->world = world; nc->comm = nccl_comm;
    if( nccl_load( ctx, nc ) ) { x264cu_nccl_close( nc ); return -1; }
    *out = nc;
    return 0;
}

int x264cu_nccl_open( x264cu_ctx_t *ctx, int rank, int world, const char *id_path, int timeout_s, x264cu_nccl_t **out )
{
    X264CU_ENTER( ctx );
    if( !ctx || !out || !id_path || world < 1 || rank < 0 || rank >= world ) return -1;
    *out = nullptr;
    x264cu_nccl *nc = new x264cu_nccl;
    nc->ctx = ctx; nc->rank = rank; nc->world = world; nc->own_comm = true;
    if( nccl_load( ctx, nc ) ) { x264cu_nccl_close( nc ); return -1; }
    NcclId id;
    memset( &id, 0, sizeof( id ) );
    if( rank == 0 )
    {   // written under a temporary name and renamed: the other ranks never see half an id
        int r = nc->get_id( &id );
        if( r ) { x264cu_fail( ctx, "nccl: ncclGetUniqueId -> %s", nc->err( r ) ); x264cu_nccl_close( nc ); return -1; }
        std::string tmp = std::string( id_path ) + ".tmp";
        FILE *f = fopen( tmp.c_str(), "wb" );
        if( !f || fwrite( &id, sizeof( id ), 1, f ) != 1 || fclose( f ) || rename( tmp.c_str(), id_path ) )
        { x264cu_fail( ctx, "nccl: cannot write the id file %s", id_path ); x264cu_nccl_close( nc ); return -1; }
    }
    else
    {
        const time_t until = time( nullptr ) + ( timeout_s > 0 ? timeout_s : 60 );
        for( ;; )
        {
            FILE *f = fopen( id_path, "rb" );
            const bool ok = f && fread( &id, sizeof( id ), 1, f ) == 1;
            if( f ) fclose( f );
            if( ok ) break;
            if( time( nullptr ) > until ) { x264cu_fail( ctx, "nccl: no id file %s after %d s", id_path, timeout_s ); x264cu_nccl_close( nc ); return -1; }
            usleep( 20000 );
        }
    }
    int r = nc->init_rank( &nc->comm, world, id, rank );
    if( r ) { x264cu_fail( ctx, "nccl: ncclCommInitRank -> %s", nc->err( r ) ); nc->comm = nullptr; x264cu_nccl_close( nc ); return -1; }
    *out = nc;
    return 0;
}

// x264cu_exchange_fn: phases 0 / 1 = buffers / all-gather of the search results, 2 / 3 = the same for the cost requests
int x264cu_exchange_nccl( void *user, int phase, size_t bytes_per_rank, void **d_send, void **d_recv, void *stream )
{
    x264cu_nccl *nc = (x264cu_nccl *)user;
    if( !nc || !d_send || !d_recv || phase < 0 || phase > 3 ) return -1;
    x264cu_ctx *ctx = nc->ctx;
    cudaSetDevice( ctx->device );
    const int ch = phase >> 1;
    if( !( phase & 1 ) )
    {
        const size_t want = bytes_per_rank < 16 ? 16 : bytes_per_rank;
        if( nc->cap[ch] < want )
        {   // copies out of the old receive buffer may still be in flight on the exchange stream
            CU_CHECK( ctx, cudaStreamSynchronize( (cudaStream_t)stream ) );
            cudaFree( nc->send[ch] ); cudaFree( nc->recv[ch] );
            nc->send[ch] = nc->recv[ch] = nullptr; nc->cap[ch] = 0;
            const size_t grow = want + want / 4;
            CU_CHECK( ctx, cudaMalloc( &nc->send[ch], grow ) );
            CU_CHECK( ctx, cudaMalloc( &nc->recv[ch], grow * nc->world ) );
            nc->cap[ch] = grow;
        }
        *d_send = nc->send[ch]; *d_recv = nc->recv[ch];
        return 0;
    }
    if( *d_send != nc->send[ch] || *d_recv != nc->recv[ch] || bytes_per_rank > nc->cap[ch] )
        return x264cu_fail( ctx, "nccl: exchange called with foreign buffers" );
    if( bytes_per_rank )
    {
        int r = nc->all_gather( nc->send[ch], nc->recv[ch], bytes_per_rank, /* ncclChar */ 0, nc->comm, (cudaStream_t)stream );
        if( r ) return x264cu_fail( ctx, "nccl: ncclAllGather -> %s", nc->err( r ) );
    }
    nc->calls++; nc->bytes += (unsigned long long)bytes_per_rank * nc->world;
    return 0;
}

long x264cu_nccl_calls( x264cu_nccl_t *nc, unsigned long long *bytes )
{
    if( !nc ) return -1;
    if( bytes ) *bytes = nc->bytes;
    return nc->calls;
}

} // extern "C"
