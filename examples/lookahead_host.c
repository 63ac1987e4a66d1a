/* A host written in plain C -- what the reference (a C program) would look like on its side of the boundary: it includes
 * include/x264_b200.h, links libx264_b200.so, and drives the lookahead the way x264_encoder_encode does
 * (x264_lookahead_put_frame / x264_lookahead_get_frames -> x264cu_slicetype_step), reading MB-tree's quantiser offsets for the
 * pictures it would encode.  No CUDA headers, no C++.
 *
 *   gcc -O2 -Iinclude examples/lookahead_host.c -o lookahead_host -Lx264_b200/csrc -lx264_b200 -Wl,-rpath,$PWD/x264_b200/csrc
 *   ./lookahead_host WIDTH HEIGHT FRAMES [raw 8-bit I420 file]      (without a file: a synthetic moving texture with a cut)
 *   ./lookahead_host --ranks N WIDTH HEIGHT FRAMES [file]            ONE stream sharded over N GPUs: N processes (fork), one per
 *                                                                    GPU, the all-gathers done by x264cu_exchange_nccl (NCCL over
 *                                                                    NVLink); every rank takes the same decisions, rank 0 prints
 *   options before WIDTH: --bframes B --b-adapt A --rc-lookahead L   (defaults 3 / 1 / 20; BASELINE configs[3] is 16 / 2 / 250)
 *
 * Raw I420 in (Y, Cb, Cr planes per picture, as x264's raw demuxer reads them, input/raw.c:43-170): adaptive quantisation
 * (aq-mode 1), lowres planes, lookahead, slice-type decision and MB-tree all run on the device (x264cu_slicetype_step_i420).
 *
 * Prints one line per picture in coded order: "frame <display index> type <I|P|B|b...> qp_offset_mean <f>".
 * tests/test_gpu_c_host.py builds and runs it and compares its output with the Python binding's. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#include "x264_b200.h"

static const char *type_name( int t )
{
    switch( t )
    {
        case X264CU_TYPE_IDR:  return "IDR";
        case X264CU_TYPE_I:    return "I";
        case X264CU_TYPE_P:    return "P";
        case X264CU_TYPE_BREF: return "Bref";
        case X264CU_TYPE_B:    return "B";
        default:               return "?";
    }
}

/* deterministic synthetic content: a smooth texture translated a little every picture, a hard cut at two thirds */
static void synth( uint8_t *luma, int w, int h, int i, int n )
{
    int cut = i >= ( 2 * n ) / 3;
    int ox = ( i * 3 ) % 17 + ( cut ? 91 : 0 ), oy = ( i * 2 ) % 11 + ( cut ? 47 : 0 );
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            int u = x + ox, v = y + oy;
            int val = 128 + ( ( ( u * 7 + v * 3 ) % 64 ) - 32 ) + ( ( ( u / 8 + v / 8 ) & 1 ) ? 24 : -24 ) + ( cut ? ( u * v ) % 23 : ( u + 2 * v ) % 13 );
            luma[(size_t)y * w + x] = (uint8_t)( val < 0 ? 0 : val > 255 ? 255 : val );
        }
}

int main( int argc, char **argv )
{
    int ranks = 1, rank = 0, bframes = 3, b_adapt = 1, rc_lookahead = 20;
    while( argc > 2 && argv[1][0] == '-' )
    {
        if( !strcmp( argv[1], "--ranks" ) ) ranks = atoi( argv[2] );
        else if( !strcmp( argv[1], "--bframes" ) ) bframes = atoi( argv[2] );
        else if( !strcmp( argv[1], "--b-adapt" ) ) b_adapt = atoi( argv[2] );
        else if( !strcmp( argv[1], "--rc-lookahead" ) ) rc_lookahead = atoi( argv[2] );
        else break;
        argv += 2; argc -= 2;
    }
    if( argc < 4 || ranks < 1 || ranks > 8 )
    {
        fprintf( stderr, "usage: %s [--ranks N] [--bframes B] [--b-adapt A] [--rc-lookahead L] WIDTH HEIGHT FRAMES [pictures.i420]\n", argv[0] );
        return 2;
    }
    const int w = atoi( argv[1] ), h = atoi( argv[2] );
    int n = atoi( argv[3] );
    char id_path[64];
    snprintf( id_path, sizeof( id_path ), "/tmp/x264cu_nccl_id_%d", (int)getpid() );
    pid_t kids[8];
    for( int r = 1; r < ranks; r++ )           /* one process per GPU; the parent is rank 0 */
        if( !( kids[r] = fork() ) ) { rank = r; break; }
    FILE *in = argc > 4 ? fopen( argv[4], "rb" ) : NULL;
    if( argc > 4 && !in ) { perror( argv[4] ); return 2; }

    x264cu_ctx_t *ctx;
    if( x264cu_open( &ctx, rank ) ) { fprintf( stderr, "x264cu_open: %s\n", x264cu_strerror( NULL ) ); return 1; }   /* no CPU fallback */
    x264cu_nccl_t *nc = NULL;
    if( ranks > 1 && x264cu_nccl_open( ctx, rank, ranks, id_path, 120, &nc ) ) { fprintf( stderr, "rank %d: %s\n", rank, x264cu_strerror( ctx ) ); return 1; }

    x264cu_slicetype_params_t p;
    memset( &p, 0, sizeof( p ) );
    p.la.width = w; p.la.height = h;
    p.la.subpel_refine = 7; p.la.me_method = X264CU_ME_HEX; p.la.me_range = 16; p.la.mv_range = 512;     /* preset medium */
    p.la.bframes = bframes; p.la.weighted_bipred = 1; p.la.aq_mode = 1; p.la.mb_tree = 1; p.la.weighted_pred = 0;
    p.keyint_max = 250; p.keyint_min = 25; p.scenecut_threshold = 40; p.b_adapt = b_adapt; p.b_pyramid = 2; p.rc_lookahead = rc_lookahead;
    p.psy = 0; p.frame_reference = 3; p.fps_num = 25; p.fps_den = 1; p.qcompress = 0.6f; p.aq_strength = 1.0f;
    x264cu_slicetype_t *st;
    if( x264cu_slicetype_open( ctx, &p, &st ) ) { fprintf( stderr, "slicetype_open: %s\n", x264cu_strerror( ctx ) ); return 1; }
    if( ranks > 1 && x264cu_slicetype_set_shard( st, rank, ranks, x264cu_exchange_nccl, nc ) ) { fprintf( stderr, "set_shard failed\n" ); return 1; }

    const int mbs = ( ( w + 15 ) / 16 ) * ( ( h + 15 ) / 16 );
    const int cw = ( w + 1 ) / 2, ch = ( h + 1 ) / 2;
    uint8_t *luma = x264cu_malloc_host( ctx, (size_t)w * h + 2 * (size_t)cw * ch );     /* page-locked: read in place by the copy engine */
    uint8_t *cb = luma ? luma + (size_t)w * h : NULL, *cr = cb ? cb + (size_t)cw * ch : NULL;
    float *qp = malloc( sizeof( float ) * mbs );
    if( !luma || !qp ) return 1;
    int fed = 0, frame, type, rc = 0;
    for( ;; )
    {
        const uint8_t *pic = NULL;
        if( fed < n )
        {
            const size_t pic_bytes = (size_t)w * h + 2 * (size_t)cw * ch;
            if( in ) { if( fread( luma, 1, pic_bytes, in ) != pic_bytes ) { n = fed; continue; } }
            else { synth( luma, w, h, fed, n ); memset( cb, 128, 2 * (size_t)cw * ch ); }
            pic = luma;
            fed++;
        }
        if( x264cu_slicetype_step_i420( st, pic, w, cb, cr, cw, &frame, &type ) ) { fprintf( stderr, "step: %s\n", x264cu_strerror( ctx ) ); rc = 1; break; }
        if( frame >= 0 )
        {
            double mean = 0;
            if( type != X264CU_TYPE_B && type != X264CU_TYPE_BREF && !x264cu_slicetype_get_qp_offset( st, frame, qp ) )
            {
                for( int i = 0; i < mbs; i++ ) mean += qp[i];
                mean /= mbs;
            }
            if( rank == 0 )
                printf( "frame %d type %s qp_offset_mean %.4f\n", frame, type_name( type ), mean );
            else
                fprintf( stderr, "rank %d frame %d type %s qp_offset_mean %.4f\n", rank, frame, type_name( type ), mean );
        }
        else if( !pic )
            break;                                                     /* flushing and nothing left */
    }
    free( qp );
    x264cu_free_host( ctx, luma );
    x264cu_slicetype_close( st );
    if( nc )
    {
        unsigned long long bytes = 0;
        long calls = x264cu_nccl_calls( nc, &bytes );
        fprintf( stderr, "rank %d: %ld NCCL all-gathers, %.1f MB gathered\n", rank, calls, bytes / 1e6 );
        x264cu_nccl_close( nc );
    }
    x264cu_close( ctx );
    if( in ) fclose( in );
    if( rank == 0 && ranks > 1 )
    {
        for( int r = 1; r < ranks; r++ )
        {
            int status = 0;
            waitpid( kids[r], &status, 0 );
            if( !WIFEXITED( status ) || WEXITSTATUS( status ) ) rc = 1;
        }
        remove( id_path );
    }
    return rc;
}
