/* TEST / BENCH INFRASTRUCTURE: the product's host-side slice-type logic (x264_b200/csrc/slicetype.c) driving the UNMODIFIED
 * reference's own slicetype_frame_cost (through oracle/_ref/libx264ref.so) instead of the GPU lookahead.  This is the CPU
 * arm of bench.py's lookahead workload: identical decision workload, reference cost function, host cores. */
#include "../../include/x264_b200.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

void *xref_open( int width, int height, const char *preset, const char *opts, int verbose );
void  xref_close( void *h );
void *xref_la_new( void *h, int n );
int   xref_la_set_frame( void *la, int idx, const uint8_t *luma, intptr_t stride, const uint16_t *q );
int   xref_la_frame_cost( void *la, int p0, int p1, int b );
void  xref_la_get( void *la, int idx, int what, int i, int j, void *out );
void  xref_la_free( void *la );
void  xref_la_remap( void *la, const int *frames, int p0, int p1 );

struct x264cu_lookahead { void *h, *la; int n_slots; };

static const char *g_preset = "medium";
static const char *g_opts = "";
void slicetype_ref_glue_config( const char *preset, const char *opts ) { g_preset = preset; g_opts = opts; }

int x264cu_lookahead_open( x264cu_ctx_t *ctx, const x264cu_lookahead_params_t *q, x264cu_lookahead_t **out )
{
    (void)ctx;
    x264cu_lookahead_t *la = calloc( 1, sizeof( *la ) );
    la->h = xref_open( q->width, q->height, g_preset, g_opts, 0 );
    if( !la->h ) { free( la ); return -1; }
    la->n_slots = q->n_slots;
    la->la = xref_la_new( la->h, q->n_slots + 300 );
    *out = la;
    return 0;
}
void x264cu_lookahead_close( x264cu_lookahead_t *la )
{
    if( !la ) return;
    xref_la_free( la->la );
    xref_close( la->h );
    free( la );
}
int x264cu_lookahead_frame_put( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t stride, const uint16_t *q )
{
    return xref_la_set_frame( la->la, slot, h_luma, stride, q );
}
int x264cu_lookahead_frame_put_device( x264cu_lookahead_t *la, int slot, const uint8_t *d, intptr_t st, const uint16_t *q )
{
    (void)la; (void)slot; (void)d; (void)st; (void)q; return -1;
}
int x264cu_lookahead_search_batch( x264cu_lookahead_t *la, int n, const int *a, const int *b, const int *c, const int *d )
{
    (void)la; (void)n; (void)a; (void)b; (void)c; (void)d; return 0;
}
/* answering cost requests ahead of time is a scheduling matter too: here every request is computed when it is made */
int x264cu_lookahead_finalize_batch( x264cu_lookahead_t *la, int n, const int *b, const int *p0, const int *p1, const int *d0, const int *d1 )
{
    (void)la; (void)n; (void)b; (void)p0; (void)p1; (void)d0; (void)d1; return 0;
}
int x264cu_lookahead_finalize_batch_sharded( x264cu_lookahead_t *la, int n, const int *b, const int *p0, const int *p1, const int *d0, const int *d1,
                                             const int *owner, int rank, int world, x264cu_exchange_fn fn, void *user )
{
    (void)la; (void)n; (void)b; (void)p0; (void)p1; (void)d0; (void)d1; (void)owner; (void)rank; (void)world; (void)fn; (void)user; return 0;
}
int x264cu_lookahead_frame_cost( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int *score )
{
    /* the reference takes a frames[] array of frame pointers: map indices p0..p1 to the slots' frames */
    xref_la_remap( la->la, frames, p0, p1 );
    *score = xref_la_frame_cost( la->la, p0, p1, b );
    return 0;
}
int x264cu_lookahead_get_cost_est( x264cu_lookahead_t *la, int slot, int i0, int i1, int *ce, int *ceaq, int *imb )
{
    int e[3];
    xref_la_get( la->la, slot + 300, 4, i0, i1, e );      /* +300: slot table (see ref_shim.c xref_la_get) */
    if( ce ) *ce = e[0];
    if( ceaq ) *ceaq = e[1];
    if( imb ) *imb = e[2];
    return 0;
}
void x264cu_lookahead_set_async_upload( x264cu_lookahead_t *la, int on ) { (void)la; (void)on; }
int x264cu_lookahead_weight_trivial( x264cu_lookahead_t *la, int a, int b ) { (void)la; (void)a; (void)b; return 0; }

/* ---- MB-tree entries, served by the reference's own macroblock_tree_propagate / _finish (all durations left at the frame
 * pool's initial 0, i.e. clamped to the same 0.01 s: the factors of a constant frame rate) ---- */
void xref_la_mbtree_reset( void *la, int idx );
void xref_la_mbtree_swap( void *la, int a, int b );
void xref_la_mbtree_propagate( void *la, float average_duration, int p0, int p1, int b, int referenced );
void xref_la_mbtree_finish( void *la, int idx, float average_duration, int ref0_distance );
void xref_la_get_mbtree( void *la, int idx, int what, int i, void *out );
int x264cu_lookahead_frame_set_qp_offset_aq( x264cu_lookahead_t *la, int slot, const float *aq ) { (void)la; (void)slot; (void)aq; return 0; }
int x264cu_lookahead_mbtree_reset( x264cu_lookahead_t *la, int slot ) { xref_la_mbtree_reset( la->la, slot + 300 ); return 0; }
int x264cu_lookahead_mbtree_swap( x264cu_lookahead_t *la, int a, int b ) { xref_la_mbtree_swap( la->la, a + 300, b + 300 ); return 0; }
int x264cu_lookahead_mbtree_propagate( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int referenced, float fps_factor )
{
    (void)fps_factor;
    xref_la_remap( la->la, frames, p0, p1 );
    xref_la_mbtree_propagate( la->la, 0.0f, p0, p1, b, referenced );
    return 0;
}
int x264cu_lookahead_mbtree_finish( x264cu_lookahead_t *la, int slot, int fps_factor, int ref0_distance, float strength )
{
    (void)fps_factor; (void)strength;
    xref_la_mbtree_finish( la->la, slot + 300, 0.0f, ref0_distance );
    return 0;
}
int x264cu_lookahead_get_qp_offset( x264cu_lookahead_t *la, int slot, float *out ) { xref_la_get_mbtree( la->la, slot + 300, 0, 0, out ); return 0; }
int xref_la_frame_cost_recalculate_at( void *la, int idx, int i0, int i1, int b_type, int *rows );
int x264cu_lookahead_frame_cost_recalculate( x264cu_lookahead_t *la, int slot, int i0, int i1, int b_type, int *score, int32_t *rows )
{
    *score = xref_la_frame_cost_recalculate_at( la->la, slot + 300, i0, i1, b_type, rows );
    return 0;
}
int x264cu_lookahead_get_row_satds( x264cu_lookahead_t *la, int slot, int i0, int i1, int32_t *rows )
{
    xref_la_get( la->la, slot + 300, 5, i0, i1, rows );
    return 0;
}

/* sharded-stream entries: nothing travels in the CPU harness (every search is computed where it is asked for) */
size_t x264cu_lookahead_search_bytes( x264cu_lookahead_t *la ) { (void)la; return 8; }
void *x264cu_lookahead_exchange_stream( x264cu_lookahead_t *la ) { (void)la; return 0; }
int x264cu_lookahead_export_search( x264cu_lookahead_t *la, int slot, int list, int dist, void *d ) { (void)la; (void)slot; (void)list; (void)dist; (void)d; return 0; }
int x264cu_lookahead_import_search( x264cu_lookahead_t *la, int slot, int list, int dist, const void *d ) { (void)la; (void)slot; (void)list; (void)dist; (void)d; return 0; }
int x264cu_lookahead_import_done( x264cu_lookahead_t *la ) { (void)la; return 0; }
int x264cu_lookahead_frame_put_i420( x264cu_lookahead_t *la, int slot, const uint8_t *y, intptr_t ys, const uint8_t *cb, const uint8_t *cr,
                                     intptr_t cs, int aq_mode, float aq_strength )
{
    (void)la; (void)slot; (void)y; (void)ys; (void)cb; (void)cr; (void)cs; (void)aq_mode; (void)aq_strength;
    return -1;          /* the bench's CPU arm feeds luma only */
}
