/* TEST INFRASTRUCTURE: lets the product's host-side slice-type logic (x264_b200/csrc/slicetype.c) run on a machine
 * without a GPU by standing in for the five x264cu_lookahead_* entry points it calls with the CPU oracle
 * (oracle/oracle_lookahead.c).  Built by tests/_libs.py into tests/_build/libslicetype_oracle.so; never shipped. */
#include "../../include/x264_b200.h"
#include "../../oracle/oracle.h"
#include <stdlib.h>
#include <string.h>

int orc_la_frame_cost( const orc_la_params_t *p, const uint16_t *cost_mv_centre, orc_la_frame_t **frames, int p0, int p1, int b );

struct x264cu_lookahead
{
    orc_la_params_t p;
    int n_slots;
    orc_la_frame_t **slots;
    uint16_t *tab;
    int tab_len;
};

int x264cu_lookahead_open( x264cu_ctx_t *ctx, const x264cu_lookahead_params_t *q, x264cu_lookahead_t **out )
{
    (void)ctx;
    x264cu_lookahead_t *la = calloc( 1, sizeof( *la ) );
    la->p.width = q->width; la->p.height = q->height;
    la->p.mb_width = ( q->width + 15 ) >> 4; la->p.mb_height = ( q->height + 15 ) >> 4;
    la->p.subpel_refine = q->subpel_refine; la->p.me_method = q->me_method; la->p.me_range = q->me_range;
    la->p.mv_range = q->mv_range; la->p.bframes = q->bframes; la->p.bframe_bias = q->bframe_bias;
    la->p.weighted_bipred = q->weighted_bipred; la->p.aq_mode = q->aq_mode; la->p.vbv = q->vbv;
    la->p.do_edges = q->mb_tree || q->vbv;
    la->p.weighted_pred = q->weighted_pred;
    la->n_slots = q->n_slots;
    la->slots = calloc( q->n_slots, sizeof( void * ) );
    la->tab_len = 2 * 4 * q->mv_range;
    la->tab = malloc( ( 2 * la->tab_len + 1 ) * 2 );
    orc_cost_mv_table( la->tab, la->tab_len, 1 );
    *out = la;
    return 0;
}

void x264cu_lookahead_close( x264cu_lookahead_t *la )
{
    if( !la ) return;
    for( int i = 0; i < la->n_slots; i++ ) orc_la_frame_delete( la->slots[i] );
    free( la->slots ); free( la->tab ); free( la );
}

void orc_la_frame_set_qscale( orc_la_frame_t *f, const uint16_t *inv_qscale );

int x264cu_lookahead_frame_put( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride, const uint16_t *q )
{
    orc_la_frame_delete( la->slots[slot] );
    la->slots[slot] = orc_la_frame_new( &la->p, h_luma, luma_stride );
    if( q ) orc_la_frame_set_qscale( la->slots[slot], q );
    return 0;
}

int x264cu_lookahead_frame_cost( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int *score )
{
    orc_la_frame_t *fr[300];
    for( int i = p0; i <= p1; i++ ) fr[i] = la->slots[frames[i]];
    *score = orc_la_frame_cost( &la->p, la->tab + la->tab_len, fr, p0, p1, b );
    return 0;
}

int x264cu_lookahead_get_cost_est( x264cu_lookahead_t *la, int slot, int i0, int i1, int *ce, int *ceaq, int *imb )
{
    orc_la_frame_t *f = la->slots[slot];
    if( ce ) *ce = f->cost_est[i0][i1];
    if( ceaq ) *ceaq = f->cost_est_aq[i0][i1];
    if( imb ) *imb = f->intra_mbs[i0];
    return 0;
}

int x264cu_lookahead_frame_put_device( x264cu_lookahead_t *la, int slot, const uint8_t *d, intptr_t st, const uint16_t *q )
{
    (void)la; (void)slot; (void)d; (void)st; (void)q;
    return -1;          /* no device in the CPU harness */
}

/* prefetching is a pure scheduling hint: the oracle computes every search on demand */
int x264cu_lookahead_search_batch( x264cu_lookahead_t *la, int n, const int *fenc, const int *ref, const int *list, const int *dist )
{
    (void)la; (void)n; (void)fenc; (void)ref; (void)list; (void)dist;
    return 0;
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
void x264cu_lookahead_set_async_upload( x264cu_lookahead_t *la, int on ) { (void)la; (void)on; }
int x264cu_lookahead_weight_trivial( x264cu_lookahead_t *la, int a, int b ) { (void)la; (void)a; (void)b; return 0; }

/* ---- MB-tree entries, served by oracle/oracle_lookahead.c: orc_la_mbtree_* ---- */
int x264cu_lookahead_frame_set_qp_offset_aq( x264cu_lookahead_t *la, int slot, const float *aq )
{
    if( aq ) orc_la_frame_set_qp_offset_aq( la->slots[slot], aq );
    return 0;
}
int x264cu_lookahead_mbtree_reset( x264cu_lookahead_t *la, int slot ) { orc_la_mbtree_reset( la->slots[slot] ); return 0; }
int x264cu_lookahead_mbtree_swap( x264cu_lookahead_t *la, int a, int b )
{
    uint16_t *t = la->slots[a]->propagate_cost; la->slots[a]->propagate_cost = la->slots[b]->propagate_cost; la->slots[b]->propagate_cost = t;
    return 0;
}
int x264cu_lookahead_mbtree_propagate( x264cu_lookahead_t *la, const int *frames, int p0, int p1, int b, int referenced, float fps_factor )
{
    orc_la_frame_t *fr[300];
    for( int i = p0; i <= p1; i++ ) fr[i] = la->slots[frames[i]];
    orc_la_mbtree_propagate( &la->p, fr, p0, p1, b, referenced, fps_factor );
    return 0;
}
int x264cu_lookahead_mbtree_finish( x264cu_lookahead_t *la, int slot, int fps_factor, int ref0_distance, float strength )
{
    if( !fps_factor ) memcpy( la->slots[slot]->qp_offset, la->slots[slot]->qp_offset_aq, la->slots[slot]->mb_count * sizeof(float) );
    else orc_la_mbtree_finish( la->slots[slot], fps_factor, ref0_distance, strength );
    return 0;
}
int x264cu_lookahead_get_qp_offset( x264cu_lookahead_t *la, int slot, float *out )
{
    memcpy( out, la->slots[slot]->qp_offset, la->slots[slot]->mb_count * sizeof(float) );
    return 0;
}

int x264cu_lookahead_frame_cost_recalculate( x264cu_lookahead_t *la, int slot, int i0, int i1, int b_type, int *score, int32_t *rows )
{
    orc_la_frame_t *fr[2*16 + 8] = { 0 };
    fr[i0] = la->slots[slot];
    *score = orc_la_frame_cost_recalculate( &la->p, fr, 0, i0 + i1, i0, b_type );
    if( rows ) memcpy( rows, la->slots[slot]->row_satds[i0][i1], la->p.mb_height * sizeof(int) );
    return 0;
}
int x264cu_lookahead_get_row_satds( x264cu_lookahead_t *la, int slot, int i0, int i1, int32_t *rows )
{
    memcpy( rows, la->slots[slot]->row_satds[i0][i1], la->p.mb_height * sizeof(int) );
    return 0;
}

/* sharded-stream entries: nothing travels in the CPU harness (every search is computed where it is asked for) */
size_t x264cu_lookahead_search_bytes( x264cu_lookahead_t *la ) { (void)la; return 8; }
void *x264cu_lookahead_exchange_stream( x264cu_lookahead_t *la ) { (void)la; return 0; }
int x264cu_lookahead_export_search( x264cu_lookahead_t *la, int slot, int list, int dist, void *d ) { (void)la; (void)slot; (void)list; (void)dist; (void)d; return 0; }
int x264cu_lookahead_import_search( x264cu_lookahead_t *la, int slot, int list, int dist, const void *d ) { (void)la; (void)slot; (void)list; (void)dist; (void)d; return 0; }
int x264cu_lookahead_import_done( x264cu_lookahead_t *la ) { (void)la; return 0; }

/* I420 pictures: adaptive quantisation by the oracle (oracle/oracle_aq.c) */
int x264cu_lookahead_frame_put_i420( x264cu_lookahead_t *la, int slot, const uint8_t *h_luma, intptr_t luma_stride,
                                     const uint8_t *h_cb, const uint8_t *h_cr, intptr_t chroma_stride, int aq_mode, float aq_strength )
{
    orc_la_frame_delete( la->slots[slot] );
    orc_la_frame_t *f = la->slots[slot] = orc_la_frame_new( &la->p, h_luma, luma_stride );
    orc_adaptive_quant_frame( h_luma, luma_stride, h_cb, h_cr, chroma_stride, la->p.width, la->p.height, aq_mode, aq_strength,
                              f->qp_offset_aq, f->inv_qscale_factor, NULL );
    memcpy( f->qp_offset, f->qp_offset_aq, f->mb_count * sizeof(float) );
    return 0;
}
