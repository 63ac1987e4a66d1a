/*
 * The reference-side binding of libx264_b200: the symbols the reference calls at its lookahead offload seam, where common/opencl.c
 * and encoder/slicetype-cl.c sit today --
 *
 *   x264_opencl_load_library / _close_library     common/opencl.h:796-799      encoder.c:1747, :4575
 *   x264_opencl_lookahead_init / _delete          common/opencl.h:801-804      encoder.c:1798, :4208
 *   x264_opencl_frame_delete                      common/opencl.h:806-807      frame.c:336
 *   x264_opencl_lowres_init / _motionsearch / _finalize_cost / _flush          encoder/slicetype-cl.h:29-38, slicetype.c:878-897
 *   x264_opencl_slicetype_prep / _end             encoder/slicetype-cl.h:39-42 slicetype.c:1531, :1741
 *   x264_opencl_precalculate_frame_cost           encoder/slicetype-cl.h:35    (not called by slicetype.c)
 *
 * -- implemented over the C ABI of include/x264_b200.h.  Compile it INSTEAD of common/opencl.c + encoder/slicetype-cl.c with
 * HAVE_OPENCL=1 (oracle/Makefile.ref, target b200) and link against libx264_b200.so: `--opencl` then runs the lookahead on the B200.
 * Unlike the OpenCL kernels, the results are those of the reference's CPU path bit for bit, so the encoder's output does not change.
 *
 * Division of labour.  The reference keeps everything around the hook: the memo check, do_search, x264_weights_analyse, the
 * slice-type decision, MB-tree and the rate control, all on the host arrays of x264_frame_t.  The hooks (1) keep one lookahead
 * slot per x264_frame_t, uploaded when the frame enters the lookahead, (2) answer a cost request with
 * x264cu_lookahead_frame_cost -- whose own bookkeeping follows the same rules as the reference's sentinels, so both sides agree on
 * what is searched and which temporal-direct vectors exist -- and (3) copy what the host code reads back into x264_frame_t:
 * lowres_mvs / lowres_mv_costs, lowres_costs, i_intra_cost, i_row_satds, i_cost_est(_aq), i_intra_mbs.
 * This file contains no reference code; it only includes the reference's headers for x264_t / x264_frame_t.
 */
#include "common/common.h"
#include "encoder/slicetype-cl.h"
#include "x264_b200.h"

typedef struct
{
    x264_opencl_function_t table;        /* what h->opencl.ocl nominally points to: never called, all NULL */
    x264cu_ctx_t *ctx;
    x264cu_lookahead_t *la;
    int n_slots, mb_count, mb_w, mb_h, do_edges;
    x264_frame_t **owner;                /* slot -> frame */
    int16_t *mvs; int32_t *ints; uint16_t *u16;
    long calls[4];                       /* uploads, cost requests answered, prefetch launches, searches prefetched */
} b200_t;

static b200_t *state_of( x264_t *h ) { return (b200_t *)h->opencl.ocl; }
/* the slot of a frame lives in frame->opencl.luma_hpel (a cl_mem nobody else touches): slot + 1, 0 = none */
static int slot_of( x264_frame_t *f ) { return (int)(intptr_t)f->opencl.luma_hpel - 1; }

static long g_calls[4];
long x264_b200_hooks_calls( int what ) { return what >= 0 && what < 4 ? g_calls[what] : 0; }

x264_opencl_function_t *x264_opencl_load_library( void )
{
    b200_t *s = calloc( 1, sizeof(*s) );
    return s ? &s->table : NULL;
}

void x264_opencl_close_library( x264_opencl_function_t *ocl )
{
    b200_t *s = (b200_t *)ocl;
    if( !s ) return;
    for( int i = 0; i < 4; i++ ) g_calls[i] += s->calls[i];
    if( s->la ) x264cu_lookahead_close( s->la );
    if( s->ctx ) x264cu_close( s->ctx );
    free( s->owner ); free( s->mvs ); free( s->ints ); free( s->u16 );
    free( s );
}

int x264_opencl_lookahead_init( x264_t *h )
{
    b200_t *s = state_of( h );
    if( !s ) return -1;
    if( x264cu_open( &s->ctx, h->param.i_opencl_device ) )
    {
        x264_log( h, X264_LOG_WARNING, "x264_b200: %s\n", x264cu_strerror( NULL ) );
        return -1;
    }
    x264cu_lookahead_params_t p;
    memset( &p, 0, sizeof(p) );
    p.width = h->param.i_width; p.height = h->param.i_height;
    p.subpel_refine = h->param.analyse.i_subpel_refine;
    p.me_method = h->param.analyse.i_me_method;
    p.me_range = h->param.analyse.i_me_range;
    p.mv_range = h->param.analyse.i_mv_range;
    p.bframes = h->param.i_bframe;
    p.bframe_bias = h->param.i_bframe_bias;
    p.weighted_bipred = h->param.analyse.b_weighted_bipred;
    p.aq_mode = h->param.rc.i_aq_mode != 0;
    p.mb_tree = h->param.rc.b_mb_tree;
    p.vbv = h->param.rc.i_vbv_buffer_size != 0;
    p.weighted_pred = h->param.analyse.i_weighted_pred;
    /* every x264_frame_t that can be alive in the input / lookahead pools gets a slot for its lifetime */
    s->n_slots = p.n_slots = h->frames.i_delay + h->param.i_sync_lookahead + 2*h->param.i_bframe + h->param.i_threads + 24;
    s->mb_w = h->mb.i_mb_width; s->mb_h = h->mb.i_mb_height; s->mb_count = h->mb.i_mb_count;
    s->do_edges = h->param.rc.b_mb_tree || h->param.rc.i_vbv_buffer_size || s->mb_w <= 2 || s->mb_h <= 2;      /* slicetype.c:823-833 */
    if( x264cu_lookahead_open( s->ctx, &p, &s->la ) )
    {
        x264_log( h, X264_LOG_WARNING, "x264_b200: %s\n", x264cu_strerror( s->ctx ) );
        x264cu_close( s->ctx ); s->ctx = NULL;
        return -1;
    }
    s->owner = calloc( s->n_slots, sizeof(*s->owner) );
    s->mvs = malloc( (size_t)s->mb_count * 2 * sizeof(int16_t) );
    s->ints = malloc( (size_t)( s->mb_count + s->mb_h ) * sizeof(int32_t) );
    s->u16 = malloc( (size_t)s->mb_count * sizeof(uint16_t) );
    if( !s->owner || !s->mvs || !s->ints || !s->u16 ) return -1;
    x264_log( h, X264_LOG_INFO, "x264_b200: lookahead on CUDA device %d, %d slots\n", h->param.i_opencl_device, s->n_slots );
    return 0;
}

void x264_opencl_lookahead_delete( x264_t *h )
{
    b200_t *s = state_of( h );
    if( s && s->ctx ) x264cu_sync( s->ctx );
}

void x264_opencl_frame_delete( x264_frame_t *frame )
{
    b200_t *s = (b200_t *)frame->opencl.ocl;
    const int slot = slot_of( frame );
    if( s && s->owner && slot >= 0 && slot < s->n_slots && s->owner[slot] == frame )
        s->owner[slot] = NULL;
    frame->opencl.luma_hpel = NULL;
}

static void fail( x264_t *h, const char *what )
{
    b200_t *s = state_of( h );
    x264_log( h, X264_LOG_ERROR, "x264_b200: %s: %s\n", what, x264cu_strerror( s->ctx ) );
    h->param.b_opencl = 0;
    h->opencl.b_fatal_error = 1;                 /* x264_encoder_encode returns -1 from now on, encoder.c:3332-3335 */
}

/* rows / columns the CPU path never visits keep what x264_frame_t held (slicetype.c:823-833): copy visited macroblocks only */
static void store_u16( b200_t *s, uint16_t *dst, const uint16_t *src )
{
    if( s->do_edges )
        memcpy( dst, src, s->mb_count * sizeof(uint16_t) );
    else
        for( int y = 1; y < s->mb_h - 1; y++ )
            memcpy( dst + y*s->mb_w + 1, src + y*s->mb_w + 1, ( s->mb_w - 2 ) * sizeof(uint16_t) );
}

/* upload on first sight of a picture (b_intra_calculated is the OpenCL path's "already on the device" flag, slicetype-cl.c:84-86),
 * and leave its intra results where the host code reads them (slicetype-cl.c:254-280) */
int x264_opencl_lowres_init( x264_t *h, x264_frame_t *fenc, int lambda )
{
    b200_t *s = state_of( h );
    (void)lambda;
    if( fenc->b_intra_calculated )
        return 0;
    int slot = slot_of( fenc );
    if( slot < 0 )
    {
        for( slot = 0; slot < s->n_slots && s->owner[slot]; slot++ )
            ;
        if( slot == s->n_slots ) { x264_log( h, X264_LOG_ERROR, "x264_b200: out of lookahead slots\n" ); h->opencl.b_fatal_error = 1; return -1; }
        s->owner[slot] = fenc;
        fenc->opencl.luma_hpel = (cl_mem)(intptr_t)( slot + 1 );
    }
    fenc->b_intra_calculated = 1;
    if( x264cu_lookahead_frame_put( s->la, slot, fenc->plane[0], fenc->i_stride[0], h->param.rc.i_aq_mode ? fenc->i_inv_qscale_factor : NULL ) )
        { fail( h, "frame_put" ); return -1; }
    s->calls[0]++;
    int score, est = 0, est_aq = 0, intra_mbs = 0;
    if( x264cu_lookahead_frame_cost( s->la, &slot, 0, 0, 0, &score ) ||
        x264cu_lookahead_get_intra( s->la, slot, s->ints ) ||
        x264cu_lookahead_get_cost_est( s->la, slot, 0, 0, &est, &est_aq, &intra_mbs ) )
        { fail( h, "intra cost" ); return -1; }
    for( int i = 0; i < s->mb_count; i++ )
        s->u16[i] = (uint16_t)s->ints[i];
    store_u16( s, fenc->lowres_costs[0][0], s->u16 );
    fenc->i_cost_est[0][0] = est;
    fenc->i_cost_est_aq[0][0] = est_aq;
    if( h->param.rc.i_vbv_buffer_size && x264cu_lookahead_get_row_satds( s->la, slot, 0, 0, fenc->i_row_satds[0][0] ) )
        { fail( h, "row satds" ); return -1; }
    return 0;
}

/* the search itself runs inside the cost request (x264_opencl_finalize_cost below): the backend decides by the same rules as the
 * reference's do_search whether it is due, and has usually run it ahead of time (x264_opencl_slicetype_prep) */
int x264_opencl_motionsearch( x264_t *h, x264_frame_t **frames, int b, int ref, int b_islist1, int lambda, const x264_weight_t *w )
{
    (void)h; (void)frames; (void)b; (void)ref; (void)b_islist1; (void)lambda; (void)w;
    return 0;
}

static int fetch_list( x264_t *h, b200_t *s, x264_frame_t *fenc, int slot, int list, int dist )
{
    if( x264cu_lookahead_get_mvs( s->la, slot, list, dist - 1, s->mvs, s->ints ) ) { fail( h, "get_mvs" ); return -1; }
    memcpy( fenc->lowres_mvs[list][dist-1], s->mvs, (size_t)s->mb_count * 2 * sizeof(int16_t) );
    memcpy( fenc->lowres_mv_costs[list][dist-1], s->ints, (size_t)s->mb_count * sizeof(int) );
    return 0;
}

int x264_opencl_finalize_cost( x264_t *h, int lambda, x264_frame_t **frames, int p0, int p1, int b, int dist_scale_factor )
{
    b200_t *s = state_of( h );
    (void)lambda; (void)dist_scale_factor;
    x264_frame_t *fenc = frames[b];
    int slots[X264_LOOKAHEAD_MAX + 4];
    for( int i = p0; i <= p1; i++ )
        slots[i] = i == p0 || i == p1 || i == b ? slot_of( frames[i] ) : 0;
    if( slots[p0] < 0 || slots[p1] < 0 || slots[b] < 0 ) { x264_log( h, X264_LOG_ERROR, "x264_b200: cost request on a picture that was never uploaded\n" ); h->opencl.b_fatal_error = 1; return -1; }
    int score = 0, est = 0, est_aq = 0, intra_mbs = 0;
    if( x264cu_lookahead_frame_cost( s->la, slots, p0, p1, b, &score ) ) { fail( h, "frame_cost" ); return -1; }
    s->calls[1]++;
    const int d0 = b - p0, d1 = p1 - b, slot = slots[b];
    if( d0 && fetch_list( h, s, fenc, slot, 0, d0 ) ) return -1;
    if( d1 && fetch_list( h, s, fenc, slot, 1, d1 ) ) return -1;
    if( x264cu_lookahead_get_costs( s->la, slot, d0, d1, s->u16 ) ||
        x264cu_lookahead_get_cost_est( s->la, slot, d0, d1, &est, &est_aq, &intra_mbs ) )
        { fail( h, "get_costs" ); return -1; }
    store_u16( s, fenc->lowres_costs[d0][d1], s->u16 );
    fenc->i_cost_est[d0][d1] = est;              /* already scaled for B pictures (slicetype.c:985) */
    fenc->i_cost_est_aq[d0][d1] = est_aq;
    if( b == p1 )
        fenc->i_intra_mbs[d0] = intra_mbs;
    if( h->param.rc.i_vbv_buffer_size && x264cu_lookahead_get_row_satds( s->la, slot, d0, d1, fenc->i_row_satds[d0][d1] ) )
        { fail( h, "row satds" ); return -1; }
    return 0;
}

int x264_opencl_precalculate_frame_cost( x264_t *h, x264_frame_t **frames, int lambda, int p0, int p1, int b )
{
    (void)h; (void)frames; (void)lambda; (void)p0; (void)p1; (void)b;
    return 0;
}

void x264_opencl_flush( x264_t *h ) { (void)h; }       /* every hook has left its results on the host by the time it returns */

/* The analysis window is known here: upload its pictures and launch every lowres search the decision can ask for in ONE batch, so
 * that the GPU is filled (a single search is a thin wavefront).  A search is a pure function of two pictures, so running it
 * early changes no result; with weighted prediction only the pairs whose weight analysis provably ends at its early exit are
 * searched ahead.  The host's sentinels are untouched: the reference still decides what it "has searched". */
void x264_opencl_slicetype_prep( x264_t *h, x264_frame_t **frames, int num_frames, int lambda )
{
    if( !h->param.b_opencl )
        return;
    b200_t *s = state_of( h );
    for( int i = 0; i <= num_frames; i++ )
        if( x264_opencl_lowres_init( h, frames[i], lambda ) < 0 )
            return;
    enum { MAXJ = 1024 };
    int fenc[MAXJ], ref[MAXJ], list[MAXJ], dist[MAXJ], n = 0;
    for( int b = 0; b <= num_frames; b++ )
        for( int d = 1; d <= h->param.i_bframe + 1; d++ )
        {
            if( b - d >= 0 && n < MAXJ && frames[b]->lowres_mvs[0][d-1][0][0] == 0x7FFF )
            {
                int t = x264cu_lookahead_weight_trivial( s->la, slot_of( frames[b] ), slot_of( frames[b-d] ) );
                if( t < 0 ) { fail( h, "weight_trivial" ); return; }
                if( t ) { fenc[n] = slot_of( frames[b] ); ref[n] = slot_of( frames[b-d] ); list[n] = 0; dist[n] = d; n++; }
            }
            if( d <= h->param.i_bframe && b + d <= num_frames && n < MAXJ && frames[b]->lowres_mvs[1][d-1][0][0] == 0x7FFF )
            {
                fenc[n] = slot_of( frames[b] ); ref[n] = slot_of( frames[b+d] ); list[n] = 1; dist[n] = d; n++;
            }
        }
    if( n )
    {
        if( x264cu_lookahead_search_batch( s->la, n, fenc, ref, list, dist ) ) { fail( h, "search_batch" ); return; }
        s->calls[2]++; s->calls[3] += n;
    }
}

void x264_opencl_slicetype_end( x264_t *h ) { (void)h; }
