/* Host control flow of the slice-type decision (plain C, no device code): a from-scratch restatement of
 * x264_slicetype_decide / x264_slicetype_analyse / scenecut / slicetype_path / slicetype_path_cost and of the request
 * sequence of macroblock_tree (encoder/slicetype.c:1091-1184, :1288-1974) plus the synchronous frame queue of
 * encoder/lookahead.c:192-250, issuing x264cu_lookahead_frame_cost wherever the reference issues
 * slicetype_frame_cost.  The order of those requests is part of the result (H3 in SURVEY.md): the memoised B costs
 * depend on whether the later reference's P search had already run (slicetype.c:629-642). */
#include "../../include/x264_b200.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <math.h>
static double st_now( void ) { struct timespec t; clock_gettime( CLOCK_MONOTONIC, &t ); return t.tv_sec + 1e-9 * t.tv_nsec; }

#define LOOKAHEAD_MAX 250                 /* X264_LOOKAHEAD_MAX, common/base.h:140 */
#define BFRAME_MAX X264CU_BFRAME_MAX
#define COST_MAX64 ( 1ULL << 60 )

#define T_AUTO X264CU_TYPE_AUTO
#define T_IDR X264CU_TYPE_IDR
#define T_I X264CU_TYPE_I
#define T_P X264CU_TYPE_P
#define T_BREF X264CU_TYPE_BREF
#define T_B X264CU_TYPE_B
#define T_KEYFRAME X264CU_TYPE_KEYFRAME
#define IS_I( t ) ( ( t ) == T_I || ( t ) == T_IDR )
#define IS_B( t ) ( ( t ) == T_B || ( t ) == T_BREF )
#define AUTO_OR_I( t ) ( ( t ) == T_AUTO || IS_I( t ) )
#define AUTO_OR_B( t ) ( ( t ) == T_AUTO || IS_B( t ) )

#define ST_RUN_AHEAD_MAX 32

typedef struct
{
    int i_frame;          /* display index */
    int slot;             /* lookahead slot holding its lowres planes */
    int i_type, i_forced_type;
    int b_scenecut;       /* frame.c:792 */
    int i_bframes;
    int rc_d0, rc_d1;     /* (b-p0, p1-b) of the cost x264_rc_analyse_slice reads for this picture (slicetype.c:1896-1935, :1985-1996) */
    int n_planned;        /* VBV lookahead (slicetype.c:1225-1286): i_planned_type / i_planned_satd of the coming pictures */
    int planned_type[LOOKAHEAD_MAX + 1], planned_satd[LOOKAHEAD_MAX + 1];
} st_frame_t;

struct x264cu_slicetype
{
    x264cu_ctx_t *ctx;
    x264cu_lookahead_t *la;
    x264cu_slicetype_params_t p;
    int slicetype_length, delay;          /* encoder.c:1602-1612 (one thread, no sync lookahead, cfr) */
    int b_analyse_keyframe;               /* lookahead.c:140 */
    int i_last_keyframe;
    st_frame_t *next[LOOKAHEAD_MAX + ST_RUN_AHEAD_MAX + 8];  /* lookahead->next */
    int n_next;
    st_frame_t *current[BFRAME_MAX + 4];  /* h->frames.current */
    int n_current;
    st_frame_t *last_nonb;
    int i_input;
    int n_slots;
    unsigned char *slot_used;
    long requests;
    int mb_w, mb_h;
    int failed;
    int prefetch;
    st_frame_t *recent[BFRAME_MAX + 2];   /* the last bframes+1 queued pictures, newest first (prefetch partners) */
    int n_recent;
    /* prefetch jobs gathered over a few pictures so that one launch fills the GPU (each search is a thin wavefront) */
    int pj_fenc[256], pj_ref[256], pj_list[256], pj_dist[256], pj_fframe[256], pj_rframe[256], pj_fframe2[256], n_pj, pj_pictures;
    int prefetch_group;                   /* pictures per prefetch launch */
    double t_put, t_batch, t_cost, t_step; long n_cost_calls;   /* X264CU_STATS: where the calling thread's time goes */
    int run_ahead;                        /* extra pictures queued before deciding, like param.i_sync_lookahead (encoder.c:1611) */
    float duration, qcompress;            /* f_duration of every picture (constant frame rate), rc.f_qcompress */
    /* sharded stream (x264cu_slicetype_set_shard): the previous group's jobs, waiting for their exchange */
    int shard_rank, shard_world;
    x264cu_exchange_fn shard_fn;
    void *shard_user;
    int xj_slot[256], xj_list[256], xj_dist[256], xj_owner[256], xj_frame[256], n_xj;
    int next_forced_type;                 /* pic_in->i_type of the next queued picture (x264cu_slicetype_set_next_type) */
    st_frame_t *handed;                   /* the picture the last step returned: kept (slot included) until the next step */
    char best_paths[BFRAME_MAX + 1][LOOKAHEAD_MAX + 1];   /* the trellis of b-adapt 2, slicetype.c:1559-1561 */
    int vbv_lookahead;                    /* h->param.rc.i_vbv_buffer_size && h->param.rc.i_lookahead */
};

int x264cu_slicetype_slot_of( x264cu_slicetype_t *s, int frame );

static inline int num_mbs( const x264cu_slicetype_t *s )          /* NUM_MBS, slicetype.c:794-797 */
{
    return s->mb_w > 2 && s->mb_h > 2 ? ( s->mb_w - 2 ) * ( s->mb_h - 2 ) : s->mb_w * s->mb_h;
}

/* slicetype_frame_cost( h, a, frames, p0, p1, b ) */
static int frame_cost( x264cu_slicetype_t *s, st_frame_t **frames, int p0, int p1, int b )
{
    int slots[LOOKAHEAD_MAX + 4];
    for( int i = p0; i <= p1; i++ ) slots[i] = frames[i]->slot;
    int score = 0;
    s->requests++;
    double t0_ = st_now();
    int rc_ = x264cu_lookahead_frame_cost( s->la, slots, p0, p1, b, &score );
    s->t_cost += st_now() - t0_; s->n_cost_calls++;
    if( rc_ )
    {
        s->failed = 1;
        return 0;
    }
    return score;
}

static void cost_est( x264cu_slicetype_t *s, st_frame_t *f, int i0, int i1, int *ce, int *imb )
{
    int a = 0, aq = 0, m = 0;
    if( x264cu_lookahead_get_cost_est( s->la, f->slot, i0, i1, &a, &aq, &m ) ) s->failed = 1;
    if( ce ) *ce = a;
    if( imb ) *imb = m;
}

/* slicetype_frame_cost_recalculate, slicetype.c:999-1024 */
static int frame_cost_recalculate( x264cu_slicetype_t *s, st_frame_t *f, int i0, int i1, int *rows )
{
    int score = 0;
    if( x264cu_lookahead_frame_cost_recalculate( s->la, f->slot, i0, i1, IS_B( f->i_type ), &score, rows ) ) s->failed = 1;
    return score;
}

/* vbv_frame_cost, slicetype.c:1186-1197 */
static int vbv_frame_cost( x264cu_slicetype_t *s, st_frame_t **frames, int p0, int p1, int b )
{
    int cost = frame_cost( s, frames, p0, p1, b );
    if( s->p.la.aq_mode )
    {
        if( s->p.la.mb_tree )
            return frame_cost_recalculate( s, frames[b], b - p0, p1 - b, NULL );
        int a = 0, aq = 0, m = 0;
        if( x264cu_lookahead_get_cost_est( s->la, frames[b]->slot, b - p0, p1 - b, &a, &aq, &m ) ) s->failed = 1;
        return aq;
    }
    return cost;
}

/* vbv_lookahead, slicetype.c:1225-1286: the planned types and costs of the pictures after the next non-B one, in coded order,
 * left with that picture for the rate control (constant frame rate: the cpb durations of the reference are not produced) */
static void vbv_lookahead( x264cu_slicetype_t *s, st_frame_t **frames, int num_frames, int keyframe )
{
    int last_nonb = 0, cur_nonb = 1, idx = 0;
    while( cur_nonb < num_frames && IS_B( frames[cur_nonb]->i_type ) )
        cur_nonb++;
    st_frame_t *dst = frames[keyframe ? last_nonb : cur_nonb];
    const int skip = keyframe ? -1 : cur_nonb;               /* the cost of the picture holding the plan is not part of it */
    while( cur_nonb < num_frames )
    {
        if( cur_nonb != skip )
        {
            int p0 = IS_I( frames[cur_nonb]->i_type ) ? cur_nonb : last_nonb;
            dst->planned_satd[idx] = vbv_frame_cost( s, frames, p0, cur_nonb, cur_nonb );
            dst->planned_type[idx] = frames[cur_nonb]->i_type;
            idx++;
        }
        for( int i = last_nonb + 1; i < cur_nonb; i++, idx++ )   /* the B pictures, coded after their non-B */
        {
            dst->planned_satd[idx] = vbv_frame_cost( s, frames, last_nonb, cur_nonb, i );
            dst->planned_type[idx] = T_B;
        }
        last_nonb = cur_nonb;
        cur_nonb++;
        while( cur_nonb <= num_frames && IS_B( frames[cur_nonb]->i_type ) )
            cur_nonb++;
    }
    dst->planned_type[idx] = T_AUTO;
    dst->n_planned = idx;
}

/* slicetype.c:1288-1330 */
static unsigned long long path_cost( x264cu_slicetype_t *s, st_frame_t **frames, char *path, unsigned long long threshold )
{
    unsigned long long cost = 0;
    int loc = 1, cur_nonb = 0;
    path--;                                  /* the first path element is really the second frame */
    while( path[loc] )
    {
        int next_nonb = loc;
        while( path[next_nonb] == 'B' ) next_nonb++;
        if( path[next_nonb] == 'P' )
            cost += frame_cost( s, frames, cur_nonb, next_nonb, next_nonb );
        else
            cost += frame_cost( s, frames, next_nonb, next_nonb, next_nonb );
        if( cost > threshold )
            break;
        if( s->p.b_pyramid && next_nonb - cur_nonb > 2 )
        {
            int middle = cur_nonb + ( next_nonb - cur_nonb ) / 2;
            cost += frame_cost( s, frames, cur_nonb, next_nonb, middle );
            for( int next_b = loc; next_b < middle && cost < threshold; next_b++ )
                cost += frame_cost( s, frames, cur_nonb, middle, next_b );
            for( int next_b = middle + 1; next_b < next_nonb && cost < threshold; next_b++ )
                cost += frame_cost( s, frames, middle, next_nonb, next_b );
        }
        else
            for( int next_b = loc; next_b < next_nonb && cost < threshold; next_b++ )
                cost += frame_cost( s, frames, cur_nonb, next_nonb, next_b );
        loc = next_nonb + 1;
        cur_nonb = next_nonb;
    }
    return cost;
}

/* Viterbi step, slicetype.c:1333-1382 */
static void slicetype_path( x264cu_slicetype_t *s, st_frame_t **frames, int length, char ( *best_paths )[LOOKAHEAD_MAX + 1] )
{
    char paths[2][LOOKAHEAD_MAX + 1];
    int num_paths = s->p.la.bframes + 1 < length ? s->p.la.bframes + 1 : length;
    unsigned long long best_cost = COST_MAX64;
    int best_possible = 0, idx = 0;
    for( int path = 0; path < num_paths; path++ )
    {
        int len = length - ( path + 1 );
        memcpy( paths[idx], best_paths[len % ( BFRAME_MAX + 1 )], len );
        memset( paths[idx] + len, 'B', path );
        strcpy( paths[idx] + len + path, "P" );
        int possible = 1;
        for( int i = 1; i <= length; i++ )
        {
            int t = frames[i]->i_type;
            if( t == T_AUTO ) continue;
            if( IS_B( t ) )
                possible = possible && ( i < len || i == length || paths[idx][i-1] == 'B' );
            else
            {
                possible = possible && ( i < len || paths[idx][i-1] != 'B' );
                paths[idx][i-1] = IS_I( t ) ? 'I' : 'P';
            }
        }
        if( possible || !best_possible )
        {
            if( possible && !best_possible )
                best_cost = COST_MAX64;
            unsigned long long cost = path_cost( s, frames, paths[idx], best_cost );
            if( cost < best_cost )
            {
                best_cost = cost;
                best_possible = possible;
                idx ^= 1;
            }
        }
    }
    memcpy( best_paths[length % ( BFRAME_MAX + 1 )], paths[idx ^ 1], length );
}

/* slicetype.c:1384-1428 */
static int scenecut_internal( x264cu_slicetype_t *s, st_frame_t **frames, int p0, int p1, int real_scenecut )
{
    st_frame_t *frame = frames[p1];
    (void)real_scenecut;
    frame_cost( s, frames, p0, p1, p1 );
    int icost, pcost;
    cost_est( s, frame, 0, 0, &icost, NULL );
    cost_est( s, frame, p1 - p0, 0, &pcost, NULL );
    float f_bias;
    int i_gop_size = frame->i_frame - s->i_last_keyframe;
    float f_thresh_max = s->p.scenecut_threshold / 100.0;
    float f_thresh_min = f_thresh_max * 0.25;
    if( s->p.keyint_min == s->p.keyint_max )
        f_thresh_min = f_thresh_max;
    if( i_gop_size <= s->p.keyint_min / 4 || s->p.intra_refresh )
        f_bias = f_thresh_min / 4;
    else if( i_gop_size <= s->p.keyint_min )
        f_bias = f_thresh_min * i_gop_size / s->p.keyint_min;
    else
        f_bias = f_thresh_min + ( f_thresh_max - f_thresh_min ) * ( i_gop_size - s->p.keyint_min )
                 / ( s->p.keyint_max - s->p.keyint_min );
    return pcost >= ( 1.0 - f_bias ) * icost;
}

/* slicetype.c:1430-1468 */
static int scenecut( x264cu_slicetype_t *s, st_frame_t **frames, int p0, int p1, int real_scenecut, int num_frames, int i_max_search )
{
    if( real_scenecut && s->p.la.bframes )
    {
        int origmaxp1 = p0 + 1;
        if( s->p.b_adapt == 2 )
            origmaxp1 += s->p.la.bframes;
        else
            origmaxp1++;
        int maxp1 = origmaxp1 < num_frames ? origmaxp1 : num_frames;
        for( int curp1 = p1; curp1 <= maxp1; curp1++ )
            if( !scenecut_internal( s, frames, p0, curp1, 0 ) )
                for( int i = curp1; i > p0; i-- )
                    frames[i]->b_scenecut = 0;
        for( int curp0 = p0; curp0 <= maxp1; curp0++ )
            if( origmaxp1 > i_max_search || ( curp0 < maxp1 && scenecut_internal( s, frames, curp0, maxp1, 0 ) ) )
                frames[curp0]->b_scenecut = 0;
    }
    if( !frames[p1]->b_scenecut )
        return 0;
    return scenecut_internal( s, frames, p0, p1, real_scenecut );
}

/* CLIP_DURATION, ratecontrol.h: durations outside [0.01 s, 1 s] are clamped */
static float clip_duration( float f ) { return f < 0.01f ? 0.01f : f > 1.00f ? 1.00f : f; }
#define MBTREE_PRECISION 0.5f

static void mbtree_reset( x264cu_slicetype_t *s, st_frame_t *f )
{
    if( x264cu_lookahead_mbtree_reset( s->la, f->slot ) ) s->failed = 1;
}

static void mbtree_finish( x264cu_slicetype_t *s, st_frame_t *f, float average_duration, int ref0_distance );

/* macroblock_tree_propagate, slicetype.c:1050-1089 (constant frame rate: every picture lasts s->duration) */
static void mbtree_propagate( x264cu_slicetype_t *s, st_frame_t **frames, float average_duration, int p0, int p1, int b, int referenced )
{
    int slots[LOOKAHEAD_MAX + 4];
    for( int i = p0; i <= p1; i++ ) slots[i] = frames[i]->slot;
    float fps_factor = clip_duration( s->duration ) / ( clip_duration( average_duration ) * 256.0f ) * MBTREE_PRECISION;
    if( x264cu_lookahead_mbtree_propagate( s->la, slots, p0, p1, b, referenced, fps_factor ) ) s->failed = 1;
    if( s->vbv_lookahead && referenced )                      /* slicetype.c:1087-1088 */
        mbtree_finish( s, frames[b], average_duration, b == p1 ? b - p0 : 0 );
}

/* macroblock_tree_finish, slicetype.c:1029-1048 */
static void mbtree_finish( x264cu_slicetype_t *s, st_frame_t *f, float average_duration, int ref0_distance )
{
    int fps_factor = round( clip_duration( average_duration ) / clip_duration( s->duration ) * 256 / MBTREE_PRECISION );
    float strength = 5.0f * ( 1.0f - s->qcompress );
    if( x264cu_lookahead_mbtree_finish( s->la, f->slot, fps_factor, ref0_distance, strength ) ) s->failed = 1;
}

/* macroblock_tree, slicetype.c:1091-1184: the cost requests (they matter for later memoised costs, slicetype.c:629-642) and the
 * propagation itself on the device.  Every picture carries the same duration here (constant frame rate); the reference reads
 * frame->f_duration, which x264_slicetype_decide only assigns when a picture is decided (slicetype.c:1769) -- pictures still
 * waiting in the lookahead carry whatever their recycled x264_frame_t held, i.e. the same value once the frame pool has been
 * through one cycle, zero (clamped to 0.01 s) before. */
static void macroblock_tree( x264cu_slicetype_t *s, st_frame_t **frames, int num_frames, int b_intra )
{
    int idx = !b_intra;
    int last_nonb, cur_nonb = 1, bframes = 0;
    int i = num_frames;
    float total_duration = 0.0;
    for( int j = 0; j <= num_frames; j++ )
        total_duration += s->duration;
    float average_duration = total_duration / ( num_frames + 1 );
    if( b_intra )
        frame_cost( s, frames, 0, 0, 0 );
    while( i > 0 && IS_B( frames[i]->i_type ) ) i--;
    last_nonb = i;
    if( !s->p.rc_lookahead )
    {
        if( b_intra )
        {   /* i_propagate_cost = 0, f_qp_offset = f_qp_offset_aq */
            mbtree_reset( s, frames[0] );
            if( x264cu_lookahead_mbtree_finish( s->la, frames[0]->slot, 0, 0, 0.0f ) ) s->failed = 1;
            return;
        }
        if( x264cu_lookahead_mbtree_swap( s->la, frames[last_nonb]->slot, frames[0]->slot ) ) s->failed = 1;
        mbtree_reset( s, frames[0] );
    }
    else
    {
        if( last_nonb < idx )
            return;
        mbtree_reset( s, frames[last_nonb] );
    }
    while( i-- > idx )
    {
        cur_nonb = i;
        while( IS_B( frames[cur_nonb]->i_type ) && cur_nonb > 0 ) cur_nonb--;
        if( cur_nonb < idx )
            break;
        frame_cost( s, frames, cur_nonb, last_nonb, last_nonb );
        mbtree_reset( s, frames[cur_nonb] );
        bframes = last_nonb - cur_nonb - 1;
        if( s->p.b_pyramid && bframes > 1 )
        {
            int middle = ( bframes + 1 ) / 2 + cur_nonb;
            frame_cost( s, frames, cur_nonb, last_nonb, middle );
            mbtree_reset( s, frames[middle] );
            while( i > cur_nonb )
            {
                int p0 = i > middle ? middle : cur_nonb;
                int p1 = i < middle ? middle : last_nonb;
                if( i != middle )
                {
                    frame_cost( s, frames, p0, p1, i );
                    mbtree_propagate( s, frames, average_duration, p0, p1, i, 0 );
                }
                i--;
            }
            mbtree_propagate( s, frames, average_duration, cur_nonb, last_nonb, middle, 1 );
        }
        else
            while( i > cur_nonb )
            {
                frame_cost( s, frames, cur_nonb, last_nonb, i );
                mbtree_propagate( s, frames, average_duration, cur_nonb, last_nonb, i, 0 );
                i--;
            }
        mbtree_propagate( s, frames, average_duration, cur_nonb, last_nonb, last_nonb, 1 );
        last_nonb = cur_nonb;
    }
    if( !s->p.rc_lookahead )
    {
        frame_cost( s, frames, 0, last_nonb, last_nonb );
        mbtree_propagate( s, frames, average_duration, 0, last_nonb, last_nonb, 1 );
        if( x264cu_lookahead_mbtree_swap( s->la, frames[last_nonb]->slot, frames[0]->slot ) ) s->failed = 1;
    }
    mbtree_finish( s, frames[last_nonb], average_duration, last_nonb );
    if( s->p.b_pyramid && bframes > 1 && !s->p.la.vbv )
        mbtree_finish( s, frames[last_nonb + ( bframes + 1 ) / 2], average_duration, 0 );
}

/* x264_slicetype_analyse, slicetype.c:1473-1743 */
static void slicetype_analyse( x264cu_slicetype_t *s, int intra_minigop )
{
    st_frame_t *frames[LOOKAHEAD_MAX + 3] = { NULL };
    int num_frames, orig_num_frames, keyint_limit, framecnt;
    int i_max_search = s->n_next < LOOKAHEAD_MAX ? s->n_next : LOOKAHEAD_MAX;
    const int bf = s->p.la.bframes;
    /* b_deterministic */
    if( i_max_search > s->slicetype_length + 1 - intra_minigop )
        i_max_search = s->slicetype_length + 1 - intra_minigop;
    int keyframe = !!intra_minigop;
    if( !s->last_nonb )
        return;
    frames[0] = s->last_nonb;
    for( framecnt = 0; framecnt < i_max_search; framecnt++ )
        frames[framecnt + 1] = s->next[framecnt];
    if( !framecnt )
    {
        if( s->p.la.mb_tree )
            macroblock_tree( s, frames, 0, keyframe );
        return;
    }
    keyint_limit = s->p.keyint_max - frames[0]->i_frame + s->i_last_keyframe - 1;
    orig_num_frames = num_frames = s->p.intra_refresh ? framecnt : framecnt < keyint_limit ? framecnt : keyint_limit;
    if( ( s->p.psy && s->p.la.mb_tree ) || s->vbv_lookahead )
        num_frames = framecnt;
    else if( s->p.open_gop && num_frames < framecnt )
        num_frames++;
    else if( num_frames == 0 )
    {
        frames[1]->i_type = T_I;
        return;
    }
    if( AUTO_OR_I( frames[1]->i_type ) && s->p.scenecut_threshold &&
        scenecut( s, frames, 0, 1, 1, orig_num_frames, i_max_search ) )
    {
        if( frames[1]->i_type == T_AUTO )
            frames[1]->i_type = T_I;
        return;
    }
    /* Replace forced keyframes with I/IDR-frames, slicetype.c:1534-1539 */
    for( int j = 1; j <= num_frames; j++ )
        if( frames[j]->i_type == T_KEYFRAME )
            frames[j]->i_type = s->p.open_gop ? T_I : T_IDR;
    /* Close GOP at IDR-frames */
    for( int j = 2; j <= num_frames; j++ )
        if( frames[j]->i_type == T_IDR && AUTO_OR_B( frames[j-1]->i_type ) )
            frames[j-1]->i_type = T_P;

    int num_analysed_frames = num_frames;
    int reset_start;
    if( bf )
    {
        if( s->p.b_adapt == 2 )
        {
            if( num_frames > 1 )
            {
                char ( *best_paths )[LOOKAHEAD_MAX + 1] = s->best_paths;          /* per object: several streams may run in one process */
                memset( s->best_paths, 0, sizeof( s->best_paths ) );
                strcpy( best_paths[1], "P" );
                int best_path_index = num_frames % ( BFRAME_MAX + 1 );
                for( int j = 2; j <= num_frames; j++ )
                    slicetype_path( s, frames, j, best_paths );
                for( int j = 1; j < num_frames; j++ )
                {
                    if( best_paths[best_path_index][j-1] != 'B' )
                    {
                        if( AUTO_OR_B( frames[j]->i_type ) )
                            frames[j]->i_type = T_P;
                    }
                    else if( frames[j]->i_type == T_AUTO )
                        frames[j]->i_type = T_B;
                }
            }
        }
        else if( s->p.b_adapt == 1 )
        {
            int last_nonb = 0, num_bframes = bf;
            char path[LOOKAHEAD_MAX + 1];
            for( int j = 1; j < num_frames; j++ )
            {
                if( j - 1 > 0 && IS_B( frames[j-1]->i_type ) )
                    num_bframes--;
                else
                {
                    last_nonb = j - 1;
                    num_bframes = bf;
                }
                if( !num_bframes )
                {
                    if( AUTO_OR_B( frames[j]->i_type ) )
                        frames[j]->i_type = T_P;
                    continue;
                }
                if( frames[j]->i_type != T_AUTO )
                    continue;
                if( IS_B( frames[j+1]->i_type ) )
                {
                    frames[j]->i_type = T_P;
                    continue;
                }
                int bframes = j - last_nonb - 1;
                memset( path, 'B', bframes );
                strcpy( path + bframes, "PP" );
                unsigned long long cost_p = path_cost( s, frames + last_nonb, path, COST_MAX64 );
                strcpy( path + bframes, "BP" );
                unsigned long long cost_b = path_cost( s, frames + last_nonb, path, cost_p );
                frames[j]->i_type = cost_b < cost_p ? T_B : T_P;
            }
        }
        else
        {
            int num_bframes = bf;
            for( int j = 1; j < num_frames; j++ )
            {
                if( !num_bframes )
                {
                    if( AUTO_OR_B( frames[j]->i_type ) )
                        frames[j]->i_type = T_P;
                }
                else if( frames[j]->i_type == T_AUTO )
                    frames[j]->i_type = IS_B( frames[j+1]->i_type ) ? T_P : T_B;
                if( IS_B( frames[j]->i_type ) )
                    num_bframes--;
                else
                    num_bframes = bf;
            }
        }
        if( AUTO_OR_B( frames[num_frames]->i_type ) )
            frames[num_frames]->i_type = T_P;

        int num_bframes = 0;
        while( num_bframes < num_frames && IS_B( frames[num_bframes + 1]->i_type ) )
            num_bframes++;
        /* Check scenecut on the first minigop. */
        for( int j = 1; j < num_bframes + 1; j++ )
            if( frames[j]->i_forced_type == T_AUTO && AUTO_OR_I( frames[j+1]->i_forced_type ) &&
                s->p.scenecut_threshold && scenecut( s, frames, j, j + 1, 0, orig_num_frames, i_max_search ) )
            {
                frames[j]->i_type = T_P;
                num_analysed_frames = j;
                break;
            }
        reset_start = keyframe ? 1 : ( num_bframes + 2 < num_analysed_frames + 1 ? num_bframes + 2 : num_analysed_frames + 1 );
    }
    else
    {
        for( int j = 1; j <= num_frames; j++ )
            if( AUTO_OR_B( frames[j]->i_type ) )
                frames[j]->i_type = T_P;
        reset_start = !keyframe + 1;
    }

    if( s->p.la.mb_tree )
        macroblock_tree( s, frames, num_frames < s->p.keyint_max ? num_frames : s->p.keyint_max, keyframe );

    /* Enforce keyframe limit. */
    if( !s->p.intra_refresh )
    {
        int last_keyframe = s->i_last_keyframe, last_possible = 0;
        for( int j = 1; j <= num_frames; j++ )
        {
            st_frame_t *frm = frames[j];
            int keyframe_dist = frm->i_frame - last_keyframe;
            if( AUTO_OR_I( frm->i_forced_type ) )
                if( s->p.open_gop || !IS_B( frames[j-1]->i_forced_type ) )
                    last_possible = j;
            if( keyframe_dist >= s->p.keyint_max )
            {
                if( last_possible != 0 && last_possible != j )
                {
                    j = last_possible;
                    frm = frames[j];
                    keyframe_dist = frm->i_frame - last_keyframe;
                }
                last_possible = 0;
                if( frm->i_type != T_IDR )
                    frm->i_type = s->p.open_gop ? T_I : T_IDR;
            }
            if( frm->i_type == T_I && keyframe_dist >= s->p.keyint_min )
            {
                if( s->p.open_gop )
                    last_keyframe = frm->i_frame;             /* display order (no blu-ray compatibility mode here) */
                else if( frm->i_forced_type != T_I )
                    frm->i_type = T_IDR;
            }
            if( frm->i_type == T_IDR )
            {
                last_keyframe = frm->i_frame;
                if( j > 1 && IS_B( frames[j-1]->i_type ) )
                    frames[j-1]->i_type = T_P;
            }
        }
    }
    if( s->vbv_lookahead )
        vbv_lookahead( s, frames, num_frames, keyframe );
    /* Restore frametypes for all frames that haven't actually been decided yet. */
    for( int j = reset_start; j <= num_frames; j++ )
        frames[j]->i_type = frames[j]->i_forced_type;
}

/* x264_slicetype_decide, slicetype.c:1745-1974 (type decision + the rate-control cost requests) */
static int slicetype_decide( x264cu_slicetype_t *s )
{
    st_frame_t *frames[BFRAME_MAX + 2];
    st_frame_t *frm;
    int bframes, brefs;
    if( !s->n_next )
        return 0;
    if( ( s->p.la.bframes && s->p.b_adapt ) || s->p.scenecut_threshold || s->p.la.mb_tree || s->vbv_lookahead )
        slicetype_analyse( s, 0 );

    for( bframes = 0, brefs = 0;; bframes++ )
    {
        frm = s->next[bframes];
        if( frm->i_type == T_BREF && s->p.b_pyramid < 2 && brefs == s->p.b_pyramid )
            frm->i_type = T_B;
        else if( frm->i_type == T_BREF && s->p.b_pyramid == 2 && brefs && s->p.frame_reference <= ( brefs + 3 ) )
            frm->i_type = T_B;
        if( frm->i_type == T_KEYFRAME )
            frm->i_type = s->p.open_gop ? T_I : T_IDR;
        /* Limit GOP size, slicetype.c:1831-1845 */
        if( ( !s->p.intra_refresh || frm->i_frame == 0 ) && frm->i_frame - s->i_last_keyframe >= s->p.keyint_max )
        {
            const int key = s->p.open_gop && s->i_last_keyframe >= 0 ? T_I : T_IDR;
            if( frm->i_type == T_AUTO || frm->i_type == T_I )
                frm->i_type = key;
            if( frm->i_type != T_IDR && !( s->p.open_gop && frm->i_type == T_I ) )
                frm->i_type = key;
        }
        if( frm->i_type == T_I && frm->i_frame - s->i_last_keyframe >= s->p.keyint_min )
        {
            if( s->p.open_gop )
                s->i_last_keyframe = frm->i_frame;              /* display order */
            else
                frm->i_type = T_IDR;
        }
        if( frm->i_type == T_IDR )
        {
            s->i_last_keyframe = frm->i_frame;
            if( bframes > 0 )
            {
                bframes--;
                s->next[bframes]->i_type = T_P;
            }
        }
        if( bframes == s->p.la.bframes || bframes + 1 >= s->n_next )
        {
            if( frm->i_type == T_AUTO || IS_B( frm->i_type ) )
                frm->i_type = T_P;
        }
        if( frm->i_type == T_BREF )
            brefs++;
        if( frm->i_type == T_AUTO )
            frm->i_type = T_B;
        else if( !IS_B( frm->i_type ) )
            break;
    }
    s->next[bframes]->i_bframes = bframes;
    /* insert a bref into the sequence */
    if( s->p.b_pyramid && bframes > 1 && !brefs )
    {
        s->next[( bframes - 1 ) / 2]->i_type = T_BREF;
        brefs++;
    }
    /* frame costs ahead of time for x264_rc_analyse_slice */
    if( !s->p.rc_cqp )
    {
        int p0, p1, b;
        p1 = b = bframes + 1;
        frames[0] = s->last_nonb;
        memcpy( &frames[1], s->next, ( bframes + 1 ) * sizeof( st_frame_t * ) );
        p0 = IS_I( s->next[bframes]->i_type ) ? bframes + 1 : 0;
        frame_cost( s, frames, p0, p1, b );
        frames[b]->rc_d0 = b - p0; frames[b]->rc_d1 = 0;
        if( ( p0 != p1 || bframes ) && s->p.la.vbv )
        {   /* the intra costs and the B pictures' costs for the row SATDs, slicetype.c:1916-1934 */
            frame_cost( s, frames, b, b, b );
            p0 = 0;
            for( b = 1; b <= bframes; b++ )
            {
                if( frames[b]->i_type == T_B )
                    for( p1 = b; frames[p1]->i_type == T_B; )
                        p1++;
                else
                    p1 = bframes + 1;
                frame_cost( s, frames, p0, p1, b );
                frames[b]->rc_d0 = b - p0; frames[b]->rc_d1 = p1 - b;
                if( frames[b]->i_type == T_BREF )
                    p0 = b;
            }
        }
    }
    /* shift sequence to coded order */
    if( bframes )
    {
        int idx_list[2] = { brefs + 1, 1 };
        for( int i = 0; i < bframes; i++ )
        {
            int idx = idx_list[s->next[i]->i_type == T_BREF]++;
            frames[idx] = s->next[i];
        }
        frames[0] = s->next[bframes];
        memcpy( s->next, frames, ( bframes + 1 ) * sizeof( st_frame_t * ) );
    }
    return bframes;
}

static void release_frame( x264cu_slicetype_t *s, st_frame_t *f )
{
    if( !f ) return;
    s->slot_used[f->slot] = 0;
    for( int k = 0; k < s->n_recent; k++ )
        if( s->recent[k] == f )
        {
            memmove( s->recent + k, s->recent + k + 1, ( s->n_recent - k - 1 ) * sizeof( st_frame_t * ) );
            s->n_recent--;
            break;
        }
    free( f );
}

/* x264_lookahead_get_frames without a lookahead thread, lookahead.c:223-250 */
static void lookahead_get_frames( x264cu_slicetype_t *s )
{
    if( s->n_current || !s->n_next )
        return;
    slicetype_decide( s );
    /* lookahead_update_last_nonb */
    st_frame_t *new_nonb = s->next[0];
    int shift_frames = new_nonb->i_bframes + 1;
    st_frame_t *old = s->last_nonb;
    s->last_nonb = new_nonb;
    /* the frames leave `next` for the encoder; last_nonb stays referenced until it is replaced */
    for( int i = 0; i < shift_frames; i++ )
        s->current[s->n_current++] = s->next[i];
    memmove( s->next, s->next + shift_frames, ( s->n_next - shift_frames ) * sizeof( st_frame_t * ) );
    s->n_next -= shift_frames;
    s->next[s->n_next] = NULL;
    if( old )
    {   /* released unless the encoder side still holds it (it never does: frames are handed out by value) */
        int held = 0;
        for( int i = 0; i < s->n_current; i++ ) held |= s->current[i] == old;
        if( !held ) release_frame( s, old );
    }
    if( s->b_analyse_keyframe && IS_I( s->last_nonb->i_type ) )
        slicetype_analyse( s, shift_frames );
}

int x264cu_slicetype_open( x264cu_ctx_t *ctx, const x264cu_slicetype_params_t *p, x264cu_slicetype_t **out )
{
    if( !ctx || !p || !out ) return -1;
    *out = NULL;
    /* rc.i_lookahead is clipped to X264_LOOKAHEAD_MAX by the reference (encoder.c:1112); the frame lists are sized for it */
    if( p->rc_lookahead < 0 || p->rc_lookahead > LOOKAHEAD_MAX || p->la.bframes < 0 || p->la.bframes > BFRAME_MAX ||
        ( p->la.vbv && p->intra_refresh ) || p->keyint_max < 1 || p->keyint_min < 1 || p->b_adapt < 0 || p->b_adapt > 2 || p->b_pyramid < 0 || p->b_pyramid > 2 )
        return -1;
    x264cu_slicetype_t *s = calloc( 1, sizeof( *s ) );
    if( !s ) return -1;
    s->ctx = ctx;
    s->p = *p;
    /* encoder.c:1602-1609 */
    s->delay = p->b_adapt == 2 ? ( p->la.bframes > 3 ? p->la.bframes : 3 ) * 4 : p->la.bframes;
    if( ( p->la.mb_tree || p->la.vbv ) && p->rc_lookahead > s->delay )
        s->delay = p->rc_lookahead;
    s->slicetype_length = s->delay;
    s->vbv_lookahead = p->la.vbv && p->rc_lookahead;
    s->b_analyse_keyframe = p->la.mb_tree || s->vbv_lookahead;      /* lookahead.c:140 */
    s->i_last_keyframe = -p->keyint_max;
    s->n_slots = s->delay + p->la.bframes + 8 + ST_RUN_AHEAD_MAX;
    s->slot_used = calloc( s->n_slots, 1 );
    s->p.la.n_slots = s->n_slots;
    s->mb_w = ( p->la.width + 15 ) >> 4;
    s->mb_h = ( p->la.height + 15 ) >> 4;
    {   /* slicetype.c:1769-1771 with i_duration = 2 (progressive), vui.i_num_units_in_tick = fps_den, i_time_scale = 2*fps_num */
        int num = p->fps_num > 0 ? p->fps_num : 25, den = p->fps_den > 0 ? p->fps_den : 1;
        s->duration = (double)2 * den / ( 2.0 * num );
        s->qcompress = p->qcompress > 0 ? p->qcompress : 0.6f;
    }
    s->prefetch = 1;
    /* measured at 4K (B200): 8/4 -> 1000 pictures/s, 16/8 -> 1260, 24/12 -> 1410, 32/16 -> 1490: a launch needs several
     * dozen independent wavefronts to fill the 148 SMs */
    s->prefetch_group = s->delay >= 12 ? 12 : 1;
    s->run_ahead = s->delay >= 12 ? 24 : 0;
    {   /* tuning hooks (bench experiments): X264CU_RUN_AHEAD=<0..16>, X264CU_PREFETCH_GROUP=<1..8> */
        const char *e = getenv( "X264CU_RUN_AHEAD" );
        if( e && atoi( e ) >= 0 && atoi( e ) <= ST_RUN_AHEAD_MAX ) s->run_ahead = atoi( e );
        e = getenv( "X264CU_PREFETCH_GROUP" );
        if( e && atoi( e ) >= 1 && atoi( e ) <= 16 ) s->prefetch_group = atoi( e );
    }
    if( !s->slot_used || x264cu_lookahead_open( ctx, &s->p.la, &s->la ) )
    {
        free( s->slot_used );
        free( s );
        return -1;
    }
    *out = s;
    return 0;
}

void x264cu_slicetype_close( x264cu_slicetype_t *s )
{
    if( !s ) return;
    for( int i = 0; i < s->n_next; i++ ) free( s->next[i] );
    for( int i = 0; i < s->n_current; i++ ) if( s->current[i] != s->last_nonb ) free( s->current[i] );
    if( s->handed && s->handed != s->last_nonb ) free( s->handed );
    free( s->last_nonb );
    if( getenv( "X264CU_STATS" ) )
        fprintf( stderr, "x264cu slicetype host time: step %.1f ms = frame_put %.1f + search_batch %.1f + frame_cost %.1f (%ld calls) + logic %.1f\n",
                 s->t_step * 1e3, s->t_put * 1e3, s->t_batch * 1e3, s->t_cost * 1e3, s->n_cost_calls,
                 ( s->t_step - s->t_put - s->t_batch - s->t_cost ) * 1e3 );
    x264cu_lookahead_close( s->la );
    free( s->slot_used );
    free( s );
}

/* Sharded stream: all-gather the results of the previous group's searches (launched one group ago: nobody waits) and install
 * the ones searched on other GPUs.  Every rank holds the same job list, so the layout of the exchange is known everywhere:
 * rank r's k-th job sits at r * bytes_per_rank + k * search_bytes. */
static int shard_exchange( x264cu_slicetype_t *s )
{
    if( !s->n_xj ) return 0;
    int cnt[64] = { 0 }, maxc = 0;
    for( int i = 0; i < s->n_xj; i++ ) cnt[s->xj_owner[i]]++;
    for( int r = 0; r < s->shard_world; r++ ) if( cnt[r] > maxc ) maxc = cnt[r];
    const size_t rec = x264cu_lookahead_search_bytes( s->la ), per_rank = (size_t)maxc * rec;
    void *d_send = NULL, *d_recv = NULL;
    if( s->shard_fn( s->shard_user, 0, per_rank, &d_send, &d_recv, x264cu_lookahead_exchange_stream( s->la ) ) || !d_send || !d_recv )
        return -1;
    int k = 0;
    for( int i = 0; i < s->n_xj; i++ )
        if( s->xj_owner[i] == s->shard_rank )
        {   /* a picture that has left its slot since (end of stream) has nothing to send: the block stays as it is */
            if( x264cu_slicetype_slot_of( s, s->xj_frame[i] ) == s->xj_slot[i] &&
                x264cu_lookahead_export_search( s->la, s->xj_slot[i], s->xj_list[i], s->xj_dist[i], (char *)d_send + (size_t)k * rec ) )
                return -1;
            k++;
        }
    if( s->shard_fn( s->shard_user, 1, per_rank, &d_send, &d_recv, x264cu_lookahead_exchange_stream( s->la ) ) )
        return -1;
    int pos[64] = { 0 };
    for( int i = 0; i < s->n_xj; i++ )
    {
        int r = s->xj_owner[i], kk = pos[r]++;
        if( r == s->shard_rank || x264cu_slicetype_slot_of( s, s->xj_frame[i] ) != s->xj_slot[i] )
            continue;
        if( x264cu_lookahead_import_search( s->la, s->xj_slot[i], s->xj_list[i], s->xj_dist[i], (char *)d_recv + (size_t)r * per_rank + (size_t)kk * rec ) )
            return -1;
    }
    s->n_xj = 0;
    return x264cu_lookahead_import_done( s->la );
}

/* launch the gathered searches; jobs whose pictures have left their slots in the meantime are dropped */
static int flush_prefetch( x264cu_slicetype_t *s )
{
    int n = 0;
    for( int i = 0; i < s->n_pj; i++ )
    {
        if( x264cu_slicetype_slot_of( s, s->pj_fframe[i] ) != s->pj_fenc[i] || x264cu_slicetype_slot_of( s, s->pj_rframe[i] ) != s->pj_ref[i] )
            continue;
        if( s->p.la.weighted_pred && s->pj_list[i] == 0 )
        {   /* a list-0 search is weighted if it is first requested as a P cost and the analysis picks a weight
             * (slicetype.c:857-864): it is a pure function of the two pictures only where the analysis cannot pick one */
            int t = x264cu_lookahead_weight_trivial( s->la, s->pj_fenc[i], s->pj_ref[i] );
            if( t < 0 ) return -1;
            if( !t ) continue;
        }
        s->pj_fenc[n] = s->pj_fenc[i]; s->pj_ref[n] = s->pj_ref[i]; s->pj_list[n] = s->pj_list[i]; s->pj_dist[n] = s->pj_dist[i];
        s->pj_fframe2[n] = s->pj_fframe[i];
        n++;
    }
    s->n_pj = 0;
    s->pj_pictures = 0;
    double t0_ = st_now();
    int rc_ = 0;
    if( s->shard_world > 1 )
    {   /* first the exchange of the group launched one flush ago, then this group's own share */
        if( shard_exchange( s ) ) return -1;
        int m = 0;
        for( int i = 0; i < n; i++ )
        {
            int owner = s->pj_fframe2[i] % s->shard_world;
            s->xj_slot[i] = s->pj_fenc[i]; s->xj_list[i] = s->pj_list[i]; s->xj_dist[i] = s->pj_dist[i];
            s->xj_owner[i] = owner; s->xj_frame[i] = s->pj_fframe2[i];
            if( owner == s->shard_rank )
            {
                s->pj_fenc[m] = s->pj_fenc[i]; s->pj_ref[m] = s->pj_ref[i]; s->pj_list[m] = s->pj_list[i]; s->pj_dist[m] = s->pj_dist[i];
                m++;
            }
        }
        s->n_xj = n;
        rc_ = m ? x264cu_lookahead_search_batch( s->la, m, s->pj_fenc, s->pj_ref, s->pj_list, s->pj_dist ) : 0;
    }
    else
        rc_ = n ? x264cu_lookahead_search_batch( s->la, n, s->pj_fenc, s->pj_ref, s->pj_list, s->pj_dist ) : 0;
    s->t_batch += st_now() - t0_;
    return rc_ ? -1 : 0;
}

static int step_common( x264cu_slicetype_t *s, const uint8_t *luma, int on_device, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                        int *out_frame, int *out_type, const uint8_t *cb, const uint8_t *cr, intptr_t chroma_stride )
{
    if( !s || !out_frame || !out_type ) return -1;
    *out_frame = -1; *out_type = T_AUTO;
    if( s->handed )
    {   /* the picture returned by the previous call leaves now (its slot with it), unless it is still the last non-B one */
        if( s->handed != s->last_nonb ) release_frame( s, s->handed );
        s->handed = NULL;
    }
    if( luma )
    {
        int slot = -1;
        for( int i = 0; i < s->n_slots; i++ )
            if( !s->slot_used[i] ) { slot = i; break; }
        if( slot < 0 || s->n_next >= LOOKAHEAD_MAX + ST_RUN_AHEAD_MAX + 4 ) return -1;
        double t0_ = st_now();
        int rc_ = on_device ? x264cu_lookahead_frame_put_device( s->la, slot, luma, luma_stride, h_inv_qscale )
                : cb        ? x264cu_lookahead_frame_put_i420( s->la, slot, luma, luma_stride, cb, cr, chroma_stride, s->p.la.aq_mode,
                                                               s->p.aq_strength > 0 ? s->p.aq_strength : 1.0f )
                            : x264cu_lookahead_frame_put( s->la, slot, luma, luma_stride, h_inv_qscale );
        s->t_put += st_now() - t0_;
        if( rc_ ) return -1;
        st_frame_t *f = calloc( 1, sizeof( *f ) );
        if( !f ) return -1;
        f->i_frame = s->i_input++;
        f->slot = slot;
        f->i_type = f->i_forced_type = s->next_forced_type;      /* x264_frame_copy_picture, frame.c:370-376 */
        s->next_forced_type = T_AUTO;
        f->b_scenecut = 1;
        s->slot_used[slot] = 1;
        s->next[s->n_next++] = f;
        s->next[s->n_next] = NULL;
        if( s->prefetch )
        {   /* every (picture, earlier picture) pair the decision could ask about: list 0 at distance d <= bframes+1 from the
             * new picture, list 1 at distance d <= bframes towards it */
            for( int k = 0; k < s->n_recent; k++ )
            {
                st_frame_t *o = s->recent[k];
                int d = f->i_frame - o->i_frame;
                if( !s->slot_used[o->slot] || d < 1 || d > s->p.la.bframes + 1 || s->n_pj + 2 > 256 ) continue;
                int n = s->n_pj;
                {   /* with weighted prediction flush_prefetch drops the pairs whose weight analysis is not trivial */
                    s->pj_fenc[n] = f->slot; s->pj_ref[n] = o->slot; s->pj_list[n] = 0; s->pj_dist[n] = d;
                    s->pj_fframe[n] = f->i_frame; s->pj_rframe[n] = o->i_frame; n++;
                }
                if( d <= s->p.la.bframes )
                {
                    s->pj_fenc[n] = o->slot; s->pj_ref[n] = f->slot; s->pj_list[n] = 1; s->pj_dist[n] = d;
                    s->pj_fframe[n] = o->i_frame; s->pj_rframe[n] = f->i_frame; n++;
                }
                s->n_pj = n;
            }
            if( ++s->pj_pictures >= s->prefetch_group && flush_prefetch( s ) ) return -1;
        }
        /* remember by value: the st_frame_t may be freed once the picture is encoded, its slot id stays meaningful only
         * while slot_used says so AND it still holds this picture -- tracked through the frame number */
        {
            int keep = s->p.la.bframes + 1;
            if( s->n_recent < keep ) s->n_recent++;
            for( int k = s->n_recent - 1; k > 0; k-- ) s->recent[k] = s->recent[k-1];
            s->recent[0] = f;
        }
        if( s->i_input <= s->delay + s->run_ahead )   /* encoder.c:3428: nothing to encode yet (i_delay includes the sync-lookahead pictures) */
            return 0;
    }
    if( !luma && s->n_pj && flush_prefetch( s ) ) return -1;
    if( !luma && s->shard_world > 1 && s->n_xj && shard_exchange( s ) ) return -1;
    lookahead_get_frames( s );
    if( s->failed ) return -1;
    if( !s->n_current )
        return 0;
    st_frame_t *f = s->current[0];
    memmove( s->current, s->current + 1, ( s->n_current - 1 ) * sizeof( st_frame_t * ) );
    s->n_current--;
    *out_frame = f->i_frame;
    *out_type = f->i_type;
    s->handed = f;
    return 0;
}

int x264cu_slicetype_step( x264cu_slicetype_t *s, const uint8_t *h_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                           int *out_frame, int *out_type )
{
    double t0 = st_now();
    int rc = step_common( s, h_luma, 0, luma_stride, h_inv_qscale, out_frame, out_type, NULL, NULL, 0 );
    if( s ) s->t_step += st_now() - t0;
    return rc;
}

int x264cu_slicetype_step_i420( x264cu_slicetype_t *s, const uint8_t *h_luma, intptr_t luma_stride, const uint8_t *h_cb, const uint8_t *h_cr,
                                intptr_t chroma_stride, int *out_frame, int *out_type )
{
    if( h_luma && ( !h_cb || !h_cr ) ) return -1;
    double t0 = st_now();
    int rc = step_common( s, h_luma, 0, luma_stride, NULL, out_frame, out_type, h_luma ? h_cb : NULL, h_cr, chroma_stride );
    if( s ) s->t_step += st_now() - t0;
    return rc;
}

int x264cu_slicetype_step_device( x264cu_slicetype_t *s, const uint8_t *d_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                                  int *out_frame, int *out_type )
{
    double t0 = st_now();
    int rc = step_common( s, d_luma, 1, luma_stride, h_inv_qscale, out_frame, out_type, NULL, NULL, 0 );
    if( s ) s->t_step += st_now() - t0;
    return rc;
}

void x264cu_slicetype_set_prefetch( x264cu_slicetype_t *s, int prefetch ) { if( s ) s->prefetch = !!prefetch; }

void x264cu_slicetype_set_run_ahead( x264cu_slicetype_t *s, int pictures )
{
    if( s && !s->i_input && pictures >= 0 && pictures <= ST_RUN_AHEAD_MAX ) s->run_ahead = pictures;
}

int x264cu_slicetype_set_shard( x264cu_slicetype_t *s, int rank, int world, x264cu_exchange_fn fn, void *user )
{
    if( !s || s->i_input || world < 1 || world > 64 || rank < 0 || rank >= world || ( world > 1 && !fn ) ) return -1;
    s->shard_rank = rank; s->shard_world = world; s->shard_fn = fn; s->shard_user = user;
    return 0;
}

int x264cu_slicetype_set_next_type( x264cu_slicetype_t *s, int type )
{
    if( !s || type < T_AUTO || type > T_KEYFRAME ) return -1;
    s->next_forced_type = type;
    return 0;
}

void x264cu_slicetype_set_async_upload( x264cu_slicetype_t *s, int on ) { if( s ) x264cu_lookahead_set_async_upload( s->la, on ); }

x264cu_lookahead_t *x264cu_slicetype_lookahead( x264cu_slicetype_t *s ) { return s ? s->la : NULL; }

int x264cu_slicetype_slot_of( x264cu_slicetype_t *s, int frame )
{
    if( !s ) return -1;
    if( s->last_nonb && s->last_nonb->i_frame == frame ) return s->last_nonb->slot;
    if( s->handed && s->handed->i_frame == frame ) return s->handed->slot;
    for( int i = 0; i < s->n_next; i++ ) if( s->next[i]->i_frame == frame ) return s->next[i]->slot;
    for( int i = 0; i < s->n_current; i++ ) if( s->current[i]->i_frame == frame ) return s->current[i]->slot;
    return -1;
}

int x264cu_slicetype_get_qp_offset( x264cu_slicetype_t *s, int frame, float *h_qp_offset )
{
    if( !s || !h_qp_offset ) return -1;
    int slot = x264cu_slicetype_slot_of( s, frame );
    if( slot < 0 ) return -1;
    return x264cu_lookahead_get_qp_offset( s->la, slot, h_qp_offset );
}

static st_frame_t *find_held( x264cu_slicetype_t *s, int frame )
{
    if( s->handed && s->handed->i_frame == frame ) return s->handed;
    if( s->last_nonb && s->last_nonb->i_frame == frame ) return s->last_nonb;
    for( int i = 0; i < s->n_current; i++ ) if( s->current[i]->i_frame == frame ) return s->current[i];
    return NULL;
}

/* x264_rc_analyse_slice, slicetype.c:1976-2030, for a picture the last x264cu_slicetype_step returned */
int x264cu_slicetype_rc_analyse_slice( x264cu_slicetype_t *s, int frame, int *cost_out, int *h_row_satd, int *h_row_satd_intra )
{
    if( !s || !cost_out || s->p.rc_cqp ) return -1;
    st_frame_t *f = find_held( s, frame );
    if( !f ) return -1;
    if( IS_B( f->i_type ) && !s->p.la.vbv ) return -1;      /* their costs are requested only for the VBV row SATDs (slicetype.c:1916) */
    const int i0 = IS_I( f->i_type ) ? 0 : f->rc_d0, i1 = IS_I( f->i_type ) ? 0 : f->rc_d1;
    int cost = 0, aq = 0, m = 0;
    if( x264cu_lookahead_get_cost_est( s->la, f->slot, i0, i1, &cost, &aq, &m ) || cost < 0 ) return -1;
    if( s->p.la.mb_tree )
    {
        cost = frame_cost_recalculate( s, f, i0, i1, h_row_satd );
        if( !IS_I( f->i_type ) && s->p.la.vbv )
        {   /* slicetype_frame_cost_recalculate( h, frames, b, b, b ): the intra rows with the same offsets */
            int t = 0;
            if( x264cu_lookahead_frame_cost_recalculate( s->la, f->slot, 0, 0, IS_B( f->i_type ), &t, NULL ) ) s->failed = 1;
        }
    }
    else
    {
        if( s->p.la.aq_mode ) cost = aq;
        if( h_row_satd && x264cu_lookahead_get_row_satds( s->la, f->slot, i0, i1, h_row_satd ) ) s->failed = 1;
    }
    if( h_row_satd_intra && !IS_I( f->i_type ) && x264cu_lookahead_get_row_satds( s->la, f->slot, 0, 0, h_row_satd_intra ) ) s->failed = 1;
    *cost_out = cost;
    return s->failed ? -1 : 0;
}

/* i_planned_type / i_planned_satd of a non-B picture the last step returned (VBV lookahead); returns the number of entries */
int x264cu_slicetype_get_planned( x264cu_slicetype_t *s, int frame, int *h_type, int *h_satd, int max_entries )
{
    if( !s || !s->vbv_lookahead ) return -1;
    st_frame_t *f = find_held( s, frame );
    if( !f ) return -1;
    int n = f->n_planned < max_entries ? f->n_planned : max_entries;
    for( int i = 0; i < n; i++ )
    {
        if( h_type ) h_type[i] = f->planned_type[i];
        if( h_satd ) h_satd[i] = f->planned_satd[i];
    }
    return n;
}

long x264cu_slicetype_cost_requests( x264cu_slicetype_t *s ) { return s ? s->requests : 0; }
