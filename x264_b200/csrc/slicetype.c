/* Slice-type decision on top of the GPU lookahead (plain C, no device code).
 *
 * What it has to equal: the frame types -- and, for the rate control, the costs and MB-tree offsets -- that the reference's
 * x264_slicetype_decide / x264_slicetype_analyse / macroblock_tree produce (encoder/slicetype.c:1029-1184, :1225-1974) behind
 * its synchronous frame queue (encoder/lookahead.c:192-250).  The ORDER in which (p0, p1, b) frame costs are first asked for is
 * part of that result: a B cost uses the later reference's P vectors only if those had been asked for by then
 * (slicetype.c:629-642), and the weight analysis runs only on a first P-type request (:857-864).
 *
 * How it is built here: the decision is split into PLANNERS that turn a window of picture types into ordered request lists
 * (which triples, in which order, under which stopping rule) and small interpreters that walk those lists against a memo of the
 * scores already known.  The reference's request order is thus kept as data.  Around it: the picture queue, the prefetcher that
 * launches every search a window can ask for in large groups, and the sharded mode that splits those groups between GPUs.
 */
#include "../../include/x264_b200.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

enum
{
    WIN_MAX = 250,                       /* X264_LOOKAHEAD_MAX, common/base.h:140 */
    GAP_MAX = X264CU_BFRAME_MAX,         /* most B pictures between two anchors */
    AHEAD_MAX = 64,                      /* run-ahead pictures (set_run_ahead) */
    QUEUE_MAX = WIN_MAX + AHEAD_MAX + 8,
    JOBS_MAX = 1024                      /* searches per prefetch launch group: 12 pictures x (2 bframes + 1) at bframes 16 = 396 */
};
#define SCORE_INF ( 1ULL << 60 )         /* COST_MAX64, encoder/me.h:31 */

enum { T_AUTO = X264CU_TYPE_AUTO, T_IDR = X264CU_TYPE_IDR, T_I = X264CU_TYPE_I, T_P = X264CU_TYPE_P, T_BREF = X264CU_TYPE_BREF,
       T_B = X264CU_TYPE_B, T_KEY = X264CU_TYPE_KEYFRAME };
static inline int intra_type( int t ) { return t == T_I || t == T_IDR; }
static inline int bi_type( int t ) { return t == T_B || t == T_BREF; }
static inline int open_or_intra( int t ) { return t == T_AUTO || intra_type( t ); }
static inline int open_or_bi( int t ) { return t == T_AUTO || bi_type( t ); }

typedef struct
{
    int number;                          /* display index */
    int slot;                            /* lookahead slot holding its lowres planes */
    int type, asked;                     /* decided so far / what the caller forced (pic_in->i_type) */
    int cut_candidate;                   /* x264_frame_t.b_scenecut (frame.c:792): may still turn out to open a scene */
    int trailing_b;                      /* i_bframes of a non-B picture: the B pictures coded after it */
    int rc0, rc1;                        /* (b-p0, p1-b) of the cost the rate control reads for this picture (slicetype.c:1896-1935) */
    int n_plan, plan_type[WIN_MAX + 1], plan_satd[WIN_MAX + 1];   /* VBV lookahead: i_planned_type / i_planned_satd */
    int memo[GAP_MAX + 2][GAP_MAX + 2];  /* scores already returned for this picture as b, by (b-p0, p1-b); -1 = not yet */
} picture_t;

struct x264cu_slicetype
{
    x264cu_ctx_t *ctx;
    x264cu_lookahead_t *la;
    x264cu_slicetype_params_t p;
    int horizon;                         /* i_slicetype_length = frames.i_delay (encoder.c:1602-1612, one thread, cfr) */
    int keyframe_pass;                   /* lookahead->b_analyse_keyframe, lookahead.c:140 */
    int vbv_plan;                        /* rc.i_vbv_buffer_size && rc.i_lookahead */
    int last_key;                        /* lookahead->i_last_keyframe */
    picture_t *queue[QUEUE_MAX];         /* lookahead->next */
    int n_queue;
    picture_t *ready[GAP_MAX + 4];       /* h->frames.current: decided, coded order */
    int n_ready;
    picture_t *anchor;                   /* lookahead->last_nonb */
    picture_t *out;                      /* the picture the last step returned: kept, slot included, until the next step */
    int fed;                             /* pictures queued so far */
    int n_slots;
    unsigned char *slot_busy;
    long requests;
    int far_list1;                       /* the largest p1-b any request named for a B picture: what the list-1 prefetch has to cover */
    int mb_w, mb_h;
    int broken;                          /* a lookahead call failed */
    float tick, qcompress, aq_strength;  /* picture duration (constant frame rate), rc.f_qcompress, rc.f_aq_strength */
    int next_asked;                      /* type forced on the next queued picture */
    /* prefetch: every search the decision could ask for is launched ahead of time, in groups */
    int prefetch, group, in_group, run_ahead;
    int speculate;                       /* cost requests computed with the searches (x264cu_lookahead_finalize_batch): 0 / 1, -1 = when it pays (launch_group) */
    picture_t *recent[GAP_MAX + 2];      /* the last bframes+1 queued pictures, newest first */
    int n_recent;
    struct { int fenc_slot, ref_slot, list, dist, fenc_no, ref_no; } job[JOBS_MAX];
    int n_job;
    /* sharded stream: the previous group's searches, waiting for their exchange */
    int rank, world;
    x264cu_exchange_fn exchange;
    void *exchange_user;
    struct { int slot, list, dist, owner, number; } sent[JOBS_MAX];
    int n_sent;
    unsigned char trellis[GAP_MAX + 1][WIN_MAX + 1];   /* b-adapt 2: best sequence per length, modulo GAP_MAX+1 (slicetype.c:1559) */
};

int x264cu_slicetype_slot_of( x264cu_slicetype_t *s, int frame );

/* ================================================================================================================
 * scores
 * ============================================================================================================== */

/* slicetype_frame_cost( h, a, frames, p0, p1, b ) through the memo; w[] is the analysis window (w[0] = the last anchor) */
static int score3( x264cu_slicetype_t *s, picture_t **w, int p0, int p1, int b )
{
    int *known = &w[b]->memo[b - p0][p1 - b];
    s->requests++;
    if( b < p1 && b > p0 && p1 - b > s->far_list1 )
        s->far_list1 = p1 - b;
    if( *known >= 0 )
        return *known;
    int slots[WIN_MAX + 4], v = 0;
    for( int i = p0; i <= p1; i++ )
        slots[i] = w[i]->slot;
    if( x264cu_lookahead_frame_cost( s->la, slots, p0, p1, b, &v ) )
    {
        s->broken = 1;
        return 0;
    }
    return *known = v;
}

static void estimates( x264cu_slicetype_t *s, picture_t *f, int d0, int d1, int *plain, int *aq )
{
    int a = 0, q = 0, m = 0;
    if( x264cu_lookahead_get_cost_est( s->la, f->slot, d0, d1, &a, &q, &m ) )
        s->broken = 1;
    if( plain ) *plain = a;
    if( aq ) *aq = q;
}

/* slicetype_frame_cost_recalculate (slicetype.c:999-1024) */
static int rescored( x264cu_slicetype_t *s, picture_t *f, int d0, int d1, int *rows )
{
    int v = 0;
    if( x264cu_lookahead_frame_cost_recalculate( s->la, f->slot, d0, d1, bi_type( f->type ), &v, rows ) )
        s->broken = 1;
    return v;
}

/* ================================================================================================================
 * planner 1: what one mini-GOP asks for.  A mini-GOP is (left anchor, right anchor) with B pictures in between; the reference
 * asks for the right anchor's cost first, then the B pictures' -- through the middle one when B pyramids are on -- and stops
 * early against a bound in three different ways (slicetype.c:1288-1330).
 * ============================================================================================================== */
enum { ASK_ANCHOR,          /* always asked; the walk ends when the sum then EXCEEDS the bound */
       ASK_ALWAYS,          /* always asked */
       ASK_BELOW };         /* asked only while the sum is still BELOW the bound */
typedef struct { short p0, p1, b; char rule; } ask_t;

static int asks_of_minigop( ask_t *out, int left, int right, int intra_anchor, int pyramid )
{
    int n = 0;
    out[n++] = (ask_t){ intra_anchor ? right : left, right, right, ASK_ANCHOR };
    if( pyramid && right - left > 2 )
    {
        const int mid = left + ( right - left ) / 2;
        out[n++] = (ask_t){ left, right, mid, ASK_ALWAYS };
        for( int b = left + 1; b < mid; b++ )
            out[n++] = (ask_t){ left, mid, b, ASK_BELOW };
        for( int b = mid + 1; b < right; b++ )
            out[n++] = (ask_t){ mid, right, b, ASK_BELOW };
    }
    else
        for( int b = left + 1; b < right; b++ )
            out[n++] = (ask_t){ left, right, b, ASK_BELOW };
    return n;
}

enum { K_B = 0, K_P = 1, K_I = 2 };      /* the letters of the reference's path strings */

/* slicetype_path_cost: kind[i] describes picture i+1 of the window, n of them, the last one an anchor */
static unsigned long long score_sequence( x264cu_slicetype_t *s, picture_t **w, const unsigned char *kind, int n, unsigned long long bound )
{
    unsigned long long sum = 0;
    ask_t asks[GAP_MAX + 2];
    for( int left = 0, right; left < n; left = right )
    {
        for( right = left + 1; right < n && kind[right - 1] == K_B; right++ )
            ;
        const int n_asks = asks_of_minigop( asks, left, right, kind[right - 1] == K_I, s->p.b_pyramid );
        for( int i = 0; i < n_asks; i++ )
        {
            const ask_t *a = &asks[i];
            if( a->rule == ASK_BELOW && sum >= bound )
                continue;
            sum += score3( s, w, a->p0, a->p1, a->b );
            if( a->rule == ASK_ANCHOR && sum > bound )
                return sum;
        }
    }
    return sum;
}

/* ================================================================================================================
 * b-adapt 2: dynamic programme over window lengths (slicetype.c:1333-1382, :1559-1579).  best(n) = the cheapest type sequence
 * of the first n pictures that ends in an anchor = best(n-k-1) followed by k B pictures and a P, over k <= bframes.
 * ============================================================================================================== */
static void trellis_extend( x264cu_slicetype_t *s, picture_t **w, int n )
{
    unsigned char cand[2][WIN_MAX + 1];
    int keep = 0;                                   /* cand[keep ^ 1] holds the best so far */
    unsigned long long best = SCORE_INF;
    int best_feasible = 0;
    const int most_b = s->p.la.bframes + 1 < n ? s->p.la.bframes + 1 : n;
    for( int k = 0; k < most_b; k++ )
    {
        unsigned char *c = cand[keep];
        const int head = n - k - 1;
        memcpy( c, s->trellis[head % ( GAP_MAX + 1 )], head );
        memset( c + head, K_B, k );
        c[n - 1] = K_P;
        /* types the caller has fixed: a sequence contradicting them is only a fallback */
        int feasible = 1;
        for( int i = 1; i <= n; i++ )
        {
            const int t = w[i]->type;
            if( t == T_AUTO )
                continue;
            if( bi_type( t ) )
                feasible &= i < head || i == n || c[i - 1] == K_B;
            else
            {
                feasible &= i < head || c[i - 1] != K_B;
                c[i - 1] = intra_type( t ) ? K_I : K_P;
            }
        }
        if( !feasible && best_feasible )
            continue;
        if( feasible && !best_feasible )
            best = SCORE_INF;                        /* the first feasible sequence beats any infeasible one */
        const unsigned long long v = score_sequence( s, w, c, n, best );
        if( v < best )
        {
            best = v;
            best_feasible = feasible;
            keep ^= 1;
        }
    }
    memcpy( s->trellis[n % ( GAP_MAX + 1 )], cand[keep ^ 1], n );
}

/* ================================================================================================================
 * scene cuts (slicetype.c:1384-1468)
 * ============================================================================================================== */
/* how much cheaper than intra a P picture must be not to count as a cut; grows with the distance from the last keyframe */
static float cut_margin( const x264cu_slicetype_t *s, int since_key )
{
    const float hi = s->p.scenecut_threshold / 100.0;
    float lo = hi * 0.25;
    if( s->p.keyint_min == s->p.keyint_max )
        lo = hi;
    if( since_key <= s->p.keyint_min / 4 || s->p.intra_refresh )
        return lo / 4;
    if( since_key <= s->p.keyint_min )
        return lo * since_key / s->p.keyint_min;
    return lo + ( hi - lo ) * ( since_key - s->p.keyint_min ) / ( s->p.keyint_max - s->p.keyint_min );
}

static int looks_like_cut( x264cu_slicetype_t *s, picture_t **w, int p0, int p1 )
{
    picture_t *f = w[p1];
    int as_intra, as_p;
    score3( s, w, p0, p1, p1 );
    estimates( s, f, 0, 0, &as_intra, NULL );
    estimates( s, f, p1 - p0, 0, &as_p, NULL );
    const float margin = cut_margin( s, f->number - s->last_key );
    return as_p >= ( 1.0 - margin ) * as_intra;
}

/* real = the decision for the first picture of the window: flashes (a cut that is undone a few pictures later) are filtered
 * out first by comparing pictures across the suspected cut */
static int scene_cut( x264cu_slicetype_t *s, picture_t **w, int p0, int p1, int real, int n_orig, int reach )
{
    if( real && s->p.la.bframes )
    {
        const int far = p0 + 1 + ( s->p.b_adapt == 2 ? s->p.la.bframes : 1 );
        const int last = far < n_orig ? far : n_orig;
        /* a picture further on that still matches p0: everything up to it was a flash */
        for( int q = p1; q <= last; q++ )
            if( !looks_like_cut( s, w, p0, q ) )
                for( int i = q; i > p0; i-- )
                    w[i]->cut_candidate = 0;
        /* a picture before the last one that does not match it: not a lasting change either */
        for( int q = p0; q <= last; q++ )
            if( far > reach || ( q < last && looks_like_cut( s, w, q, last ) ) )
                w[q]->cut_candidate = 0;
    }
    if( !w[p1]->cut_candidate )
        return 0;
    return looks_like_cut( s, w, p0, p1 );
}

/* ================================================================================================================
 * planner 2: MB-tree (slicetype.c:1091-1184) as a list of operations.  Walking back from the last anchor of the window, each
 * mini-GOP asks for its costs and pushes every picture's importance onto the pictures it was predicted from.
 * ============================================================================================================== */
enum { TREE_ASK, TREE_ZERO, TREE_SWAP, TREE_PUSH, TREE_SETTLE, TREE_PLAIN };
typedef struct { char op; char referenced; short a, b, c; } treeop_t;

static int tree_minigop( treeop_t *o, int left, int right, int pyramid )
{
    int n = 0;
    const int gap = right - left - 1;
    o[n++] = (treeop_t){ TREE_ASK, 0, left, right, right };
    o[n++] = (treeop_t){ TREE_ZERO, 0, left, 0, 0 };
    if( pyramid && gap > 1 )
    {
        const int mid = left + ( gap + 1 ) / 2;
        o[n++] = (treeop_t){ TREE_ASK, 0, left, right, mid };
        o[n++] = (treeop_t){ TREE_ZERO, 0, mid, 0, 0 };
        for( int b = right - 1; b > left; b-- )
            if( b != mid )
            {
                const int p0 = b > mid ? mid : left, p1 = b < mid ? mid : right;
                o[n++] = (treeop_t){ TREE_ASK, 0, p0, p1, b };
                o[n++] = (treeop_t){ TREE_PUSH, 0, p0, p1, b };
            }
        o[n++] = (treeop_t){ TREE_PUSH, 1, left, right, mid };
    }
    else
        for( int b = right - 1; b > left; b-- )
        {
            o[n++] = (treeop_t){ TREE_ASK, 0, left, right, b };
            o[n++] = (treeop_t){ TREE_PUSH, 0, left, right, b };
        }
    o[n++] = (treeop_t){ TREE_PUSH, 1, left, right, right };
    return n;
}

/* -> number of operations; n = pictures of the window taken into account, after_key = the window starts at a keyframe that is
 * itself part of the tree */
static int plan_tree( const x264cu_slicetype_t *s, picture_t **w, int n, int after_key, treeop_t *o )
{
    const int first = after_key ? 0 : 1;             /* the left-most picture that may receive importance */
    int k = 0, right = n, gap = 0;
    if( after_key )
        o[k++] = (treeop_t){ TREE_ASK, 0, 0, 0, 0 };
    while( right > 0 && bi_type( w[right]->type ) )
        right--;
    if( !s->p.rc_lookahead )
    {   /* no lookahead: the tree of the previous picture is carried over in the last anchor's array */
        if( after_key )
        {
            o[k++] = (treeop_t){ TREE_ZERO, 0, 0, 0, 0 };
            o[k++] = (treeop_t){ TREE_PLAIN, 0, 0, 0, 0 };
            return k;
        }
        o[k++] = (treeop_t){ TREE_SWAP, 0, right, 0, 0 };
        o[k++] = (treeop_t){ TREE_ZERO, 0, 0, 0, 0 };
    }
    else
    {
        if( right < first )
            return k;
        o[k++] = (treeop_t){ TREE_ZERO, 0, right, 0, 0 };
    }
    while( right > first )
    {
        int left = right - 1;
        while( left > 0 && bi_type( w[left]->type ) )
            left--;
        if( left < first )
            break;
        k += tree_minigop( o + k, left, right, s->p.b_pyramid );
        gap = right - left - 1;
        right = left;
    }
    if( !s->p.rc_lookahead )
    {
        o[k++] = (treeop_t){ TREE_ASK, 0, 0, right, right };
        o[k++] = (treeop_t){ TREE_PUSH, 1, 0, right, right };
        o[k++] = (treeop_t){ TREE_SWAP, 0, right, 0, 0 };
    }
    o[k++] = (treeop_t){ TREE_SETTLE, 0, right, right, 0 };
    if( s->p.b_pyramid && gap > 1 && !s->p.la.vbv )
        o[k++] = (treeop_t){ TREE_SETTLE, 0, right + ( gap + 1 ) / 2, 0, 0 };
    return k;
}

static float clamp_duration( float f ) { return f < 0.01f ? 0.01f : f > 1.00f ? 1.00f : f; }      /* CLIP_DURATION, ratecontrol.h */
#define TREE_PRECISION 0.5f                                                                       /* MBTREE_PRECISION */

static void tree_settle( x264cu_slicetype_t *s, picture_t *f, float mean_tick, int ref0_distance )
{   /* macroblock_tree_finish, slicetype.c:1029-1048 */
    const int factor = round( clamp_duration( mean_tick ) / clamp_duration( s->tick ) * 256 / TREE_PRECISION );
    if( x264cu_lookahead_mbtree_finish( s->la, f->slot, factor, ref0_distance, 5.0f * ( 1.0f - s->qcompress ) ) )
        s->broken = 1;
}

/* Every picture carries the same duration here (constant frame rate); the reference reads frame->f_duration, which
 * x264_slicetype_decide only assigns when a picture is decided (slicetype.c:1769) -- pictures still waiting carry whatever their
 * recycled x264_frame_t held: the same value once the pool has been through one cycle, zero (clamped to 0.01 s) before. */
static void grow_tree( x264cu_slicetype_t *s, picture_t **w, int n, int after_key )
{
    treeop_t ops[4 * WIN_MAX + 16];
    const int n_ops = plan_tree( s, w, n, after_key, ops );
    float total = 0.0;
    for( int j = 0; j <= n; j++ )
        total += s->tick;
    const float mean_tick = total / ( n + 1 );
    int slots[WIN_MAX + 4];
    for( int i = 0; i < n_ops && !s->broken; i++ )
    {
        const treeop_t *o = &ops[i];
        switch( o->op )
        {
            case TREE_ASK:
                score3( s, w, o->a, o->b, o->c );
                break;
            case TREE_ZERO:
                if( x264cu_lookahead_mbtree_reset( s->la, w[o->a]->slot ) ) s->broken = 1;
                break;
            case TREE_SWAP:
                if( x264cu_lookahead_mbtree_swap( s->la, w[o->a]->slot, w[0]->slot ) ) s->broken = 1;
                break;
            case TREE_PLAIN:      /* f_qp_offset = f_qp_offset_aq (slicetype.c:1121-1123) */
                if( x264cu_lookahead_mbtree_finish( s->la, w[o->a]->slot, 0, 0, 0.0f ) ) s->broken = 1;
                break;
            case TREE_PUSH:
            {   /* macroblock_tree_propagate, slicetype.c:1050-1089 */
                for( int j = o->a; j <= o->b; j++ )
                    slots[j] = w[j]->slot;
                const float factor = clamp_duration( s->tick ) / ( clamp_duration( mean_tick ) * 256.0f ) * TREE_PRECISION;
                if( x264cu_lookahead_mbtree_propagate( s->la, slots, o->a, o->b, o->c, o->referenced, factor ) ) s->broken = 1;
                if( s->vbv_plan && o->referenced )                                   /* slicetype.c:1087-1088 */
                    tree_settle( s, w[o->c], mean_tick, o->c == o->b ? o->c - o->a : 0 );
                break;
            }
            case TREE_SETTLE:
                tree_settle( s, w[o->a], mean_tick, o->b );
                break;
        }
    }
}

/* ================================================================================================================
 * VBV lookahead (slicetype.c:1186-1286): the types and costs of the pictures coded after the next anchor, for the rate control
 * ============================================================================================================== */
static int vbv_score( x264cu_slicetype_t *s, picture_t **w, int p0, int p1, int b )
{
    const int plain = score3( s, w, p0, p1, b );
    if( !s->p.la.aq_mode )
        return plain;
    if( s->p.la.mb_tree )
        return rescored( s, w[b], b - p0, p1 - b, NULL );
    int aq = 0;
    estimates( s, w[b], b - p0, p1 - b, NULL, &aq );
    return aq;
}

static void plan_for_vbv( x264cu_slicetype_t *s, picture_t **w, int n, int after_key )
{
    int left = 0, right = 1, k = 0;
    while( right < n && bi_type( w[right]->type ) )
        right++;
    picture_t *holder = w[after_key ? 0 : right];
    const int own = after_key ? -1 : right;                    /* the holder's own cost is not part of its plan */
    while( right < n )
    {
        if( right != own )
        {
            holder->plan_satd[k] = vbv_score( s, w, intra_type( w[right]->type ) ? right : left, right, right );
            holder->plan_type[k++] = w[right]->type;
        }
        for( int b = left + 1; b < right; b++ )                 /* coded after their anchor */
        {
            holder->plan_satd[k] = vbv_score( s, w, left, right, b );
            holder->plan_type[k++] = T_B;
        }
        left = right++;
        while( right <= n && bi_type( w[right]->type ) )
            right++;
    }
    holder->plan_type[k] = T_AUTO;
    holder->n_plan = k;
}

/* ================================================================================================================
 * the analysis of one window (x264_slicetype_analyse, slicetype.c:1473-1743), stage by stage
 * ============================================================================================================== */
typedef struct
{
    picture_t *pic[WIN_MAX + 3];         /* pic[0] = last anchor, pic[1..avail] = queued pictures */
    int avail;                           /* pictures looked at (framecnt) */
    int span, span_keyint;               /* pictures typed in this pass (num_frames) / before MB-tree widened it (orig_num_frames) */
    int reach;                           /* i_max_search */
    int after_key;                       /* second pass after a keyframe was decided (intra_minigop != 0) */
} window_t;

static void types_by_trellis( x264cu_slicetype_t *s, window_t *v )
{
    if( v->span <= 1 )
        return;
    memset( s->trellis, 0, sizeof( s->trellis ) );
    s->trellis[1][0] = K_P;
    for( int n = 2; n <= v->span; n++ )
        trellis_extend( s, v->pic, n );
    const unsigned char *best = s->trellis[v->span % ( GAP_MAX + 1 )];
    for( int j = 1; j < v->span; j++ )
    {
        picture_t *f = v->pic[j];
        if( best[j - 1] == K_B )
        {
            if( f->type == T_AUTO ) f->type = T_B;
        }
        else if( open_or_bi( f->type ) )
            f->type = T_P;
    }
}

/* b-adapt 1: one more B picture in the current run, or close it with a P?  Two short sequences from the run's anchor are
 * compared: (run of B) P P  against  (run of B) B P */
static void types_by_lookahead_of_one( x264cu_slicetype_t *s, window_t *v )
{
    const int most = s->p.la.bframes;
    int run_anchor = 0, room = most;
    for( int j = 1; j < v->span; j++ )
    {
        picture_t *f = v->pic[j];
        if( j > 1 && bi_type( v->pic[j - 1]->type ) )
            room--;
        else
        {
            run_anchor = j - 1;
            room = most;
        }
        if( !room )
        {
            if( open_or_bi( f->type ) ) f->type = T_P;
            continue;
        }
        if( f->type != T_AUTO )
            continue;
        if( bi_type( v->pic[j + 1]->type ) )
        {
            f->type = T_P;
            continue;
        }
        unsigned char seq[GAP_MAX + 3];
        const int run = j - run_anchor - 1;
        memset( seq, K_B, run );
        seq[run] = K_P; seq[run + 1] = K_P;
        const unsigned long long closed = score_sequence( s, v->pic + run_anchor, seq, run + 2, SCORE_INF );
        seq[run] = K_B;
        const unsigned long long longer = score_sequence( s, v->pic + run_anchor, seq, run + 2, closed );
        f->type = longer < closed ? T_B : T_P;
    }
}

static void types_fixed_pattern( x264cu_slicetype_t *s, window_t *v )
{
    int room = s->p.la.bframes;
    for( int j = 1; j < v->span; j++ )
    {
        picture_t *f = v->pic[j];
        if( !room )
        {
            if( open_or_bi( f->type ) ) f->type = T_P;
        }
        else if( f->type == T_AUTO )
            f->type = bi_type( v->pic[j + 1]->type ) ? T_P : T_B;
        room = bi_type( f->type ) ? room - 1 : s->p.la.bframes;
    }
}

/* keyint: no picture further than keyint_max from the last keyframe without becoming one; I pictures at least keyint_min after
 * it are promoted to IDR (slicetype.c:1668-1718) */
static void enforce_keyint( x264cu_slicetype_t *s, window_t *v )
{
    int key = s->last_key, candidate = 0;
    for( int j = 1; j <= v->span; j++ )
    {
        picture_t *f = v->pic[j];
        int dist = f->number - key;
        if( open_or_intra( f->asked ) && ( s->p.open_gop || !bi_type( v->pic[j - 1]->asked ) ) )
            candidate = j;                                   /* the last place a keyframe could go without breaking a forced type */
        if( dist >= s->p.keyint_max )
        {
            if( candidate && candidate != j )
            {
                j = candidate;
                f = v->pic[j];
                dist = f->number - key;
            }
            candidate = 0;
            if( f->type != T_IDR )
                f->type = s->p.open_gop ? T_I : T_IDR;
        }
        if( f->type == T_I && dist >= s->p.keyint_min )
        {
            if( s->p.open_gop )
                key = f->number;                             /* display order (no blu-ray compatibility mode) */
            else if( f->asked != T_I )
                f->type = T_IDR;
        }
        if( f->type == T_IDR )
        {
            key = f->number;
            if( j > 1 && bi_type( v->pic[j - 1]->type ) )
                v->pic[j - 1]->type = T_P;
        }
    }
}

static void analyse_window( x264cu_slicetype_t *s, int shifted )
{
    window_t win, *v = &win;
    memset( v->pic, 0, sizeof( v->pic ) );
    v->after_key = shifted != 0;
    v->reach = s->n_queue < WIN_MAX ? s->n_queue : WIN_MAX;
    if( v->reach > s->horizon + 1 - shifted )                 /* b_deterministic, slicetype.c:1480-1485 */
        v->reach = s->horizon + 1 - shifted;
    if( !s->anchor )
        return;
    v->pic[0] = s->anchor;
    for( v->avail = 0; v->avail < v->reach; v->avail++ )
        v->pic[v->avail + 1] = s->queue[v->avail];
    if( !v->avail )
    {
        if( s->p.la.mb_tree )
            grow_tree( s, v->pic, 0, v->after_key );
        return;
    }
    const int to_keyint = s->p.keyint_max - v->pic[0]->number + s->last_key - 1;
    v->span_keyint = v->span = s->p.intra_refresh ? v->avail : v->avail < to_keyint ? v->avail : to_keyint;
    if( ( s->p.psy && s->p.la.mb_tree ) || s->vbv_plan )
        v->span = v->avail;                                   /* the tree / the plan want the whole window */
    else if( s->p.open_gop && v->span < v->avail )
        v->span++;
    else if( !v->span )
    {
        v->pic[1]->type = T_I;
        return;
    }
    picture_t **w = v->pic;
    if( open_or_intra( w[1]->type ) && s->p.scenecut_threshold && scene_cut( s, w, 0, 1, 1, v->span_keyint, v->reach ) )
    {
        if( w[1]->type == T_AUTO )
            w[1]->type = T_I;
        return;
    }
    for( int j = 1; j <= v->span; j++ )                       /* forced keyframes, slicetype.c:1534-1539 */
        if( w[j]->type == T_KEY )
            w[j]->type = s->p.open_gop ? T_I : T_IDR;
    for( int j = 2; j <= v->span; j++ )                       /* nothing may lean across an IDR picture */
        if( w[j]->type == T_IDR && open_or_bi( w[j - 1]->type ) )
            w[j - 1]->type = T_P;

    int typed = v->span, undo_from;
    if( s->p.la.bframes )
    {
        if( s->p.b_adapt == 2 ) types_by_trellis( s, v );
        else if( s->p.b_adapt == 1 ) types_by_lookahead_of_one( s, v );
        else types_fixed_pattern( s, v );
        if( open_or_bi( w[v->span]->type ) )
            w[v->span]->type = T_P;
        int lead_b = 0;
        while( lead_b < v->span && bi_type( w[lead_b + 1]->type ) )
            lead_b++;
        /* a cut inside the first mini-GOP ends it there */
        for( int j = 1; j < lead_b + 1; j++ )
            if( w[j]->asked == T_AUTO && open_or_intra( w[j + 1]->asked ) && s->p.scenecut_threshold &&
                scene_cut( s, w, j, j + 1, 0, v->span_keyint, v->reach ) )
            {
                w[j]->type = T_P;
                typed = j;
                break;
            }
        undo_from = v->after_key ? 1 : ( lead_b + 2 < typed + 1 ? lead_b + 2 : typed + 1 );
    }
    else
    {
        for( int j = 1; j <= v->span; j++ )
            if( open_or_bi( w[j]->type ) )
                w[j]->type = T_P;
        undo_from = v->after_key ? 1 : 2;
    }
    if( s->p.la.mb_tree )
        grow_tree( s, w, v->span < s->p.keyint_max ? v->span : s->p.keyint_max, v->after_key );
    if( !s->p.intra_refresh )
        enforce_keyint( s, v );
    if( s->vbv_plan )
        plan_for_vbv( s, w, v->span, v->after_key );
    for( int j = undo_from; j <= v->span; j++ )               /* only the first mini-GOP is final */
        w[j]->type = w[j]->asked;
}

/* ================================================================================================================
 * x264_slicetype_decide (slicetype.c:1745-1974): close the next mini-GOP, ask for the rate control's costs, coded order
 * ============================================================================================================== */
static void decide_minigop( x264cu_slicetype_t *s )
{
    if( !s->n_queue )
        return;
    if( ( s->p.la.bframes && s->p.b_adapt ) || s->p.scenecut_threshold || s->p.la.mb_tree || s->vbv_plan )
        analyse_window( s, 0 );

    int n_b = 0, n_bref = 0;
    for( ;; n_b++ )
    {
        picture_t *f = s->queue[n_b];
        const int since_key = f->number - s->last_key;
        if( f->type == T_BREF )
        {   /* no room for another reference B */
            if( s->p.b_pyramid < 2 ? n_bref == s->p.b_pyramid : ( n_bref && s->p.frame_reference <= n_bref + 3 ) )
                f->type = T_B;
        }
        if( f->type == T_KEY )
            f->type = s->p.open_gop ? T_I : T_IDR;
        if( ( !s->p.intra_refresh || f->number == 0 ) && since_key >= s->p.keyint_max )
        {   /* GOP size limit, slicetype.c:1831-1845 */
            if( f->type != T_IDR )                           /* whatever was wanted here, a keyframe is due */
                f->type = s->p.open_gop && s->last_key >= 0 ? T_I : T_IDR;
        }
        if( f->type == T_I && since_key >= s->p.keyint_min )
        {
            if( s->p.open_gop )
                s->last_key = f->number;
            else
                f->type = T_IDR;
        }
        if( f->type == T_IDR )
        {
            s->last_key = f->number;
            if( n_b > 0 )
                s->queue[--n_b]->type = T_P;                  /* the picture before an IDR closes its mini-GOP */
        }
        if( ( n_b == s->p.la.bframes || n_b + 1 >= s->n_queue ) && open_or_bi( f->type ) )
            f->type = T_P;
        if( f->type == T_BREF )
            n_bref++;
        if( f->type == T_AUTO )
            f->type = T_B;
        else if( !bi_type( f->type ) )
            break;
    }
    picture_t *closing = s->queue[n_b];
    closing->trailing_b = n_b;
    if( s->p.b_pyramid && n_b > 1 && !n_bref )
    {
        s->queue[( n_b - 1 ) / 2]->type = T_BREF;
        n_bref++;
    }
    picture_t *w[GAP_MAX + 2];
    if( !s->p.rc_cqp )
    {   /* the costs x264_rc_analyse_slice will read (slicetype.c:1896-1935) */
        w[0] = s->anchor;
        memcpy( w + 1, s->queue, ( n_b + 1 ) * sizeof( picture_t * ) );
        const int last = n_b + 1, from = intra_type( closing->type ) ? last : 0;
        score3( s, w, from, last, last );
        closing->rc0 = last - from; closing->rc1 = 0;
        if( ( from != last || n_b ) && s->p.la.vbv )
        {   /* row SATDs: the intra costs and the B pictures' costs, slicetype.c:1916-1934 */
            score3( s, w, last, last, last );
            int p0 = 0;
            for( int b = 1; b <= n_b; b++ )
            {
                int p1 = last;
                if( w[b]->type == T_B )
                    for( p1 = b; w[p1]->type == T_B; )
                        p1++;
                score3( s, w, p0, p1, b );
                w[b]->rc0 = b - p0; w[b]->rc1 = p1 - b;
                if( w[b]->type == T_BREF )
                    p0 = b;
            }
        }
    }
    if( n_b )
    {   /* coded order: anchor, reference B pictures, plain B pictures */
        int at_bref = 1, at_b = 1 + n_bref;
        for( int i = 0; i < n_b; i++ )
            w[s->queue[i]->type == T_BREF ? at_bref++ : at_b++] = s->queue[i];
        w[0] = closing;
        memcpy( s->queue, w, ( n_b + 1 ) * sizeof( picture_t * ) );
    }
}

/* ================================================================================================================
 * the picture queue (encoder/lookahead.c:192-250 without a lookahead thread)
 * ============================================================================================================== */
static void drop_picture( x264cu_slicetype_t *s, picture_t *f )
{
    if( !f ) return;
    s->slot_busy[f->slot] = 0;
    for( int k = 0; k < s->n_recent; k++ )
        if( s->recent[k] == f )
        {
            memmove( s->recent + k, s->recent + k + 1, ( s->n_recent - k - 1 ) * sizeof( picture_t * ) );
            s->n_recent--;
            break;
        }
    free( f );
}

static void pull_decided( x264cu_slicetype_t *s )
{
    if( s->n_ready || !s->n_queue )
        return;
    decide_minigop( s );
    picture_t *was = s->anchor;
    s->anchor = s->queue[0];
    const int moved = s->anchor->trailing_b + 1;
    for( int i = 0; i < moved; i++ )
        s->ready[s->n_ready++] = s->queue[i];
    memmove( s->queue, s->queue + moved, ( s->n_queue - moved ) * sizeof( picture_t * ) );
    s->n_queue -= moved;
    s->queue[s->n_queue] = NULL;
    if( was )
    {
        int held = 0;
        for( int i = 0; i < s->n_ready; i++ )
            held |= s->ready[i] == was;
        if( !held )
            drop_picture( s, was );
    }
    if( s->keyframe_pass && intra_type( s->anchor->type ) )
        analyse_window( s, moved );       /* MB-tree / the VBV plan also want the tree behind a keyframe, lookahead.c:243-245 */
}

int x264cu_slicetype_open( x264cu_ctx_t *ctx, const x264cu_slicetype_params_t *p, x264cu_slicetype_t **out )
{
    if( !ctx || !p || !out ) return -1;
    *out = NULL;
    /* rc.i_lookahead is clipped to X264_LOOKAHEAD_MAX by the reference (encoder.c:1112); the windows are sized for it.  Not
     * built: the intra-refresh column correction of the VBV path (slicetype.c:2015-2036), hence no VBV with intra refresh. */
    if( p->rc_lookahead < 0 || p->rc_lookahead > WIN_MAX || p->la.bframes < 0 || p->la.bframes > GAP_MAX ||
        ( p->la.vbv && p->intra_refresh ) || p->keyint_max < 1 || p->keyint_min < 1 || p->b_adapt < 0 || p->b_adapt > 2 ||
        p->b_pyramid < 0 || p->b_pyramid > 2 || p->qcompress > 1.0f )
        return -1;
    x264cu_slicetype_t *s = calloc( 1, sizeof( *s ) );
    if( !s ) return -1;
    s->ctx = ctx;
    s->p = *p;
    /* frames.i_delay, encoder.c:1602-1609 */
    s->horizon = p->b_adapt == 2 ? ( p->la.bframes > 3 ? p->la.bframes : 3 ) * 4 : p->la.bframes;
    if( ( p->la.mb_tree || p->la.vbv ) && p->rc_lookahead > s->horizon )
        s->horizon = p->rc_lookahead;
    s->vbv_plan = p->la.vbv && p->rc_lookahead;
    s->keyframe_pass = p->la.mb_tree || s->vbv_plan;
    s->last_key = -p->keyint_max;
    s->n_slots = s->horizon + p->la.bframes + 8 + AHEAD_MAX;
    s->slot_busy = calloc( s->n_slots, 1 );
    s->p.la.n_slots = s->n_slots;
    s->mb_w = ( p->la.width + 15 ) >> 4;
    s->mb_h = ( p->la.height + 15 ) >> 4;
    {   /* f_duration of a progressive picture at a constant frame rate (slicetype.c:1769-1771): i_duration = 2 ticks of
         * num_units_in_tick / time_scale = fps_den / (2 fps_num) */
        const int num = p->fps_num > 0 ? p->fps_num : 25, den = p->fps_den > 0 ? p->fps_den : 1;
        s->tick = (double)2 * den / ( 2.0 * num );
    }
    /* negative = "the reference's default"; zero is a legal value of both (qcompress 0: MB-tree strength 5; aq_strength 0:
     * adaptive quantisation off, encoder.c:1094-1097) */
    s->qcompress = p->qcompress < 0 ? 0.6f : p->qcompress;
    s->aq_strength = p->aq_strength < 0 ? 1.0f : p->aq_strength;
    if( s->aq_strength == 0 )
        s->p.la.aq_mode = 0;
    s->prefetch = 1;
    s->speculate = -1;
    /* measured at 4K (B200), pictures per launch / run-ahead: 4/8 -> 1000 pictures/s, 8/16 -> 1260, 12/24 -> 1410: a launch
     * needs several dozen independent wavefronts to fill the 148 SMs */
    s->group = s->horizon >= 12 ? 12 : 1;
    s->run_ahead = s->horizon >= 12 ? 24 : 0;
    if( !s->slot_busy || x264cu_lookahead_open( ctx, &s->p.la, &s->la ) )
    {
        free( s->slot_busy );
        free( s );
        return -1;
    }
    *out = s;
    return 0;
}

void x264cu_slicetype_close( x264cu_slicetype_t *s )
{
    if( !s ) return;
    for( int i = 0; i < s->n_queue; i++ ) free( s->queue[i] );
    for( int i = 0; i < s->n_ready; i++ ) if( s->ready[i] != s->anchor ) free( s->ready[i] );
    if( s->out && s->out != s->anchor ) free( s->out );
    free( s->anchor );
    x264cu_lookahead_close( s->la );
    free( s->slot_busy );
    free( s );
}

/* ================================================================================================================
 * prefetch and the sharded stream
 * ============================================================================================================== */

/* Sharded stream: all-gather the results of the group's searches (queued behind them on the exchange stream: the calling thread
 * does not wait) and install the ones searched on other GPUs.  Every rank holds the same job list, so the layout of the exchange is known everywhere:
 * rank r's k-th job sits at r * bytes_per_rank + k * search_bytes. */
static int exchange_group( x264cu_slicetype_t *s )
{
    if( !s->n_sent ) return 0;
    int per_owner[64] = { 0 }, most = 0;
    for( int i = 0; i < s->n_sent; i++ ) per_owner[s->sent[i].owner]++;
    for( int r = 0; r < s->world; r++ ) if( per_owner[r] > most ) most = per_owner[r];
    const size_t rec = x264cu_lookahead_search_bytes( s->la ), per_rank = (size_t)most * rec;
    void *d_send = NULL, *d_recv = NULL, *stream = x264cu_lookahead_exchange_stream( s->la );
    if( s->exchange( s->exchange_user, 0, per_rank, &d_send, &d_recv, stream ) || !d_send || !d_recv )
        return -1;
    int k = 0;
    for( int i = 0; i < s->n_sent; i++ )
        if( s->sent[i].owner == s->rank )
        {   /* a picture that has left its slot since (end of stream) has nothing to send: the block stays as it is */
            if( x264cu_slicetype_slot_of( s, s->sent[i].number ) == s->sent[i].slot &&
                x264cu_lookahead_export_search( s->la, s->sent[i].slot, s->sent[i].list, s->sent[i].dist, (char *)d_send + (size_t)k * rec ) )
                return -1;
            k++;
        }
    if( s->exchange( s->exchange_user, 1, per_rank, &d_send, &d_recv, stream ) )
        return -1;
    int seen[64] = { 0 };
    for( int i = 0; i < s->n_sent; i++ )
    {
        const int r = s->sent[i].owner, at = seen[r]++;
        if( r == s->rank || x264cu_slicetype_slot_of( s, s->sent[i].number ) != s->sent[i].slot )
            continue;
        if( x264cu_lookahead_import_search( s->la, s->sent[i].slot, s->sent[i].list, s->sent[i].dist,
                                            (char *)d_recv + (size_t)r * per_rank + (size_t)at * rec ) )
            return -1;
    }
    s->n_sent = 0;
    return x264cu_lookahead_import_done( s->la );
}

/* Every cost request the decision can make once these searches exist: a list-0 search of picture X against picture Y, d = X - Y
 * apart, completes the P triple (Y, X, X) and the B triples (Y, X, b) of the pictures in between (their own two searches were
 * launched when b and X arrived). */
static int speculate_group( x264cu_slicetype_t *s, int n, const int *fenc, const int *ref, const int *list, const int *dist, const int *number )
{
    enum { TRIPLES_MAX = 4096 };
    int *t = malloc( 6 * TRIPLES_MAX * sizeof( int ) ), m = 0;
    if( !t ) return -1;
    int *tb = t, *t0 = t + TRIPLES_MAX, *t1 = t + 2 * TRIPLES_MAX, *td0 = t + 3 * TRIPLES_MAX, *td1 = t + 4 * TRIPLES_MAX, *own = t + 5 * TRIPLES_MAX;
    /* With a B pyramid the trellis (slicetype_path_cost, slicetype.c:1310-1321) and MB-tree (:1143-1160) price a mini-GOP Y..X longer
     * than two as: its middle picture against (Y, X), the pictures of each half against (Y, middle) / (middle, X).  So of the spans
     * longer than a half can be, (bframes+1) - (bframes+1)/2, only the middle picture's triple is ever asked for: 61 instead of 150
     * triples per picture at bframes 16.  (Anything else is still answered, on demand.) */
    const int pyramid_paths = s->p.b_pyramid && s->p.b_adapt == 2 && s->p.la.bframes > 1 && !s->p.la.vbv;
    const int half_span = ( s->p.la.bframes + 1 ) - ( s->p.la.bframes + 1 ) / 2;
    for( int i = 0; i < n; i++ )
    {
        if( list[i] )
            continue;
        const int x_no = number[i], d = dist[i];
        for( int k = 0; k < d && m < TRIPLES_MAX; k++ )
        {   /* k = 0: the P triple; k > 0: the B picture k pictures before X */
            if( k && pyramid_paths && d > half_span && k != d - d / 2 )
                continue;
            const int b_slot = k ? x264cu_slicetype_slot_of( s, x_no - k ) : fenc[i];
            if( b_slot < 0 )
                continue;
            tb[m] = b_slot; t0[m] = ref[i]; t1[m] = fenc[i]; td0[m] = d - k; td1[m] = k;
            own[m] = s->world > 1 ? ( x_no - k ) % s->world : 0;    /* the rank that searched picture b computes its costs too */
            m++;
        }
    }
    const int rc = !m ? 0 : s->world > 1 ? x264cu_lookahead_finalize_batch_sharded( s->la, m, tb, t0, t1, td0, td1, own, s->rank, s->world,
                                                                                    s->exchange, s->exchange_user )
                                         : x264cu_lookahead_finalize_batch( s->la, m, tb, t0, t1, td0, td1 );
    free( t );
    return rc;
}

/* launch the gathered searches; jobs whose pictures have left their slots in the meantime are dropped */
static int launch_group( x264cu_slicetype_t *s )
{
    int fenc[JOBS_MAX], ref[JOBS_MAX], list[JOBS_MAX], dist[JOBS_MAX], number[JOBS_MAX], n = 0;
    for( int i = 0; i < s->n_job; i++ )
    {
        if( x264cu_slicetype_slot_of( s, s->job[i].fenc_no ) != s->job[i].fenc_slot || x264cu_slicetype_slot_of( s, s->job[i].ref_no ) != s->job[i].ref_slot )
            continue;
        if( s->p.la.weighted_pred && s->job[i].list == 0 )
        {   /* a list-0 search is weighted if it is first requested as a P cost and the analysis picks a weight
             * (slicetype.c:857-864): it is a pure function of the two pictures only where the analysis cannot pick one */
            const int t = x264cu_lookahead_weight_trivial( s->la, s->job[i].fenc_slot, s->job[i].ref_slot );
            if( t < 0 ) return -1;
            if( !t ) continue;
        }
        fenc[n] = s->job[i].fenc_slot; ref[n] = s->job[i].ref_slot; list[n] = s->job[i].list; dist[n] = s->job[i].dist;
        number[n] = s->job[i].fenc_no;
        n++;
    }
    s->n_job = 0;
    s->in_group = 0;
    int all = n;
    int a_fenc[JOBS_MAX], a_ref[JOBS_MAX], a_list[JOBS_MAX], a_dist[JOBS_MAX];
    if( s->world > 1 )
    {   /* this rank's share of the searches; the results are exchanged right behind them (nothing waits on the host) */
        memcpy( a_fenc, fenc, n * sizeof( int ) ); memcpy( a_ref, ref, n * sizeof( int ) );
        memcpy( a_list, list, n * sizeof( int ) ); memcpy( a_dist, dist, n * sizeof( int ) );
        int mine = 0;
        for( int i = 0; i < n; i++ )
        {
            const int owner = number[i] % s->world;
            s->sent[i].slot = fenc[i]; s->sent[i].list = list[i]; s->sent[i].dist = dist[i];
            s->sent[i].owner = owner; s->sent[i].number = number[i];
            if( owner == s->rank )
            {
                fenc[mine] = fenc[i]; ref[mine] = ref[i]; list[mine] = list[i]; dist[mine] = dist[i];
                mine++;
            }
        }
        s->n_sent = n;
        n = mine;
    }
    if( n && x264cu_lookahead_search_batch( s->la, n, fenc, ref, list, dist ) )
        return -1;
    if( s->world > 1 && exchange_group( s ) )
        return -1;
    /* Sharded, the cost requests are the replicated part that has to be split: always.  On one GPU it depends on the window: with
     * the trellis over a B pyramid a new GOP's first analysis asks for ~18 000 triples at once (each one a launch + read-back when
     * computed on demand), and computing them with the searches wins (8K, bframes 16: 57 -> 64 pictures/s); with the short windows
     * of b-adapt 1 the lookahead is bound by the searches' throughput and the extra triples cost more than the waits they save
     * (4K preset medium: 1 396 -> 1 327 pictures/s). */
    const int long_windows = s->p.b_adapt == 2 && s->p.b_pyramid && s->p.la.bframes > 1 && !s->p.la.vbv;
    if( !( s->speculate < 0 ? s->world > 1 || long_windows : s->speculate ) )
        return 0;
    return s->world > 1 ? speculate_group( s, all, a_fenc, a_ref, a_list, a_dist, number )
                        : speculate_group( s, n, fenc, ref, list, dist, number );
}

/* every (picture, earlier picture) pair the decision could ask about: list 0 at distance d <= bframes+1 from the new
 * picture, list 1 at distance d <= bframes towards it -- or, with a B pyramid, no further than half a mini-GOP: whoever prices
 * a B picture then does it against the middle picture or the nearer anchor (the b-adapt 1 loop asks for (i, i+2, i+1) only,
 * slicetype.c:1610; trellis :1310-1321; MB-tree :1143-1160; the rate control's nearest references, ratecontrol.c; only the VBV
 * plan, :1262-1266, spans the whole mini-GOP).  Anything else would still be searched on demand. */
static void note_searches_of( x264cu_slicetype_t *s, picture_t *f )
{
    const int span = s->p.la.bframes + 1;
    const int l1_max = s->p.b_pyramid && s->p.la.bframes > 1 && !s->p.la.vbv ? span - span / 2 : s->p.la.bframes;
    for( int k = 0; k < s->n_recent; k++ )
    {
        picture_t *o = s->recent[k];
        const int d = f->number - o->number;
        if( !s->slot_busy[o->slot] || d < 1 || d > s->p.la.bframes + 1 || s->n_job + 2 > JOBS_MAX )
            continue;
        s->job[s->n_job].fenc_slot = f->slot; s->job[s->n_job].ref_slot = o->slot; s->job[s->n_job].list = 0; s->job[s->n_job].dist = d;
        s->job[s->n_job].fenc_no = f->number; s->job[s->n_job].ref_no = o->number;
        s->n_job++;
        if( d <= l1_max )
        {
            s->job[s->n_job].fenc_slot = o->slot; s->job[s->n_job].ref_slot = f->slot; s->job[s->n_job].list = 1; s->job[s->n_job].dist = d;
            s->job[s->n_job].fenc_no = o->number; s->job[s->n_job].ref_no = f->number;
            s->n_job++;
        }
    }
}

/* ================================================================================================================
 * one step of the encoder's loop as far as the lookahead is concerned (encoder.c:3323-3450)
 * ============================================================================================================== */
typedef struct { const uint8_t *luma, *cb, *cr; intptr_t stride, cstride; const uint16_t *qscale; int on_device; } input_t;

static int step( x264cu_slicetype_t *s, const input_t *in, int *out_frame, int *out_type )
{
    if( !s || !out_frame || !out_type ) return -1;
    *out_frame = -1; *out_type = T_AUTO;
    if( s->out )
    {   /* the picture returned by the previous call leaves now (its slot with it), unless it is still the last anchor */
        if( s->out != s->anchor ) drop_picture( s, s->out );
        s->out = NULL;
    }
    if( in->luma )
    {
        int slot = 0;
        while( slot < s->n_slots && s->slot_busy[slot] )
            slot++;
        if( slot == s->n_slots || s->n_queue >= QUEUE_MAX - 4 ) return -1;
        const int rc = in->on_device ? x264cu_lookahead_frame_put_device( s->la, slot, in->luma, in->stride, in->qscale )
                     : in->cb        ? x264cu_lookahead_frame_put_i420( s->la, slot, in->luma, in->stride, in->cb, in->cr, in->cstride,
                                                                        s->p.la.aq_mode, s->aq_strength )
                                     : x264cu_lookahead_frame_put( s->la, slot, in->luma, in->stride, in->qscale );
        if( rc ) return -1;
        picture_t *f = calloc( 1, sizeof( *f ) );
        if( !f ) return -1;
        memset( f->memo, -1, sizeof( f->memo ) );
        f->number = s->fed++;
        f->slot = slot;
        f->type = f->asked = s->next_asked;                  /* x264_frame_copy_picture, frame.c:370-376 */
        s->next_asked = T_AUTO;
        f->cut_candidate = 1;
        s->slot_busy[slot] = 1;
        s->queue[s->n_queue++] = f;
        s->queue[s->n_queue] = NULL;
        if( s->prefetch )
        {
            note_searches_of( s, f );
            if( ++s->in_group >= s->group && launch_group( s ) ) return -1;
        }
        const int keep = s->p.la.bframes + 1;
        if( s->n_recent < keep ) s->n_recent++;
        memmove( s->recent + 1, s->recent, ( s->n_recent - 1 ) * sizeof( picture_t * ) );
        s->recent[0] = f;
        if( s->fed <= s->horizon + s->run_ahead )              /* encoder.c:3428: nothing to encode yet (i_delay includes the sync-lookahead pictures) */
            return 0;
    }
    else
    {
        if( s->n_job && launch_group( s ) ) return -1;
        if( s->world > 1 && s->n_sent && exchange_group( s ) ) return -1;
    }
    pull_decided( s );
    if( s->broken ) return -1;
    if( !s->n_ready )
        return 0;
    picture_t *f = s->ready[0];
    memmove( s->ready, s->ready + 1, ( s->n_ready - 1 ) * sizeof( picture_t * ) );
    s->n_ready--;
    *out_frame = f->number;
    *out_type = f->type;
    s->out = f;
    return 0;
}

int x264cu_slicetype_step( x264cu_slicetype_t *s, const uint8_t *h_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                           int *out_frame, int *out_type )
{
    const input_t in = { h_luma, NULL, NULL, luma_stride, 0, h_inv_qscale, 0 };
    return step( s, &in, out_frame, out_type );
}

int x264cu_slicetype_step_i420( x264cu_slicetype_t *s, const uint8_t *h_luma, intptr_t luma_stride, const uint8_t *h_cb, const uint8_t *h_cr,
                                intptr_t chroma_stride, int *out_frame, int *out_type )
{
    if( h_luma && ( !h_cb || !h_cr ) ) return -1;
    const input_t in = { h_luma, h_luma ? h_cb : NULL, h_cr, luma_stride, chroma_stride, NULL, 0 };
    return step( s, &in, out_frame, out_type );
}

int x264cu_slicetype_step_device( x264cu_slicetype_t *s, const uint8_t *d_luma, intptr_t luma_stride, const uint16_t *h_inv_qscale,
                                  int *out_frame, int *out_type )
{
    const input_t in = { d_luma, NULL, NULL, luma_stride, 0, h_inv_qscale, 1 };
    return step( s, &in, out_frame, out_type );
}

void x264cu_slicetype_set_prefetch( x264cu_slicetype_t *s, int prefetch ) { if( s ) s->prefetch = !!prefetch; }

void x264cu_slicetype_set_run_ahead( x264cu_slicetype_t *s, int pictures )
{
    if( s && !s->fed && pictures >= 0 && pictures <= AHEAD_MAX ) s->run_ahead = pictures;
}

void x264cu_slicetype_set_speculation( x264cu_slicetype_t *s, int speculate ) { if( s ) s->speculate = !!speculate; }

void x264cu_slicetype_set_prefetch_group( x264cu_slicetype_t *s, int pictures )
{
    if( s && !s->fed && pictures >= 1 && pictures <= 32 ) s->group = pictures;
}

int x264cu_slicetype_set_shard( x264cu_slicetype_t *s, int rank, int world, x264cu_exchange_fn fn, void *user )
{
    if( !s || s->fed || world < 1 || world > 64 || rank < 0 || rank >= world || ( world > 1 && !fn ) ) return -1;
    s->rank = rank; s->world = world; s->exchange = fn; s->exchange_user = user;
    return 0;
}

int x264cu_slicetype_set_next_type( x264cu_slicetype_t *s, int type )
{
    if( !s || type < T_AUTO || type > T_KEY ) return -1;
    s->next_asked = type;
    return 0;
}

void x264cu_slicetype_set_async_upload( x264cu_slicetype_t *s, int on ) { if( s ) x264cu_lookahead_set_async_upload( s->la, on ); }

x264cu_lookahead_t *x264cu_slicetype_lookahead( x264cu_slicetype_t *s ) { return s ? s->la : NULL; }

static picture_t *held_picture( x264cu_slicetype_t *s, int frame )
{
    if( s->out && s->out->number == frame ) return s->out;
    if( s->anchor && s->anchor->number == frame ) return s->anchor;
    for( int i = 0; i < s->n_ready; i++ ) if( s->ready[i]->number == frame ) return s->ready[i];
    return NULL;
}

int x264cu_slicetype_slot_of( x264cu_slicetype_t *s, int frame )
{
    if( !s ) return -1;
    picture_t *f = held_picture( s, frame );
    if( f ) return f->slot;
    for( int i = 0; i < s->n_queue; i++ ) if( s->queue[i]->number == frame ) return s->queue[i]->slot;
    return -1;
}

int x264cu_slicetype_get_qp_offset( x264cu_slicetype_t *s, int frame, float *h_qp_offset )
{
    if( !s || !h_qp_offset ) return -1;
    const int slot = x264cu_slicetype_slot_of( s, frame );
    return slot < 0 ? -1 : x264cu_lookahead_get_qp_offset( s->la, slot, h_qp_offset );
}

/* x264_rc_analyse_slice, slicetype.c:1976-2030, for a picture the last x264cu_slicetype_step returned */
int x264cu_slicetype_rc_analyse_slice( x264cu_slicetype_t *s, int frame, int *cost_out, int *h_row_satd, int *h_row_satd_intra )
{
    if( !s || !cost_out || s->p.rc_cqp ) return -1;
    picture_t *f = held_picture( s, frame );
    if( !f ) return -1;
    if( bi_type( f->type ) && !s->p.la.vbv ) return -1;      /* their costs are requested only for the VBV row SATDs (slicetype.c:1916) */
    const int d0 = intra_type( f->type ) ? 0 : f->rc0, d1 = intra_type( f->type ) ? 0 : f->rc1;
    int cost = 0, aq = 0;
    estimates( s, f, d0, d1, &cost, &aq );
    if( s->broken || cost < 0 ) return -1;
    if( s->p.la.mb_tree )
    {
        cost = rescored( s, f, d0, d1, h_row_satd );
        if( !intra_type( f->type ) && s->p.la.vbv )
            rescored( s, f, 0, 0, NULL );                      /* the intra rows with the same offsets */
    }
    else
    {
        if( s->p.la.aq_mode ) cost = aq;
        if( h_row_satd && x264cu_lookahead_get_row_satds( s->la, f->slot, d0, d1, h_row_satd ) ) s->broken = 1;
    }
    if( h_row_satd_intra && !intra_type( f->type ) && x264cu_lookahead_get_row_satds( s->la, f->slot, 0, 0, h_row_satd_intra ) ) s->broken = 1;
    *cost_out = cost;
    return s->broken ? -1 : 0;
}

/* i_planned_type / i_planned_satd of a non-B picture the last step returned (VBV lookahead); returns the number of entries */
int x264cu_slicetype_get_planned( x264cu_slicetype_t *s, int frame, int *h_type, int *h_satd, int max_entries )
{
    if( !s || !s->vbv_plan ) return -1;
    picture_t *f = held_picture( s, frame );
    if( !f ) return -1;
    const int n = f->n_plan < max_entries ? f->n_plan : max_entries;
    for( int i = 0; i < n; i++ )
    {
        if( h_type ) h_type[i] = f->plan_type[i];
        if( h_satd ) h_satd[i] = f->plan_satd[i];
    }
    return n;
}

long x264cu_slicetype_cost_requests( x264cu_slicetype_t *s ) { return s ? s->requests : 0; }
int x264cu_slicetype_farthest_list1( x264cu_slicetype_t *s ) { return s ? s->far_list1 : 0; }
