/* TEST INFRASTRUCTURE -- CPU restatement (oracle) of x264's adaptive quantisation for the lookahead's inputs:
 * x264_adaptive_quant_frame with aq-mode 0 / 1 (encoder/ratecontrol.c:225-420): per-macroblock AC energy of luma and both chroma
 * planes (ac_energy_mb, :261-303; pixel_var, common/pixel.c:183-203), f_qp_offset_aq = strength * (x264_log2(energy) - 14.427)
 * (:397), i_inv_qscale_factor = x264_exp2fix8( qp ) (common/base.h:218-224), frame sums i_pixel_sum / i_pixel_ssd (:405-414).
 * Auto-variance modes 2 / 3 (:352-392): qp = (energy+1)^(1/8), frame means of qp and qp^2 accumulated in raster order in single
 * precision -- in the form the reference's -ffast-math build computes them (powf(x, .125f) = three square roots, qp*qp taken
 * from the second one; disassembly of oracle/_ref).  Pinned against the compiled reference by tests/test_oracle_aq.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file. */
#include "oracle.h"
#include <math.h>
#include <string.h>

static int exp2fix8( float x )                                   /* x264_exp2fix8, common/base.h:218-224 */
{
    static uint8_t lut[64];
    static int init = 0;
    if( !init )
    {   /* x264_exp2_lut (common/tables.c:58-64): round( 256 * (2^(i/64) - 1) ) */
        for( int i = 0; i < 64; i++ ) lut[i] = (uint8_t)( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) + 0.5 );
        init = 1;
    }
    int i = x * ( -64.f / 6.f ) + 512.5f;
    if( i < 0 ) return 0;
    if( i > 1023 ) return 0xffff;
    return ( lut[i & 63] + 256 ) << ( i >> 6 ) >> 8;
}

/* pixel_var_WxH: sum + (sum of squares << 32); ac_energy_var: ssd - (sum*sum >> shift), 32-bit */
static uint32_t energy_block( const uint8_t *p, intptr_t stride, int x0, int y0, int bw, int bh, int pw, int ph, int shift,
                              uint64_t *acc_sum, uint64_t *acc_ssd )
{
    uint32_t sum = 0, sqr = 0;
    for( int y = 0; y < bh; y++ )
        for( int x = 0; x < bw; x++ )
        {   /* the picture is edge-replicated to the macroblock grid (x264_frame_expand_border_mod16, frame.c:640-665) */
            int yy = y0 + y < ph ? y0 + y : ph - 1, xx = x0 + x < pw ? x0 + x : pw - 1;
            uint32_t v = p[(intptr_t)yy * stride + xx];
            sum += v; sqr += v * v;
        }
    *acc_sum += sum; *acc_ssd += sqr;
    return sqr - (uint32_t)( (uint64_t)sum * sum >> shift );
}

/* luma w x h (stride), cb / cr (w+1)/2 x (h+1)/2 (cstride).  aq_mode 0: offsets 0, factors 256 (the MB-tree initialisation,
 * ratecontrol.c:314-336).  stats[6] = i_pixel_sum[3], i_pixel_ssd[3] after the mean removal. */
void orc_adaptive_quant_frame( const uint8_t *luma, intptr_t stride, const uint8_t *cb, const uint8_t *cr, intptr_t cstride,
                               int width, int height, int aq_mode, float aq_strength, float *qp_offset_aq, uint16_t *inv_qscale,
                               uint64_t *stats )
{
    const int mb_w = ( width + 15 ) >> 4, mb_h = ( height + 15 ) >> 4;
    const int cw = ( width + 1 ) >> 1, ch = ( height + 1 ) >> 1;
    uint64_t sum[3] = { 0, 0, 0 }, ssd[3] = { 0, 0, 0 };
    const int active = aq_mode != 0 && aq_strength != 0;
    float strength = aq_strength * 1.0397f, avg_adj = 0.f, bias_strength = 0.f;
    if( active && aq_mode >= 2 )
    {   /* first pass, ratecontrol.c:352-372 */
        uint64_t dsum[3] = { 0, 0, 0 }, dssd[3] = { 0, 0, 0 };           /* the statistics are accumulated by the second pass */
        float avg_adj_pow2 = 0.f;
        for( int mb_y = 0; mb_y < mb_h; mb_y++ )
            for( int mb_x = 0; mb_x < mb_w; mb_x++ )
            {
                uint32_t energy = energy_block( luma, stride, 16*mb_x, 16*mb_y, 16, 16, width, height, 8, &dsum[0], &dssd[0] );
                energy += energy_block( cb, cstride, 8*mb_x, 8*mb_y, 8, 8, cw, ch, 6, &dsum[1], &dssd[1] );
                energy += energy_block( cr, cstride, 8*mb_x, 8*mb_y, 8, 8, cw, ch, 6, &dsum[2], &dssd[2] );
                float q4 = sqrtf( sqrtf( (float)energy + 1.f ) );         /* (energy+1)^(1/4) = qp_adj * qp_adj under -ffast-math */
                float qp = sqrtf( q4 );                                   /* powf( energy + 1, 0.125f ) */
                qp_offset_aq[mb_x + mb_y * mb_w] = qp;
                avg_adj += qp;
                avg_adj_pow2 += q4;
            }
        avg_adj /= (float)( mb_w * mb_h );
        avg_adj_pow2 /= (float)( mb_w * mb_h );
        strength = avg_adj * aq_strength;
        avg_adj = ( 0.5f * ( 14.f - avg_adj_pow2 ) ) / avg_adj + avg_adj;  /* avg_adj - 0.5f * (avg_adj_pow2 - 14.f) / avg_adj */
        bias_strength = aq_strength;
    }
    for( int mb_y = 0; mb_y < mb_h; mb_y++ )
        for( int mb_x = 0; mb_x < mb_w; mb_x++ )
        {
            uint32_t energy = energy_block( luma, stride, 16*mb_x, 16*mb_y, 16, 16, width, height, 8, &sum[0], &ssd[0] );
            energy += energy_block( cb, cstride, 8*mb_x, 8*mb_y, 8, 8, cw, ch, 6, &sum[1], &ssd[1] );
            energy += energy_block( cr, cstride, 8*mb_x, 8*mb_y, 8, 8, cw, ch, 6, &sum[2], &ssd[2] );
            const int mb = mb_x + mb_y * mb_w;
            float qp_adj = 0.f;
            if( active && aq_mode == 3 )
            {
                float qp = qp_offset_aq[mb];
                qp_adj = ( qp - avg_adj ) * strength + ( 1.f - 14.f / ( qp * qp ) ) * bias_strength;
            }
            else if( active && aq_mode == 2 )
                qp_adj = ( qp_offset_aq[mb] - avg_adj ) * strength;
            else if( active )
            {   /* strength * (x264_log2(energy) - 14.427f) in the association the -ffast-math build of the reference uses
                 * (disassembly of oracle/_ref, ratecontrol.c:397): (integer part - 14.427) + table entry */
                uint32_t e = energy > 1 ? energy : 1;
                float ipart = (float)( 31 - __builtin_clz( e ) );
                float frac = orc_log2_frac( e );
                qp_adj = strength * ( ( ipart - 14.427f ) + frac );
            }
            qp_offset_aq[mb] = qp_adj;
            inv_qscale[mb] = active ? (uint16_t)exp2fix8( qp_adj ) : 256;
        }
    if( stats )
        for( int i = 0; i < 3; i++ )
        {
            uint64_t w = 16 * mb_w >> ( i ? 1 : 0 ), h = 16 * mb_h >> ( i ? 1 : 0 );
            stats[i] = sum[i];
            stats[3+i] = ssd[i] - ( sum[i] * sum[i] + w * h / 2 ) / ( w * h );
        }
}
