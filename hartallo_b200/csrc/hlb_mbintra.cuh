// hlb_mbintra.cuh -- intra mode decision of one macroblock (I pictures, and inside P pictures after the inter search):
//   hl_codec_264_rdo_mb_guess_best_intra_pred_avc            source/h264/hl_codec_264_rdo.c:99-300
//   _hl_codec_264_rdo_mb_guess_best_intra16x16_pred          rdo.c:1526-1812
//   _hl_codec_264_rdo_mb_guess_best_intra4x4_pred            rdo.c:1814-2090
//   _hl_codec_264_rdo_mb_reconstruct_intra16x16_luma         rdo.c:2092-2137, source/h264/hl_codec_264_transf.c:298-373
//   neighbouring samples                                     source/h264/hl_codec_264_pred_intra.c:325-461
//   prev_intra4x4_pred_mode / rem_intra4x4_pred_mode         pred_intra.c:541-614
// Included by hlb_mbcore.cuh (uses MbWork / FrameCtx and the lane-phase execution model described there).
#pragma once

namespace hlb {

// reconstructed luma sample at (x,y) relative to the macroblock (may lie in a neighbouring macroblock), HLB_NA when unavailable
// (6.4.11.1; constrained_intra_pred_flag = 0, so every already-coded neighbour counts)
// Reconstructed samples around the macroblock (left column, corner, top row up to x = 23) are fetched ONCE per macroblock into
// w.nbr_y / w.nbr_c by intra_fetch_borders (one coalesced pass over the lanes); the derivations below then read shared memory only.
// nbr_y: [0] (-1,-1), [1..16] (-1, 0..15), [17..40] (0..23, -1).  nbr_c[c]: [0] (-1,-1), [1..8] (-1, 0..7), [9..16] (0..7, -1).
HLB_HD int intra_luma_fetch(const MbWork& w, const FrameCtx& f, int x, int y)
{
    bool ok;
    if (y < 0 && x >= 0 && x <= 15) ok = w.availB;
    else if (y < 0 && x > 15) ok = w.availC;
    else if (y < 0 && x < 0) ok = w.availD;
    else if (x < 0 && y >= 0 && y <= 15) ok = w.availA;
    else ok = false;
    if (!ok) return HLB_NA;
    return HLB_LDCG(f.cur[0] + (w.mby * 16 + y) * f.W + w.mbx * 16 + x);
}
HLB_HD int intra_chroma_fetch(const MbWork& w, const FrameCtx& f, int c, int x, int y)
{
    bool ok;
    if (y < 0 && x >= 0) ok = w.availB;
    else if (y < 0 && x < 0) ok = w.availD;
    else ok = w.availA;
    if (!ok) return HLB_NA;
    return HLB_LDCG(f.cur[1 + c] + (w.mby * 8 + y) * (f.W >> 1) + w.mbx * 8 + x);
}
HLB_FN void intra_fetch_borders(MbWork& w, const FrameCtx& f, int lane, int nl)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
#pragma unroll 1
    for (int i = lane; i < 41 + 34; i += nl) {
        if (i < 41) w.nbr_y[i] = (i == 0 ? intra_luma_fetch(w, f, -1, -1) : (i < 17 ? intra_luma_fetch(w, f, -1, i - 1) : intra_luma_fetch(w, f, i - 17, -1)));
        else {
            const int k = i - 41, c = k / 17, j = k - c * 17;
            w.nbr_c[c][j] = (j == 0 ? intra_chroma_fetch(w, f, c, -1, -1) : (j < 9 ? intra_chroma_fetch(w, f, c, -1, j - 1) : intra_chroma_fetch(w, f, c, j - 9, -1)));
        }
    }
    HLB_LANE_SYNC();
}
HLB_HD int intra_luma_at(const MbWork& w, const FrameCtx& f, int x, int y)
{
    (void)f;
    if (x >= 0 && x <= 15 && y >= 0 && y <= 15) return w.rec_y[y * 16 + x];
    if (y == -1 && x >= -1 && x <= 23) return x < 0 ? w.nbr_y[0] : w.nbr_y[17 + x];
    if (x == -1 && y >= 0 && y <= 15) return w.nbr_y[1 + y];
    return HLB_NA;
}
HLB_HD int intra_chroma_at(const MbWork& w, const FrameCtx& f, int c, int x, int y)
{
    (void)f;
    if (y < 0) return x < 0 ? w.nbr_c[c][0] : w.nbr_c[c][9 + x];
    return w.nbr_c[c][1 + y];
}

// ------------------------------------------------------------------------------------------------------------------
// CMD_I16_EVAL: 64 lanes = (mode, luma4x4BlkIdx)
// ------------------------------------------------------------------------------------------------------------------
struct I16Shared { I16Params q; };

HLB_FN void i16_phase(MbWork& w, const FrameCtx& f, int phase, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (phase == 0) {   // prediction, residual, transform, AC quantisation (intra offset), CAVLC info of the AC list: packed formulation of hlb_fast.cuh
        if (lane >= 64) return;
        const int m = lane >> 4, blk = lane & 15;
        if (!w.t_mode_ok[m]) return;
        const int bx = blk_x(blk), by = blk_y(blk);
        I16Params q;
        q.dc = w.t_dcbits[0]; q.a = w.t_dcbits[1]; q.b = w.t_dcbits[2]; q.c = w.t_dcbits[3];  // parameters parked by the master
        Rows4 sv, pv;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            pv.r[r] = (uint32_t)i16_pred_px(m, w.p33, q, bx, by + r) | ((uint32_t)i16_pred_px(m, w.p33, q, bx + 1, by + r) << 8) |
                      ((uint32_t)i16_pred_px(m, w.p33, q, bx + 2, by + r) << 16) | ((uint32_t)i16_pred_px(m, w.p33, q, bx + 3, by + r) << 24);
            ((uint32_t*)w.t_pred[m])[((by + r) * 16 + bx) >> 2] = pv.r[r];
            sv.r[r] = ((const uint32_t*)w.src_y)[((by + r) * 16 + bx) >> 2];
        }
        QuantK qi = f.qk;
        qi.f_pos = (1 << qi.qbits) / 3; qi.f_neg = (1 << qi.qbits) - 1 - qi.f_pos;
        int c[16];
        fast_fwd_transform(sv, pv, c);
        w.t_dcw[m][blk] = c[0];
        fast_quant(c, qi);
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) any |= (uint32_t)c[i];   // the block counts as non-zero when ANY of the 16 quantised values is (rdo.c:1640), the DC position included
        // AC list = zig-zag positions 1..15 (0 1 4 8 5 2 3 6 9 12 13 10 7 11 14 15), element 15 = 0
        int l16[16] = {c[1], c[4], c[8], c[5], c[2], c[3], c[6], c[9], c[12], c[13], c[10], c[7], c[11], c[14], c[15], 0};
        uint32_t* acw = (uint32_t*)w.t_ac[m][blk];
#pragma unroll
        for (int i = 0; i < 8; ++i) acw[i] = pack16(l16[2 * i], l16[2 * i + 1]);
        const CavlcInfo ci = cavlc_block_info16(l16, level_mask16(l16));
        w.t_nz[m][blk] = any ? 1 : 0; w.t_tc[m][blk] = ci.total_coeff; w.t_t1[m][blk] = ci.trailing_ones; w.t_sc[m][blk] = ci.single_ctr; w.t_bits[m][blk] = ci.bits_rest;
    } else if (phase == 1) {  // luma DC of each mode: Hadamard (>>1), quantisation, scan (rdo.c:1667-1670); then its de-scaled form for the reconstruction (8.5.10, transf.c:498)
        if (lane >= 4 || !w.t_mode_ok[lane]) return;
        const int m = lane;
        int d[16], lv[16];
#pragma unroll
        for (int blk = 0; blk < 16; ++blk) d[(blk_y(blk) >> 2) * 4 + (blk_x(blk) >> 2)] = w.t_dcw[m][blk];
        hadamard4x4(d);
        const int qb1 = f.qk.qbits + 1, f2 = ((1 << f.qk.qbits) / 3) << 1, mf = f.qk.mf[0];
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int v = d[i] >> 1, z = (iabs(v) * mf + f2) >> qb1; d[i] = v >= 0 ? z : -z; }
        zigzag4x4(d, lv);
#pragma unroll
        for (int i = 0; i < 8; ++i) ((uint32_t*)w.t_dc[m])[i] = pack16(lv[2 * i], lv[2 * i + 1]);
        hadamard4x4(d);   // d still holds the quantised values in raster order
        const int q6 = f.qk.qbits - 15, ls = f.qk.dq_shift ? f.qk.dq_mul[0] : (f.qk.dq_mul[0] >> (q6 - 4));   // LevelScale(QP % 6, 0, 0)
#pragma unroll
        for (int k = 0; k < 16; ++k) w.t_dcw[m][k] = q6 >= 6 ? ((d[k] * ls) << (q6 - 6)) : ((d[k] * ls + (1 << (5 - q6))) >> (6 - q6));   // by raster position of the 4x4 block
    } else {  // reconstruction + distortion per block (rdo.c:1683-1760)
        if (lane >= 64) return;
        const int m = lane >> 4, blk = lane & 15;
        if (!w.t_mode_ok[m]) return;
        const int bx = blk_x(blk), by = blk_y(blk);
        Rows4 sv, pv;
#pragma unroll
        for (int r = 0; r < 4; ++r) { pv.r[r] = ((const uint32_t*)w.t_pred[m])[((by + r) * 16 + bx) >> 2]; sv.r[r] = ((const uint32_t*)w.src_y)[((by + r) * 16 + bx) >> 2]; }
        Rows4 rec = pv;
        if (w.t_cbp[m]) {
            const int dcv = w.t_dcw[m][(by >> 2) * 4 + (bx >> 2)];
            const int16_t* e = w.t_ac[m][blk];
            int c[16] = {dcv, e[0], e[4], e[5], e[1], e[3], e[6], e[11], e[2], e[7], e[10], e[12], e[8], e[9], e[13], e[14]};   // zig-zag list back to raster positions
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) any |= (uint32_t)c[i];
            if (any) {
                fast_dequant_inverse(c, f.qk, /*keep_dc*/ true);
                rec = fast_recon_clip(pv, c);
#pragma unroll
                for (int r = 0; r < 4; ++r) ((uint32_t*)w.t_pred[m])[((by + r) * 16 + bx) >> 2] = rec.r[r];   // becomes the reconstruction of mode m
            }
        }
        uint32_t dist = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) dist = p_sad4(sv.r[r], rec.r[r], dist);
        w.t_dist[m][blk] = (int)dist;
    }
}
HLB_FN void i16_recon_phase(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    (void)f;
    if (lane >= 16) return;  // copy the reconstruction of the chosen Intra16x16 mode (identical to transf.c:298 on the same levels)
    for (int i = 0; i < 16; ++i) w.rec_y[lane * 16 + i] = w.t_pred[w.i16_mode][lane * 16 + i];
}

// ------------------------------------------------------------------------------------------------------------------
// CMD_I4_EVAL for block w.i4_blk: 9 lanes = Intra4x4 modes (rdo.c:1903-2010)
// ------------------------------------------------------------------------------------------------------------------
HLB_FN void i4_phase(MbWork& w, const FrameCtx& f, int phase, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (phase != 0 || lane >= 9) return;
    const int mode = lane, blk = w.i4_blk, bx = blk_x(blk), by = blk_y(blk);
    w.q_ok[mode] = i4_mode_allowed(mode, w.p13) ? 1 : 0;
    if (!w.q_ok[mode]) return;
    int pred[16];
    intra4x4_pred(mode, w.p13, pred);
    // packed from here on (hlb_fast.cuh): the block as four words, the intra rounding offset (1/3) in the quantiser
    Rows4 sv, pv;
    bool rnz = false;
    uint32_t* qp4 = (uint32_t*)w.q_pred[mode];
    uint32_t* lvw = (uint32_t*)w.q_lv[mode];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        pv.r[r] = (uint32_t)pred[r * 4] | ((uint32_t)pred[r * 4 + 1] << 8) | ((uint32_t)pred[r * 4 + 2] << 16) | ((uint32_t)pred[r * 4 + 3] << 24);
        sv.r[r] = ((const uint32_t*)w.src_y)[((by + r) * 16 + bx) >> 2];
        rnz |= sv.r[r] != pv.r[r];
        qp4[r] = pv.r[r];
    }
    w.q_res0[mode] = rnz ? 0 : 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) lvw[i] = 0;
    w.q_nz[mode] = 0; w.q_bits[mode] = 0; w.q_tc[mode] = 0; w.q_t1[mode] = 0; w.q_sc[mode] = 9; w.q_dist[mode] = 0;
    if (!rnz) return;
    QuantK qi = f.qk;
    qi.f_pos = (1 << qi.qbits) / 3; qi.f_neg = (1 << qi.qbits) - 1 - qi.f_pos;
    int m[16];
    fast_fwd_transform(sv, pv, m);
    fast_quant(m, qi);
    int lv[16];
    zigzag4x4(m, lv);
    const uint32_t mask = level_mask16(lv);
    Rows4 rec = pv;
    if (mask) {
#pragma unroll
        for (int i = 0; i < 8; ++i) lvw[i] = pack16(lv[2 * i], lv[2 * i + 1]);
        const CavlcInfo ci = cavlc_block_info16(lv, mask);
        w.q_nz[mode] = 1; w.q_bits[mode] = ci.bits_rest; w.q_tc[mode] = ci.total_coeff; w.q_t1[mode] = ci.trailing_ones; w.q_sc[mode] = ci.single_ctr;
        fast_dequant_inverse(m, qi, false);
        rec = fast_recon_clip(pv, m);
#pragma unroll
        for (int r = 0; r < 4; ++r) qp4[r] = rec.r[r];   // reconstruction of this mode (the prediction when nothing is coded)
    }
    uint32_t dist = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) dist = p_sad4(sv.r[r], rec.r[r], dist);
    w.q_dist[mode] = (int)dist;
}
HLB_HD void i4_commit_phase(MbWork& w, const FrameCtx& f, int lane) { (void)w; (void)f; (void)lane; }

// ------------------------------------------------------------------------------------------------------------------
// CMD_PRED_CHROMA_INTRA: 8 lanes = (plane, 4x4 block); p17 and the mode are set by the master (pred_intra.c:1044-1230)
// ------------------------------------------------------------------------------------------------------------------
HLB_FN void intra_chroma_pred_phase(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    (void)f;
    if (lane >= 8) return;
    const int c = lane >> 2, b = lane & 3, x0 = (b & 1) * 4, y0 = (b >> 1) * 4;
    const ICParams q = ic_params(w.p17[c]);
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 4; ++k) w.pred_c[c][(y0 + r) * 8 + x0 + k] = (uint8_t)ic_pred_px(w.intra_chroma_mode, w.p17[c], q, x0 + k, y0 + r);
}

// ------------------------------------------------------------------------------------------------------------------
// master: rdo.c:99-300.  Returns MBK_I16 / MBK_I4 and the best intra cost; leaves the reconstruction in w.rec_y / w.rec_c and
// in the picture, levels / modes / CBP flags in w.
// ------------------------------------------------------------------------------------------------------------------
template <class X>
HLB_FN int mb_encode_intra(X& x, MbWork& w, const FrameCtx& f, double& intra_cost)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    intra_fetch_borders(w, f, x.lane(), x.nlanes());
    // ---- Intra16x16 (no reconstruction into the picture yet) ----
    // the serial copies of this function are strided over the lanes of the master warp (x.lane(), x.nlanes(); one lane on the CPU harness)
#pragma unroll 1
    for (int i = x.lane(); i < 33; i += x.nlanes()) w.p33[i] = i == 0 ? intra_luma_at(w, f, -1, -1) : (i < 17 ? intra_luma_at(w, f, -1, i - 1) : intra_luma_at(w, f, i - 17, -1));
    x.sync();
#pragma unroll 1
    for (int m = 0; m < 4; ++m) w.t_mode_ok[m] = i16_mode_allowed(m, w.p33) ? 1 : 0;
    {
        const I16Params q = i16_params(w.p33);
        w.t_dcbits[0] = q.dc; w.t_dcbits[1] = q.a; w.t_dcbits[2] = q.b; w.t_dcbits[3] = q.c;
    }
    x.run(CMD_I16_EVAL, 64);
    // The Single_ctr of a block whose AC levels are all zero is whatever the previous residual_block call left behind
    // (residual.c:882 only runs when TotalCoeffs > 0).  Fetch the chain value of the raster predecessor only when this
    // macroblock has not produced one itself and the first flagged block would consume it.
    w.arg1 = 0;
    if (w.last_sctr < 0) {
        bool need = false, done = false;
#pragma unroll 1
        for (int m = 0; m < 4 && !done; ++m) {
            if (!w.t_mode_ok[m]) continue;
#pragma unroll 1
            for (int b = 0; b < 16 && !done; ++b)
                if (w.t_nz[m][b]) { need = (w.t_tc[m][b] == 0); done = true; }
        }
        if (need) w.arg1 = x.prev_sctr(w.mb);
    }
    // Rate with the evolving nC state, single-coefficient elimination, DC token (rdo.c:1617-1679).  The four modes follow each other (a mode leaves its TotalCoeff
    // behind for the next one); inside a mode the blocks go over the lanes: a block's nC reads its left / upper neighbours, which precede it, so "the state when
    // block b is reached" is the state after every non-zero block of the mode has been written -- written first, read after one barrier.
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
        if (!w.t_mode_ok[m]) continue;
#pragma unroll 1
        for (int blk = x.lane(); blk < 16; blk += x.nlanes())
            if (w.t_nz[m][blk]) w.tc[blk] = w.t_tc[m][blk];
        x.sync();
        int bits = 0, cbp = 0;
#pragma unroll 1
        for (int blk = x.lane(); blk < 16; blk += x.nlanes())
            if (w.t_nz[m][blk]) { bits += w.t_bits[m][blk] + coeff_token_len(luma_nc(w, w.tc, blk), w.t_tc[m][blk], w.t_t1[m][blk]); cbp |= 1 << blk; }
        bits = x.reduce_add(bits); cbp = x.reduce_add(cbp);
        // the Single_ctr chain is a scan over the blocks in order: every lane walks it (16 short steps on shared memory)
        int sctr = 0, last = w.last_sctr, need_prev = 0;
#pragma unroll 1
        for (int blk = 0; blk < 16; ++blk) {
            if (!w.t_nz[m][blk]) continue;
            if (w.t_tc[m][blk] > 0) last = w.t_sc[m][blk];
            else if (last < 0) { need_prev = 1; last = w.arg1; }   // chain value of the raster predecessor (residual.c:882 is skipped when TotalCoeffs == 0)
            sctr += last;
        }
        if (cbp && sctr < 6) cbp = 0;
        int tc0 = -1;
        if (cbp) {
            int l16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) l16[i] = w.t_dc[m][i];
            const CavlcInfo ci = cavlc_block_info16(l16, level_mask16(l16));
            bits += ci.bits_rest + coeff_token_len(luma_nc(w, w.tc, 0), ci.total_coeff, ci.trailing_ones);
            tc0 = ci.total_coeff;
            if (ci.total_coeff > 0) last = ci.single_ctr;
        }
        x.sync();
        if (x.lane() == 0) {
            w.last_sctr = last;
            if (need_prev) w.need_prev_sctr = 1;
            if (tc0 >= 0) w.tc[0] = (uint8_t)tc0;
            w.t_cbp[m] = cbp; w.t_rate[m] = bits;
        }
        x.sync();
    }
    x.run(CMD_I16_RATE, 64);   // reconstruction + distortion of every (mode, block)
#pragma unroll 1
    for (int m = 0; m < 4; ++m) w.stat_intra += w.t_mode_ok[m] ? 16u : 0u;
    double best16 = DBL_MAX;
    int dist16 = 0;
    w.i16_mode = 2; w.i16_cbp4x4 = 0;
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
        if (!w.t_mode_ok[m]) continue;
        double d = 0;
#pragma unroll 1
        for (int b = 0; b < 16; ++b) d += w.t_dist[m][b];
        const double cost = d + (f.lambda * (double)w.t_rate[m]);
        HLB_DBG("  I16 mode %d: dist %.0f rate %d cbp %x cost %.4f\n", m, d, w.t_rate[m], w.t_cbp[m], cost);
        if (cost < best16) { best16 = cost; dist16 = (int)d; w.i16_mode = m; w.i16_cbp4x4 = w.t_cbp[m]; }
    }
    x.sync();   // i16_mode / i16_cbp4x4 are written by every lane with the same value
#pragma unroll 1
    for (int i = x.lane(); i < 256; i += x.nlanes()) (&w.i16_ac[0][0])[i] = (&w.t_ac[w.i16_mode][0][0])[i];
#pragma unroll 1
    for (int i = x.lane(); i < 16; i += x.nlanes()) w.i16_dc[i] = w.i16_cbp4x4 ? w.t_dc[w.i16_mode][i] : 0;
    x.sync();

    // ---- Intra4x4 (reconstructs block after block: later blocks predict from it) ----
    double cost4 = DBL_MAX;
    int dist4 = INT_MAX, cbp4 = 0;
    if (best16 != 0) {
        cost4 = 0; dist4 = 0;
#pragma unroll 1
        for (int blk = 0; blk < 16; ++blk) {
            const int bx = blk_x(blk), by = blk_y(blk);
            // the 13 neighbouring samples, one per lane: [0] corner, [1..4] left column, [5..12] top row and top-right (pred_intra.c:325-461)
#pragma unroll 1
            for (int i = x.lane(); i < 13; i += x.nlanes())
                w.p13[i] = i == 0 ? intra_luma_at(w, f, bx - 1, by - 1) : (i < 5 ? intra_luma_at(w, f, bx - 1, by + i - 1) : ((i > 8 && (blk == 3 || blk == 11)) ? HLB_NA : intra_luma_at(w, f, bx + i - 5, by - 1)));
            x.sync();
            // in-MB positions to the right of an uncoded area (blocks 5, 7, 13, 15) fall outside the macroblock => HLB_NA from intra_luma_at; 8.3.1.2: substituted by p[3,-1]
            if (w.p13[9] == HLB_NA && w.p13[10] == HLB_NA && w.p13[11] == HLB_NA && w.p13[12] == HLB_NA && w.p13[8] != HLB_NA) {
                x.sync();
                if (x.lane() == 0) w.p13[9] = w.p13[10] = w.p13[11] = w.p13[12] = w.p13[8];
                x.sync();
            }
            if (x.lane() == 0) { w.i4_blk = blk; w.i4_mode[blk] = 2; }
            x.sync();
            x.run(CMD_I4_EVAL, 9);
            // The reference walks the nine modes in order (rdo.c:1903-2010): a mode whose prediction equals the source wins outright and ends the walk; every
            // earlier mode that coded something leaves its TotalCoeff / Single_ctr behind (the last one stands); the cheapest mode wins, the first among equals.
            // One mode per lane, three reductions.
            const int nC = luma_nc(w, w.tc, blk);
            int first_res0 = 99, last_nz = -1, n_ok = 0;
#pragma unroll 1
            for (int mode = x.lane(); mode < 9; mode += x.nlanes()) {
                if (!w.q_ok[mode]) continue;
                ++n_ok;
                if (w.q_res0[mode] && mode < first_res0) first_res0 = mode;
            }
            first_res0 = x.reduce_min(first_res0);
            n_ok = x.reduce_add(n_ok);
            double mc = DBL_MAX;
            int mi = 99;
#pragma unroll 1
            for (int mode = x.lane(); mode < 9; mode += x.nlanes()) {
                if (!w.q_ok[mode] || mode >= first_res0) continue;
                int bits = 0;
                if (w.q_nz[mode]) { bits = w.q_bits[mode] + coeff_token_len(nC, w.q_tc[mode], w.q_t1[mode]); if (mode > last_nz) last_nz = mode; }
                const double cost = (double)w.q_dist[mode] + (f.lambda * (double)bits);
                if (cost < mc) { mc = cost; mi = mode; }
            }
            last_nz = x.reduce_max(last_nz);
            x.reduce_argmin(mc, mi);
            double min_cost, min_dist;
            int best_mode, best_allzero;
            if (first_res0 < 99) { min_cost = 0; min_dist = 0; best_mode = first_res0; best_allzero = 1; }
            else { min_cost = mc; min_dist = w.q_dist[mi]; best_mode = mi; best_allzero = !w.q_nz[mi]; }
            x.sync();
            if (x.lane() == 0) {
                w.stat_intra += (unsigned)n_ok;
                if (last_nz >= 0) { w.tc[blk] = w.q_tc[last_nz]; w.last_sctr = w.q_sc[last_nz]; }
                w.i4_mode[blk] = (uint8_t)best_mode;
            }
            HLB_DBG("  I4 blk %d: mode %d cost %.4f dist %.0f nC %d\n", blk, best_mode, min_cost, min_dist, nC);
            x.sync();
#pragma unroll 1
            for (int i = x.lane(); i < 16; i += x.nlanes()) {
                w.luma_level[blk][i] = w.q_lv[best_mode][i];
                w.rec_y[(by + (i >> 2)) * 16 + bx + (i & 3)] = w.q_pred[best_mode][i];
            }
            cost4 += min_cost; dist4 = (int)(dist4 + min_dist);
            if (!best_allzero) cbp4 |= 1 << blk;
            x.sync();   // the next block predicts from these samples
        }
    }
    w.i4_cbp4x4 = cbp4;
    // chroma prediction mode follows the Intra16x16 mode (rdo.c:165-179)
    w.intra_chroma_mode = w.i16_mode == 0 ? 2 : (w.i16_mode == 3 ? 3 : (w.i16_mode == 1 ? 1 : 0));
    int kind = MBK_I16, mad = dist16;
    if (cost4 < best16) {
        kind = MBK_I4;
        w.cbp_luma4x4 = cbp4;
        int zeros = 0;
#pragma unroll 1
        for (int blk = 0; blk < 16; ++blk) {  // pred_intra.c:541-614
            const int bx = blk_x(blk), by = blk_y(blk);
            int mA = -1, mB = -1;  // -1: neighbour macroblock not available
            if (bx > 0) mA = w.i4_mode[blk_idx_from_xy(bx - 4, by)];
            else if (w.availA) { const MbState& s = *(const MbState*)w.nbw[1]; mA = s.kind == MBK_I4 ? s.i4_mode[blk_idx_from_xy(12, by)] : 2; }
            if (by > 0) mB = w.i4_mode[blk_idx_from_xy(bx, by - 4)];
            else if (w.availB) { const MbState& s = *(const MbState*)w.nbw[2]; mB = s.kind == MBK_I4 ? s.i4_mode[blk_idx_from_xy(bx, 12)] : 2; }
            if (mA < 0 || mB < 0) mA = mB = 2;
            const int pm = mA < mB ? mA : mB, cur = w.i4_mode[blk];
            if (pm == cur) { w.prev_i4[blk] = 1; }
            else { w.prev_i4[blk] = 0; w.rem_i4[blk] = (uint8_t)(cur < pm ? cur : cur - 1); ++zeros; }
        }
        cost4 += f.lambda * (double)(16 + zeros * 3);
        mad = dist4;
    }
    HLB_DBG("  intra: best16 %.4f cost4 %.4f\n", best16, cost4);
    if (best16 <= cost4) { kind = MBK_I16; w.cbp_luma4x4 = w.i16_cbp4x4; mad = dist16; }
    intra_cost = best16 < cost4 ? best16 : cost4;
    w.arg0 = mad;  // parked for mb_commit_intra
    // ---- chroma (rdo.c:216-246) ----
    w.mb_is_intra = 1;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        w.p17[c][0] = intra_chroma_at(w, f, c, -1, -1);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) { w.p17[c][1 + i] = intra_chroma_at(w, f, c, -1, i); w.p17[c][9 + i] = intra_chroma_at(w, f, c, i, -1); }
    }
    const int mad_keep = w.arg0;
    x.run(CMD_PRED_CHROMA_INTRA, 8);
    x.run(CMD_CHROMA, 8);
    if (kind == MBK_I16) x.run(CMD_I16_RECON, 16);
    w.arg0 = 3; x.run(CMD_STORE, 96);
    w.arg0 = mad_keep;
    return kind;
}

HLB_FN void mb_commit(MbWork& w, const FrameCtx& f, int kind, int cbp_luma, int cbp_chroma, int coded_block_pattern, int mb_type, const int16_t mvd[4][4][2], int mad, int lane, int nl);
HLB_HD int guess_cbp_luma(int cbp4x4, bool i16);
HLB_HD int guess_cbp_chroma(const MbWork& w);

HLB_FN void mb_commit_intra(MbWork& w, const FrameCtx& f, int kind, int lane, int nl)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int mad = w.arg0;
    const int cbp_luma = guess_cbp_luma(w.cbp_luma4x4, kind == MBK_I16);
    int cbp_chroma = guess_cbp_chroma(w);
    int cbp = (cbp_chroma << 4) | cbp_luma;
    if (cbp > 47) { cbp -= 16; cbp_chroma = cbp >> 4; }
    int mb_type = kind == MBK_I4 ? 0 : 1 + (cbp_chroma << 2) + w.i16_mode + (cbp_luma ? 12 : 0);
    if (f.is_p) mb_type += 5;
    w.fin_mode = 0; w.fin_sub[0] = w.fin_sub[1] = w.fin_sub[2] = w.fin_sub[3] = 0;
    for (int p = 0; p < 4; ++p) {
        w.fin_ref[p] = 0;
        for (int s = 0; s < 4; ++s) w.fin_mv[p][s][0] = w.fin_mv[p][s][1] = 0;
    }
    mb_commit(w, f, kind, cbp_luma, cbp_chroma, cbp, mb_type, nullptr, mad, lane, nl);
}

}  // namespace hlb
