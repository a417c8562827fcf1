// hlb_svc.cuh -- SVC enhancement-layer inter macroblock (base_mode_flag = 1), SURVEY 8a row a14:
// hl_codec_264_rdo_mb_guess_best_inter_pred_svc, source/h264/hl_codec_264_rdo.c:1273-1521.
//
// The reference does no search for these macroblocks: partitions and motion vectors come from the inter-layer derivation
// (utils.c:966-2439, host side), the macroblock is predicted from RefPicList0[0] of its own layer
// (rdo.c:1340-1424), the luma residual is transformed and quantised with the INTRA rounding offset (rdo.c:1468), chroma goes
// through the shared _hl_codec_264_rdo_mb_reconstruct_chroma (rdo.c:1500, :2502-2700) with the macroblock counted as inter.
// Prediction, residual coding and reconstruction are fused here: one lane owns one 4x4 block (16 luma + 2x4 chroma lanes per
// macroblock) from the reference window to the reconstructed samples, so the prediction never travels through HBM
// (1,152 B per macroblock = 384 reference + 384 source read, 384 reconstruction written, plus 768 B of levels).
//
// Two phases per lane with one exchange area between them (the 2x2 chroma DC stage and the single-coefficient elimination
// look at all four blocks of a plane).  Only the prediction differs between luma and chroma lanes; residual, transform,
// quantisation, scan (phase A) and de-quantisation, inverse transform, add/clip, stores (phase B) are the same instructions
// for all 24 lanes, so the warp executes them once (converged) instead of once per kind of lane.  The phases are `HLB_HD`: under nvcc they are the body of k_svc_inter_recon
// (hlb_batch.cu), compiled as plain C++ they run lane by lane in tools/emu/svc_emu.cpp -- the CPU tier checks the same source
// against the reference's per-macroblock trace.
#pragma once
#include "../../include/hlb200.h"
#include "hlb_prims.cuh"
#include "hlb_fast.cuh"   // packed-byte primitives (funnel shift, byte dot product, saturating pack) for svc_resample_row4

namespace hlb {

// partition that contains luma position (bx, by); same geometry as part_of() in hlb_common.cuh (kept here so this header
// compiles without the CUDA runtime)
struct SvcPart { int part, sub, ox, oy; };
HLB_HD SvcPart svc_part_of(int part_mode, const uint8_t sub_mode[4], int bx, int by)
{
    SvcPart g;
    g.sub = 0;
    if (part_mode == 0) { g.part = 0; g.ox = 0; g.oy = 0; }
    else if (part_mode == 1) { g.part = by >> 3; g.ox = 0; g.oy = g.part * 8; }
    else if (part_mode == 2) { g.part = bx >> 3; g.ox = g.part * 8; g.oy = 0; }
    else {
        g.part = ((by >> 3) << 1) | (bx >> 3);
        const int px = (g.part & 1) * 8, py = (g.part >> 1) * 8, lx = bx & 7, ly = by & 7, sm = sub_mode[g.part];
        if (sm == 0) { g.ox = px; g.oy = py; }
        else if (sm == 1) { g.sub = ly >> 2; g.ox = px; g.oy = py + g.sub * 4; }
        else if (sm == 2) { g.sub = lx >> 2; g.ox = px + g.sub * 4; g.oy = py; }
        else { g.sub = ((ly >> 2) << 1) | (lx >> 2); g.ox = px + (g.sub & 1) * 4; g.oy = py + (g.sub >> 1) * 4; }
    }
    return g;
}

// what the four block lanes of a chroma plane tell each other (shared memory on the device, one per macroblock)
struct SvcXchg {
    int32_t dc_coef[2][4];   // W00 of the block before quantisation (rdo.c:2591)
    uint8_t nnz[2][4];       // non-zero AC levels of a block whose AC bit is set
    uint8_t big[2][4];       // any |level| > 1 among them
    uint8_t coded[2][4];     // CodedBlockPatternChromaAC4x4 bit before the elimination
    uint8_t luma_coded[16];  // CodedBlockPatternLuma4x4 bits
};
// what a lane keeps between its two phases (registers on the device)
struct SvcLane {
    uint8_t pv[16];   // prediction of the lane's 4x4 block
    int16_t lv[16];   // luma lane: LumaLevel[blk]; chroma lane: ChromaACLevel[plane][blk] as the macroblock object holds it after the forward pass
                      // (15 AC levels + the never-written [15]; stale when the residual is zero)
    bool coded;       // luma lane: its CodedBlockPatternLuma4x4 bit
};

// four samples of a row as one word (plane bases are 4-byte aligned, widths multiples of 16)
HLB_HD uint32_t svc_ld4(const uint8_t* p) { return HLB_LDG(reinterpret_cast<const uint32_t*>(p)); }
HLB_HD void svc_st4(uint8_t* p, int a, int b, int c, int d) { *reinterpret_cast<uint32_t*>(p) = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24); }

// bytes sh8/8 .. sh8/8+3 of the eight bytes {lo, hi} (little endian); sh8 in {0, 8, 16, 24}
HLB_HD uint32_t svc_funnel(uint32_t lo, uint32_t hi, int sh8)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh8);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> sh8);
#endif
}

struct SvcPlanes {
    const uint8_t *src_y, *src_u, *src_v;
    const uint8_t *ref_y, *ref_u, *ref_v;   // inter (base mode): reference picture of the layer, or
                                            // I_BL: the PREDICTION planes (base-layer reconstruction resampled by the host, G.8.6.2)
    uint8_t *rec_y, *rec_u, *rec_v;
    int W, H;
};

// ---- luma lane (blk = luma4x4BlkIdx): prediction -> residual -> T -> Q (intra offset) -> Q^-1 -> T^-1 -> reconstruction (rdo.c:1428-1496) ----
HLB_HD void svc_luma_predict(const SvcPlanes& P, int mbx, int mby, int blk, const hlb200_mb_motion_t& m, uint8_t pv[16])
{
    const int bx = blk_x(blk), by = blk_y(blk), W = P.W, H = P.H;
    const SvcPart g = svc_part_of(m.part_mode, m.sub_mode, bx, by);
    const int mvx = m.mv[g.part][g.sub][0], mvy = m.mv[g.part][g.sub][1];
    // the origin clip applies to the PARTITION origin (pred_inter.c:395-396); samples are fetched with the per-sample
    // clamp of the reference's index table (interpol.c:108-131)
    const int X0 = clip3(-17, W + 17, mbx * 16 + g.ox + (mvx >> 2)) + (bx - g.ox);
    const int Y0 = clip3(-17, H + 17, mby * 16 + g.oy + (mvy >> 2)) + (by - g.oy);
    // 9 rows x 12 bytes, kept as words.  Inside the picture a row is fetched as four ALIGNED 32-bit words and shifted into place (the r01d profile of
    // k_interp_luma showed the byte-wise window fetch L1-bound: 81 byte loads per block); the words of a row never leave that row of the plane.
    uint32_t win[27];
    const int xa = (X0 - 2) & ~3, sh8 = ((X0 - 2) & 3) * 8;
    if (X0 >= 2 && Y0 >= 2 && xa + 16 <= W && Y0 + 7 <= H) {
        const uint8_t* p = P.ref_y + (Y0 - 2) * W + xa;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(p + r * W);
            const uint32_t w0 = HLB_LDG(q), w1 = HLB_LDG(q + 1), w2 = HLB_LDG(q + 2), w3 = HLB_LDG(q + 3);
            win[r * 3] = svc_funnel(w0, w1, sh8); win[r * 3 + 1] = svc_funnel(w1, w2, sh8); win[r * 3 + 2] = svc_funnel(w2, w3, sh8);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint8_t* row = P.ref_y + clip3(0, H - 1, Y0 - 2 + r) * W;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) w |= (uint32_t)HLB_LDG(row + clip3(0, W - 1, X0 - 2 + 4 * k + j)) << (8 * j);
                win[r * 3 + k] = w;
            }
        }
    }
    interp_luma_4x4(reinterpret_cast<const uint8_t*>(win) + 2 * 12 + 2, 12, mvx & 3, mvy & 3, pv);
}
// I_BL: the prediction is a plane (rdo.c:363, mbPredL)
HLB_HD void svc_load_pred4x4(const uint8_t* plane, int off, int pitch, uint8_t pv[16])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t w = svc_ld4(plane + off + r * pitch);
        pv[r * 4] = (uint8_t)w; pv[r * 4 + 1] = (uint8_t)(w >> 8); pv[r * 4 + 2] = (uint8_t)(w >> 16); pv[r * 4 + 3] = (uint8_t)(w >> 24);
    }
}
// Where a macroblock's prediction comes from.  Normally itself.  A macroblock of a P picture whose base macroblock is intra reaches the reference's inter
// function without any partition (predFlagL0 = 0, NumSubMbPart = 0): the prediction loops (rdo.c:1350-1424) do not run and predMbL/Cb/Cr -- scratch blocks
// that the allocator hands out at the same position on every call (hl_memory.h:226-241) -- still hold the prediction of the last macroblock that had
// partitions.  The macroblock is then coded against THAT prediction, and because the SVC initialisation process flags it intra, with the intra offset in
// the 2x2 chroma DC quantisation (rdo.c:2660).  The host marks such a macroblock in hlb200_mb_motion_t::pad: pad[0] bit 0, pad[1..2] = address of the
// macroblock whose prediction it inherits (traced: tag 10 of oracle/ref_driver.c; tests/test_svc_inter.py::test_stale_prediction_model).
struct SvcPredSrc { int mbx, mby; const hlb200_mb_motion_t* m; bool inherited; };
HLB_HD SvcPredSrc svc_pred_src(const hlb200_mb_motion_t* pic_motion, int mb, int mbw, int nmb = 0x7fffffff)
{
    SvcPredSrc s;
    const hlb200_mb_motion_t* m = pic_motion + mb;
    s.inherited = (m->pad[0] & 1) != 0;
    int from = s.inherited ? ((int)m->pad[1] | ((int)m->pad[2] << 8)) : mb;
    if (from >= nmb) { from = mb; s.inherited = false; }   // a 16-bit address outside the picture: caller error, never read out of bounds
    s.m = pic_motion + from; s.mbx = from % mbw; s.mby = from / mbw;
    return s;
}

// chroma prediction of one 4x4 block (plane 0 = Cb, 1 = Cr; blk raster 0..3)
HLB_HD void svc_chroma_predict(const SvcPlanes& P, int mbx, int mby, int plane, int blk, const hlb200_mb_motion_t& m, SvcLane& L)
{
    const int Wc = P.W >> 1, Hc = P.H >> 1, bx = (blk & 1) * 4, by = (blk >> 1) * 4;
    const uint8_t* ref = plane ? P.ref_v : P.ref_u;
    // A 4x4 chroma block is an 8x8 luma area: one motion vector unless the macroblock is 8x8-partitioned with sub-partitions there (never after the dyadic
    // inter-layer derivation).  Then, inside the picture, the 5x5 window is fetched as two aligned words per row and shifted into place.
    if (m.part_mode != 3 || m.sub_mode[blk] == 0) {
        const int part = m.part_mode == 0 ? 0 : (m.part_mode == 1 ? (blk >> 1) : (m.part_mode == 2 ? (blk & 1) : blk));
        const int mvx = m.mv[part][0][0], mvy = m.mv[part][0][1], xf = mvx & 7, yf = mvy & 7;
        const int x0 = mbx * 8 + bx + (mvx >> 3), y0 = mby * 8 + by + (mvy >> 3), xa = x0 & ~3, sh8 = (x0 & 3) * 8;
        if (x0 >= 0 && y0 >= 0 && xa + 8 <= Wc && y0 + 5 <= Hc) {
            uint32_t lo[5], hi[5];   // samples x0..x0+3 and x0+4 of rows y0..y0+4
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const uint32_t* q = reinterpret_cast<const uint32_t*>(ref + (y0 + r) * Wc + xa);
                const uint32_t w0 = HLB_LDG(q), w1 = HLB_LDG(q + 1);
                lo[r] = svc_funnel(w0, w1, sh8); hi[r] = (w1 >> sh8) & 255u;
            }
#pragma unroll
            for (int y = 0; y < 4; ++y)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int A = (lo[y] >> (8 * x)) & 255u, C = (lo[y + 1] >> (8 * x)) & 255u;
                    const int B = x < 3 ? (int)((lo[y] >> (8 * x + 8)) & 255u) : (int)hi[y], D = x < 3 ? (int)((lo[y + 1] >> (8 * x + 8)) & 255u) : (int)hi[y + 1];
                    L.pv[y * 4 + x] = (uint8_t)interp_chroma_px(A, B, C, D, xf, yf);
                }
            return;
        }
    }
    // 8.4.2.2.2 per sample with the reference's clamp; a 2x2 chroma area is the smallest one with its own motion vector
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; x += 2) {
            const SvcPart g = svc_part_of(m.part_mode, m.sub_mode, (bx + x) * 2, (by + y) * 2);
            const int mvx = m.mv[g.part][g.sub][0], mvy = m.mv[g.part][g.sub][1];
            const int x0 = mbx * 8 + bx + x + (mvx >> 3), y0 = mby * 8 + by + y + (mvy >> 3), xf = mvx & 7, yf = mvy & 7;
            const int xa = clip3(0, Wc - 1, x0), xb = clip3(0, Wc - 1, x0 + 1), xc = clip3(0, Wc - 1, x0 + 2);
            const int ya = clip3(0, Hc - 1, y0) * Wc, yc = clip3(0, Hc - 1, y0 + 1) * Wc;
            const int a0 = HLB_LDG(ref + ya + xa), a1 = HLB_LDG(ref + ya + xb), a2 = HLB_LDG(ref + ya + xc);
            const int c0 = HLB_LDG(ref + yc + xa), c1 = HLB_LDG(ref + yc + xb), c2 = HLB_LDG(ref + yc + xc);
            L.pv[y * 4 + x] = (uint8_t)interp_chroma_px(a0, a1, c0, c1, xf, yf);
            L.pv[y * 4 + x + 1] = (uint8_t)interp_chroma_px(a1, a2, c1, c2, xf, yf);
        }
}
// ---- phase A, lanes 0..23 (0..15: luma4x4BlkIdx; 16..19: Cb blocks; 20..23: Cr blocks): prediction, then residual -> T -> Q (intra rounding offset for
// luma, rdo.c:1468, and always for chroma AC, rdo.c:2588) -> scan (rdo.c:1428-1466, :2561-2638) ----
// ps: motion and position the prediction is formed with (the macroblock itself, or the one it inherits from); unused for BL
template <bool BL>
HLB_HD void svc_lane_a(const SvcPlanes& P, int mbx, int mby, int lane, const SvcPredSrc& ps, int qp, int qpc, const hlb200_svc_mb_state_t& st, SvcLane& L, SvcXchg& X)
{
    const bool is_luma = lane < 16;
    const int plane = (lane - 16) >> 2, cblk = (lane - 16) & 3;
    const int pitch = is_luma ? P.W : (P.W >> 1);
    const int bx = is_luma ? blk_x(lane) : (cblk & 1) * 4, by = is_luma ? blk_y(lane) : (cblk >> 1) * 4;
    const int off = is_luma ? (mby * 16 + by) * pitch + mbx * 16 + bx : (mby * 8 + by) * pitch + mbx * 8 + bx;
    const uint8_t* src = is_luma ? P.src_y : (plane ? P.src_v : P.src_u);
    if (BL) svc_load_pred4x4(is_luma ? P.ref_y : (plane ? P.ref_v : P.ref_u), off, pitch, L.pv);
    else if (is_luma) svc_luma_predict(P, ps.mbx, ps.mby, lane, *ps.m, L.pv);
    else svc_chroma_predict(P, ps.mbx, ps.mby, plane, cblk, *ps.m, L);
    int mm[16];
    bool nz = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t sw = svc_ld4(src + off + r * pitch);
#pragma unroll
        for (int c = 0; c < 4; ++c) { const int d = (int)((sw >> (8 * c)) & 255u) - (int)L.pv[r * 4 + c]; mm[r * 4 + c] = d; nz |= (d != 0); }
    }
    // ChromaACLevel lives in the macroblock object from picture to picture: a block whose residual is zero keeps what an earlier picture left there, and
    // transf.c:236-245 reads it again whenever the block's de-quantised DC is not zero.  Element [15] is never written.  LumaLevel is cleared (rdo.c:1453,1462).
#pragma unroll
    for (int i = 0; i < 16; ++i) L.lv[i] = is_luma ? (int16_t)0 : st.chroma_ac_level[is_luma ? 0 : plane][cblk][i];
    int dc = 0;
    if (nz) {
        int z[16];
        fwd_transform4x4(mm);
        dc = mm[0];   // chroma: W00 before quantisation, input of the 2x2 DC stage (rdo.c:2591)
        quant4x4_ac(mm, is_luma ? qp : qpc, /*intra f*/ true);
        zigzag4x4(mm, z);
        if (is_luma) {
#pragma unroll
            for (int i = 0; i < 16; ++i) L.lv[i] = (int16_t)z[i];
        } else {
#pragma unroll
            for (int i = 1; i < 16; ++i) L.lv[i - 1] = (int16_t)z[i];   // Scan4x4_AC_C, utils.h:183
        }
    }
    bool coded = false;
    int nnz = 0, big = 0;
    if (nz) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { coded |= (L.lv[i] != 0); nnz += (L.lv[i] != 0); big |= (iabs(L.lv[i]) > 1); }
    }
    L.coded = coded;
    if (is_luma) X.luma_coded[lane] = coded ? 1 : 0;
    else {
        X.dc_coef[plane][cblk] = dc;
        X.coded[plane][cblk] = coded ? 1 : 0;
        X.nnz[plane][cblk] = (uint8_t)(coded ? nnz : 0);
        X.big[plane][cblk] = (uint8_t)(coded ? big : 0);
    }
}

// ---- phase B, lanes 0..23: chroma lanes first resolve the elimination and the 2x2 DC of their plane (rdo.c:2640-2672); then Q^-1 -> T^-1 -> add/clip ->
// reconstruction and the outputs (rdo.c:1469-1496, transf.c:161-296) ----
// mb_intra: the macroblock counts as intra in rdo.c:2660 (I_BL, or a macroblock with an inherited prediction -- see SvcPredSrc)
HLB_HD void svc_lane_b(const SvcPlanes& P, int mbx, int mby, int lane, int qp, int qpc, bool mb_intra, hlb200_svc_mb_state_t& st, const SvcLane& L, const SvcXchg& X,
                       hlb200_mb_coeffs_t& out)
{
    const bool is_luma = lane < 16;
    const int plane = is_luma ? 0 : (lane - 16) >> 2, cblk = (lane - 16) & 3;
    const int pitch = is_luma ? P.W : (P.W >> 1);
    const int bx = is_luma ? blk_x(lane) : (cblk & 1) * 4, by = is_luma ? blk_y(lane) : (cblk >> 1) * 4;
    const int off = is_luma ? (mby * 16 + by) * pitch + mbx * 16 + bx : (mby * 8 + by) * pitch + mbx * 8 + bx;
    uint8_t* rec = is_luma ? P.rec_y : (plane ? P.rec_v : P.rec_u);
    int l2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) l2[i] = L.lv[i];
    bool use_res = L.coded;
    if (!is_luma) {
        int tot = 0, anybig = 0;
        unsigned ac_mask = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { tot += X.nnz[plane][k]; anybig |= X.big[plane][k]; ac_mask |= (unsigned)X.coded[plane][k] << k; }
        if (tot == 1 && !anybig) ac_mask = 0;   // exactly one +-1 AC coefficient in the plane: Single_ctr < 7 && TotalCoeffs == 1, rdo.c:2641-2649
        int dcl[4], dcr[4] = {0, 0, 0, 0};
        unsigned dc_mask = 0;
        const bool dc_tent = (X.dc_coef[plane][0] | X.dc_coef[plane][1] | X.dc_coef[plane][2] | X.dc_coef[plane][3]) != 0;
        if (dc_tent) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dcl[k] = X.dc_coef[plane][k];
            hadamard2x2(dcl);
            // rdo.c:2660 uses the macroblock's own intra flag: an inferred macroblock with inter prediction is not intra (mb.h:46,57), an I_BL one is
            quant_dc(dcl, 4, qpc, /*isIntra(MB)*/ mb_intra);
#pragma unroll
            for (int k = 0; k < 4; ++k) dc_mask |= (unsigned)(dcl[k] != 0) << k;
            if (dc_mask) {   // transf.c:612: f = H.c.H ; dcC = ((f*LS00) << (qP/6)) >> 5
#pragma unroll
                for (int k = 0; k < 4; ++k) dcr[k] = dcl[k];
                hadamard2x2(dcr);
                const int ls = 16 * kNormAdjust[qpc % 6][0];
#pragma unroll
                for (int k = 0; k < 4; ++k) dcr[k] = ((dcr[k] * ls) << (qpc / 6)) >> 5;
            }
        } else {   // ChromaDCLevel keeps its old content (rdo.c:2653: not entered)
#pragma unroll
            for (int k = 0; k < 4; ++k) dcl[k] = st.chroma_dc_level[plane][k];
        }
        const int mydc = dcr[cblk];
        use_res = mydc != 0 || ((ac_mask >> cblk) & 1);   // AC levels are used whenever the DC is non-zero (transf.c:236)
#pragma unroll
        for (int i = 15; i >= 1; --i) l2[i] = L.lv[i - 1];
        l2[0] = mydc;
        // outputs = the macroblock object's fields after the call; state = the same fields, carried to the next picture of the layer
#pragma unroll
        for (int i = 0; i < 16; ++i) { out.chroma_ac_level[plane][cblk][i] = L.lv[i]; st.chroma_ac_level[plane][cblk][i] = L.lv[i]; }
        if (cblk == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { out.chroma_dc_level[plane][k] = (int16_t)dcl[k]; st.chroma_dc_level[plane][k] = (int16_t)dcl[k]; }
            out.cbp_chroma_dc4x4[plane] = (uint8_t)dc_mask;
            out.cbp_chroma_ac4x4[plane] = (uint8_t)ac_mask;
        }
    } else {   // two levels per store; all zero when the block is not coded
        uint32_t* o = reinterpret_cast<uint32_t*>(&out.luma_level[lane][0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (uint32_t)(uint16_t)L.lv[2 * i] | ((uint32_t)(uint16_t)L.lv[2 * i + 1] << 16);
    }
    int c[16];
    if (use_res) {
        inv_zigzag4x4(l2, c);
        dequant4x4(c, is_luma ? qp : qpc, /*keep_dc*/ !is_luma);
        inv_transform4x4(c);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = 0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
        svc_st4(rec + off + r * pitch, clip255((int)L.pv[r * 4] + c[r * 4]), clip255((int)L.pv[r * 4 + 1] + c[r * 4 + 1]), clip255((int)L.pv[r * 4 + 2] + c[r * 4 + 2]),
                clip255((int)L.pv[r * 4 + 3] + c[r * 4 + 3]));
}

HLB_HD unsigned svc_luma_cbp(const SvcXchg& X)
{
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) m |= (unsigned)X.luma_coded[i] << i;
    return m;
}

// ---- Intra_Base resampling: the I_BL prediction of an enhancement-layer I picture from the reference layer's reconstruction (first half) ----
// _hl_codec_264_decode_svc_resample_intra_colour_comps decode_svc.c:2864 -> array construction :2952 (in an I picture every reference macroblock is intra: the
// array is a clamped gather, (G-280)/(G-281)) -> _hl_codec_264_decode_svc_interpol_intra_base :3071 (vertical pass (G-301), horizontal pass + clip (G-305);
// chroma always takes the two-tap table, :3116,:3150), sample locations by utils.c:1064-1157 (G.6.3) for frame macroblocks, no cropping, chroma phases 0
// (sps.c:810-813); fixed-point precision (G-43): shift = 16 for level_idc <= 30, else 31 - ceil(log2(refDim)) (QCIF -> CIF -> 4CIF runs its top layer at the
// latter).  The reference's per-macroblock window offsets are multiples of 16 samples, so the result is a function of the plane position only: one output
// sample per call, no state.  For a reference dimension that is a power of two and level_idc > 30 the reference's int32 `refDim << shift` is 2^31 (undefined): refused
// by the callers (svc_rs_precision_ok).
struct SvcRsAxis { int scale, add, shift4; };   // (G-45), (G-48), shift - 4
HLB_HD int svc_ceil_log2(int v) { int k = 0; while ((1 << k) < v) ++k; return k; }
HLB_HD bool svc_rs_precision_ok(int refDim, int level_idc) { return level_idc <= 30 || (refDim & (refDim - 1)) != 0; }
HLB_HD SvcRsAxis svc_rs_axis(int refDim, int scaledDim, int level_idc)
{
    SvcRsAxis a;
    const int shift = level_idc <= 30 ? 16 : 31 - svc_ceil_log2(refDim);
    a.scale = ((refDim << shift) + (scaledDim >> 1)) / scaledDim;
    a.add = (((refDim * 2) << (shift - 2)) + (scaledDim >> 1)) / scaledDim + (1 << (shift - 5));
    a.shift4 = shift - 4;
    return a;
}
HLB_HD int svc_rs_ref16(int p, const SvcRsAxis& a) { return ((p * a.scale + a.add) >> a.shift4) - 8; }   // (G-59)/(G-60), deltaX = 8

// Table G-9 (tables.h:626-643), 4 signed bytes per phase packed low byte first; chroma: {32 - 2p, 2p} (tables.h:647-664)
HLB_TABLE static const uint32_t kSvcRsLuma[16] = { 0x00002000u, 0xFF0220FFu, 0xFF041FFEu, 0xFF061EFDu, 0xFF081CFDu, 0xFF0B1AFCu, 0xFE0E18FCu, 0xFD1016FDu,
                                                   0xFD1313FDu, 0xFD1610FDu, 0xFC180EFEu, 0xFC1A0BFFu, 0xFD1C08FFu, 0xFD1E06FFu, 0xFE1F04FFu, 0xFF2002FFu };
HLB_HD int svc_rs_luma_tap(uint32_t packed, int k) { return (int)(int8_t)(packed >> (8 * k)); }

HLB_HD uint8_t svc_resample_px(const uint8_t* ref, int refW, int refH, const SvcRsAxis& ax, const SvcRsAxis& ay, int x, int y, bool chroma)
{
    const int x16 = svc_rs_ref16(x, ax), y16 = svc_rs_ref16(y, ay), xr = x16 >> 4, xp = x16 & 15, yr = y16 >> 4, yp = y16 & 15;
    int v = 0;
    if (chroma) {
        const uint8_t* r0 = ref + clip3(0, refH - 1, yr) * refW;
        const uint8_t* r1 = ref + clip3(0, refH - 1, yr + 1) * refW;
        const int xa = clip3(0, refW - 1, xr), xb = clip3(0, refW - 1, xr + 1);
        const int t0 = (32 - 2 * yp) * (int)HLB_LDG(r0 + xa) + 2 * yp * (int)HLB_LDG(r1 + xa);
        const int t1 = (32 - 2 * yp) * (int)HLB_LDG(r0 + xb) + 2 * yp * (int)HLB_LDG(r1 + xb);
        v = (32 - 2 * xp) * t0 + 2 * xp * t1;
    }
    else {
        int t[4] = {0, 0, 0, 0};
        const uint32_t fy = kSvcRsLuma[yp], fx = kSvcRsLuma[xp];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* row = ref + clip3(0, refH - 1, yr - 1 + k) * refW;
            const int f = svc_rs_luma_tap(fy, k);
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] += f * (int)HLB_LDG(row + clip3(0, refW - 1, xr - 1 + j));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v += svc_rs_luma_tap(fx, j) * t[j];
    }
    return (uint8_t)clip255((v + 512) >> 10);
}

// Four horizontally adjacent output samples (x0 .. x0 + 3, x0 a multiple of 4) of row y as one word -- what a thread of k_svc_resample_intra stores.  The two passes
// of the reference ((G-301) vertical, (G-305) horizontal) carry no rounding between them, so the filter is the exact 2-D sum  sum_k fy[k] sum_j fx[j] s[k][j]  and
// the inner (horizontal) sum is a byte dot product: the row's samples xs .. xs + 7 are fetched ONCE as three aligned words and shifted into an 8-byte window, each
// output takes its four bytes from it by a funnel shift (its reference column differs from the first output's by at most 3 when the layer is not smaller than the
// reference layer) and multiplies them with its packed taps (Table G-9 rows / the two-tap chroma weights).  12 word loads per thread instead of 64 byte loads with
// per-sample clamps.  Threads whose window would leave the row (picture edges) and layers that shrink take the per-sample form; the rows are clamped as the reference
// clamps them.  Same results as four svc_resample_px calls (tests/test_svc_bl_resample.py runs this form on the CPU against the oracle).
HLB_HD uint32_t svc_resample_row4(const uint8_t* ref, int refW, int refH, const SvcRsAxis& ax, const SvcRsAxis& ay, int x0, int y, bool chroma)
{
    const int y16 = svc_rs_ref16(y, ay), yr = y16 >> 4, yp = y16 & 15;
    int xr[4];
    uint32_t fx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x16 = svc_rs_ref16(x0 + i, ax), xp = x16 & 15;
        xr[i] = x16 >> 4;
        fx[i] = chroma ? (uint32_t)(32 - 2 * xp) | ((uint32_t)(2 * xp) << 8) : kSvcRsLuma[xp];
    }
    const int xs = xr[0] - (chroma ? 0 : 1), xa = xs & ~3;
    const int d1 = xr[1] - xr[0], d2 = xr[2] - xr[0], d3 = xr[3] - xr[0];
    const bool fast = xs >= 0 && xa + 12 <= refW && (unsigned)d1 <= 3u && (unsigned)d2 <= 3u && (unsigned)d3 <= 3u;
    if (!fast) {
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) word |= (uint32_t)svc_resample_px(ref, refW, refH, ax, ay, x0 + i, y, chroma) << (8 * i);
        return word;
    }
    // the vertical taps as four plain integers (chroma: the two-tap weights, rows yr and yr + 1; luma: Table G-9 row yp, rows yr - 1 .. yr + 2)
    const uint32_t fyl = kSvcRsLuma[yp], sh = (uint32_t)(xs & 3) * 8, s1 = (uint32_t)d1 * 8, s2 = (uint32_t)d2 * 8, s3 = (uint32_t)d3 * 8;
    const int fy0 = chroma ? 0 : svc_rs_luma_tap(fyl, 0), fy1 = chroma ? 32 - 2 * yp : svc_rs_luma_tap(fyl, 1), fy2 = chroma ? 2 * yp : svc_rs_luma_tap(fyl, 2),
              fy3 = chroma ? 0 : svc_rs_luma_tap(fyl, 3);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(ref + xa);   // row r of the window starts refW / 4 words further (plane widths are multiples of 8)
    const int pitch4 = refW >> 2, r0 = yr - 1;
    int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int fyk = k == 0 ? fy0 : (k == 1 ? fy1 : (k == 2 ? fy2 : fy3));
        if (chroma && (k == 0 || k == 3)) continue;
        const uint32_t* q = base + clip3(0, refH - 1, r0 + k) * pitch4;
        const uint32_t w0 = HLB_LDG(q), w1 = HLB_LDG(q + 1), w2 = HLB_LDG(q + 2);
        const uint32_t lo = p_shf_r(w0, w1, sh), hi = p_shf_r(w1, w2, sh);   // samples xs .. xs + 7 of the row
        acc0 += fyk * p_dp4a_us(lo, fx[0], 0);
        acc1 += fyk * p_dp4a_us(p_shf_r(lo, hi, s1), fx[1], 0);
        acc2 += fyk * p_dp4a_us(p_shf_r(lo, hi, s2), fx[2], 0);
        acc3 += fyk * p_dp4a_us(p_shf_r(lo, hi, s3), fx[3], 0);
    }
    return pack4_sat((acc0 + 512) >> 10, (acc1 + 512) >> 10, (acc2 + 512) >> 10, (acc3 + 512) >> 10);
}

}  // namespace hlb
