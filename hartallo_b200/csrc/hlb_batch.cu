// hlb_batch.cu -- whole-frame batch kernels of the stateless part of the hot path:
// luma/chroma fractional interpolation, residual -> T -> Q -> Q^-1 -> T^-1 -> reconstruction, SAD/SATD and the
// independent ME candidate cost.  All are HBM-bound integer/byte kernels (no tensor cores: 4x4 butterflies are not
// a dense contraction); one thread owns one 4x4 block in registers, one warp owns one (or two) macroblocks.
#include <cuda.h>

#include "hlb_common.cuh"
#include "hlb_fast.cuh"
#include "hlb_svc.cuh"
#include "hlb_svc_derive.cuh"

namespace hlb {

// ------------------------------------------------------------------------------------------------------------------
// 9x9 window around a 4x4 block (rows/cols -2..+6), fetched with the per-sample clamp of the reference's index
// table (source/h264/hl_codec_264_interpol.c:108-131).  (X,Y) = position of output pixel (0,0) AFTER the partition
// origin clip of pred_inter.c:395-396.  Fully unrolled: the window lives in registers.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_win9(const uint8_t* __restrict__ plane, int W, int H, int X, int Y, uint8_t t[81])
{
    if (X >= 2 && Y >= 2 && X + 7 <= W && Y + 7 <= H) {
        const uint8_t* p = plane + (Y - 2) * W + (X - 2);
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
            for (int c = 0; c < 9; ++c) t[r * 9 + c] = __ldg(p + r * W + c);
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int y = clip3(0, H - 1, Y - 2 + r);
#pragma unroll
            for (int c = 0; c < 9; ++c) t[r * 9 + c] = __ldg(plane + y * W + clip3(0, W - 1, X - 2 + c));
        }
    }
}

// ---------------- luma interpolation: one thread per 4x4 block, blocks in PICTURE raster order -------------------------------------------------
// (a warp stores 32 adjacent words of each of its four output rows; the 9x9 window of an interior block is read straight from the plane as aligned 32-bit
// words through the read-only path and funnel-shifted into place, the prediction is formed by the packed formulation of hlb_fast.cuh).
// blockIdx.y = picture of a batch: the planes of consecutive pictures lie `stride` bytes apart, their motion fields nmb entries apart
#ifndef HLB_IL_MINB
#define HLB_IL_MINB 5   // 48 registers: +5 % (profiles/r02v2_variants.log)
#endif
__global__ void __launch_bounds__(256, HLB_IL_MINB) k_interp_luma(const uint8_t* __restrict__ ref, int W, int H, int mbw, int nmb,
                                                     const hlb200_mb_motion_t* __restrict__ motion, uint8_t* __restrict__ pred, size_t stride, uint32_t rcp_mbw)
{
    const int bw = W >> 2, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bw * (H >> 2)) return;
    ref += blockIdx.y * stride; pred += blockIdx.y * stride; motion += (size_t)blockIdx.y * nmb;
#ifdef HLB_IL_RASTER   // blocks in picture raster order: a warp = one row of 32 blocks = a strip of 8 macroblocks (whole sectors per store, but up to 8 x 4 partitions per warp)
    const int gy4 = t / bw, gx = (t - gy4 * bw) << 2, gy = gy4 << 2;
    const int mbx = gx >> 4, mby = gy >> 4, bx = gx & 15, by = gy & 15;
#else                  // blocks in macroblock order (luma4x4BlkIdx inside): a warp = two macroblocks, so it meets the fractional classes of two macroblocks' partitions only
    const int mb = t >> 4, b4 = t & 15, mby = div_rcp(mb, mbw, rcp_mbw), mbx = mb - mby * mbw, bx = blk_x(b4), by = blk_y(b4);
    const int gx = mbx * 16 + bx, gy = mby * 16 + by;
#endif
    const hlb200_mb_motion_t* m = motion + mby * mbw + mbx;
    const PartGeom g = part_of(m->part_mode, m->sub_mode, bx, by);
    const int mvx = m->mv[g.part][g.sub][0], mvy = m->mv[g.part][g.sub][1];
    // origin clip applies to the PARTITION origin
    const int X = clip3(-17, W + 17, mbx * 16 + g.ox + (mvx >> 2)) + (bx - g.ox);
    const int Y = clip3(-17, H + 17, mby * 16 + g.oy + (mvy >> 2)) + (by - g.oy);
    Rows4 o;
    if (X >= 2 && Y >= 2 && X + 7 <= W && Y + 7 <= H) o = fast_pred_luma_staged<LdReadOnly>(reinterpret_cast<const uint32_t*>(ref), W >> 2, X, Y, mvx & 3, mvy & 3);
    else {
        // the window touches the picture edge: staged with the reference's per-sample clamp (interpol.c:108-131), 12 bytes per row
        uint32_t win[9 * 3];
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            const uint8_t* row = ref + clip3(0, H - 1, Y - 2 + r) * W;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint32_t v = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) v |= (uint32_t)__ldg(row + clip3(0, W - 1, X - 2 + q * 4 + c)) << (8 * c);
                win[r * 3 + q] = v;
            }
        }
        o = fast_pred_luma_staged<LdPlain>(win, 3, 2, 2, mvx & 3, mvy & 3);
    }
    uint8_t* dst = pred + (size_t)gy * W + gx;
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t*>(dst + r * W) = o.r[r];
}

// ---------------- chroma interpolation: one thread per four horizontally adjacent samples of one plane, PICTURE raster order ---------------------
// (8 luma samples wide: one motion vector unless the macroblock is split into 4-wide sub-partitions, then two).  Written as ONE control flow for both cases:
// every thread forms its two sample pairs from their own vectors (equal in the common case), so warps that straddle macroblocks of different partition layouts
// do not run the two variants one after the other (the first packed version did: 360 warp instructions per thread, 17 of 32 lanes active on average).
// The sample pairs come from fast_chroma_two (hlb_fast.cuh).  Measured again at the end of round 2 (profiles/r02v8_chroma_four_dropped.log): a second path that fetches the
// two rows once for the four samples when both halves share a vector executes fewer instructions per thread (8 loads -> 4) but made the kernel 17 % SLOWER (0.141 -> 0.117
// of the HBM peak) -- the test field mixes layouts inside a warp and both paths run.
__global__ void __launch_bounds__(256) k_interp_chroma(const uint8_t* __restrict__ ref_u, const uint8_t* __restrict__ ref_v, int W, int H, int mbw, int nmb,
                                                       const hlb200_mb_motion_t* __restrict__ motion, uint8_t* __restrict__ pred_u, uint8_t* __restrict__ pred_v, size_t stride, uint32_t rcp_sw)
{
    const int Wc = W >> 1, Hc = H >> 1, sw = Wc >> 2, per_plane = sw * Hc;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * per_plane) return;
    const int plane = t >= per_plane, s = t - plane * per_plane, cy = div_rcp(s, sw, rcp_sw), cx = (s - cy * sw) << 2;
    const uint8_t* rp = (plane ? ref_v : ref_u) + blockIdx.y * stride;
    uint8_t* dst = (plane ? pred_v : pred_u) + blockIdx.y * stride + (size_t)cy * Wc + cx;
    const hlb200_mb_motion_t* m = motion + (size_t)blockIdx.y * nmb + (cy >> 3) * mbw + (cx >> 3);
    // the thread's eight luma columns lie in one macroblock partition; the sub-partition of its left / right half (part_of, hlb_common.cuh)
    const int hx = (cx >> 2) & 1, hy = (cy >> 2) & 1, pm = m->part_mode;
    const int part = pm == 0 ? 0 : (pm == 1 ? hy : (pm == 2 ? hx : hy * 2 + hx));
    const int sm = pm == 3 ? m->sub_mode[part] : 0, row = (cy >> 1) & 1;   // row: upper / lower 4 luma rows of the 8x8
    const int sub_l = (sm & 1 ? row : 0) << (sm >> 1), sub_r = sub_l + (sm >> 1);
    const uint32_t* mv = reinterpret_cast<const uint32_t*>(&m->mv[part][0][0]);
    const uint32_t vl = __ldg(mv + sub_l), vr = __ldg(mv + sub_r);
    const int lx = (int)(int16_t)(vl & 0xffffu), ly = (int)vl >> 16, rx = (int)(int16_t)(vr & 0xffffu), ry = (int)vr >> 16;
    const uint32_t out = fast_chroma_two(rp, Wc, Hc, cx + (lx >> 3), cy + (ly >> 3), lx & 7, ly & 7) | (fast_chroma_two(rp, Wc, Hc, cx + 2 + (rx >> 3), cy + (ry >> 3), rx & 7, ry & 7) << 16);
    *reinterpret_cast<uint32_t*>(dst) = out;
}

// ---------------- residual coding + reconstruction: one thread per 4x4 block, warps of one kind of block ----------------------------------------
// The luma blocks of a picture (16 per macroblock, luma4x4BlkIdx order) and its chroma blocks (8 per macroblock: Cb 0..3, Cr 0..3) form two index spaces
// served by different CTAs of the same launch, so a warp never runs the luma and the chroma control flow one after the other (the first version gave 24 of
// 32 lanes a block and executed both paths per warp: 908 warp instructions per macroblock, issue-bound at 0.29 of the HBM peak).  A lane keeps its block as
// four packed rows: residual transform = byte dot products (hlb_fast.cuh), reconstruction = saturating pack, levels leave as four 8-byte stores, transposed over groups of four lanes so that every store instruction fills whole 32-byte sectors.
__device__ __forceinline__ void load4x4(const uint8_t* __restrict__ p, int pitch, uint8_t v[16])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + r * pitch));
        v[r * 4] = w & 255; v[r * 4 + 1] = (w >> 8) & 255; v[r * 4 + 2] = (w >> 16) & 255; v[r * 4 + 3] = w >> 24;
    }
}
struct TqConst {
    QuantK luma, chroma;         // luma: inter rounding offset; chroma AC: always the intra offset (rdo.c:2588)
    int dc_qbits1, dc_f2, dc_mf; // 2x2 chroma DC quantiser (quant.c:150-189) with the macroblock's own (inter) offset (rdo.c:2660)
    int dc_ls, dc_q6;            // its de-quantisation (transf.c:612): ((f * LevelScale(qP % 6, 0, 0)) << (qP / 6)) >> 5
};
__device__ __forceinline__ void quantk_device(QuantK& k, int qp, bool intra)   // quantk_make from the device tables
{
    const int r = qp % 6, q6 = qp / 6;
    k.qbits = 15 + q6; k.f_pos = (1 << k.qbits) / (intra ? 3 : 6); k.f_neg = (1 << k.qbits) - 1 - k.f_pos;
#pragma unroll
    for (int c = 0; c < 3; ++c) { k.mf[c] = kQuantMF[r][c]; k.dq_mul[c] = qp >= 24 ? (16 * kNormAdjust[r][c]) << (q6 - 4) : 16 * kNormAdjust[r][c]; }
    k.dq_shift = qp >= 24 ? 0 : 4 - q6; k.dq_round = qp >= 24 ? 0 : 1 << (3 - q6); k.zero_sad = -1;
}
// zig-zag levels two per word: word k = lv[2k] | lv[2k+1] << 16 from the raster block m
__device__ __forceinline__ uint32_t pair16(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }
// The 32-byte level lists of four neighbouring lanes (= four consecutive blocks: 128 contiguous bytes) leave as four stores in which the four lanes together
// write ONE full 32-byte sector each, instead of four sectors a quarter full: a 4x4 transpose of 8-byte elements over the lanes (two butterfly stages).
// In: P[j] = 8-byte piece j of this lane's block.  Out: P[k] = piece (lane & 3) of the block of lane (lane & ~3) + k.
__device__ __forceinline__ void transpose4_u2(uint2 P[4], int lane)
{
    const unsigned full = 0xffffffffu;
    const bool o1 = lane & 1, o2 = lane & 2;
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        const uint2 snd = o1 ? P[j] : P[j + 1];
        const uint2 rcv = make_uint2(__shfl_xor_sync(full, snd.x, 1), __shfl_xor_sync(full, snd.y, 1));
        if (o1) P[j] = rcv; else P[j + 1] = rcv;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const uint2 snd = o2 ? P[j] : P[j + 2];
        const uint2 rcv = make_uint2(__shfl_xor_sync(full, snd.x, 2), __shfl_xor_sync(full, snd.y, 2));
        if (o2) P[j] = rcv; else P[j + 2] = rcv;
    }
}
#ifndef HLB_TQ_MINB
#define HLB_TQ_MINB 12   // 40 registers (40 bytes spilled): 48 resident warps instead of 40 -- 0.4935 -> 0.519 of the HBM peak at 128 pictures per launch (profiles/r02v_tq_variants.log; 16 blocks: 0.463)
#endif
__global__ void __launch_bounds__(128, HLB_TQ_MINB) k_tq_recon(const uint8_t* __restrict__ src_y, const uint8_t* __restrict__ src_u, const uint8_t* __restrict__ src_v,
                                                  const uint8_t* __restrict__ pred_y, const uint8_t* __restrict__ pred_u, const uint8_t* __restrict__ pred_v,
                                                  int W, int H, int mbw, int nmb, int qp, int qpc, hlb200_mb_coeffs_t* __restrict__ coeffs,
                                                  uint8_t* __restrict__ rec_y, uint8_t* __restrict__ rec_u, uint8_t* __restrict__ rec_v, size_t stride, int luma_ctas, uint32_t rcp_mbw)
{
    __shared__ TqConst K;
    if (threadIdx.x < 2) quantk_device(threadIdx.x ? K.chroma : K.luma, threadIdx.x ? qpc : qp, threadIdx.x != 0);
    if (threadIdx.x == 2) {
        K.dc_qbits1 = 16 + qpc / 6; K.dc_f2 = ((1 << (15 + qpc / 6)) / 6) << 1; K.dc_mf = kQuantMF[qpc % 6][0];
        K.dc_ls = 16 * kNormAdjust[qpc % 6][0]; K.dc_q6 = qpc / 6;
    }
    __syncthreads();
    const size_t po = blockIdx.y * stride;
    coeffs += (size_t)blockIdx.y * nmb;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < luma_ctas) {
        // ---------------- luma: t = macroblock * 16 + luma4x4BlkIdx ----------------
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = t < nmb * 16;
        const int mb = valid ? t >> 4 : 0, b = t & 15, mby = div_rcp(mb, mbw, rcp_mbw), mbx = mb - mby * mbw;
        const size_t off = po + (size_t)(mby * 16 + blk_y(b)) * W + mbx * 16 + blk_x(b);
        Rows4 sv, pv;
        bool res_nz = false;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            sv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(src_y + off + (size_t)y * W)) : 0u;
            pv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(pred_y + off + (size_t)y * W)) : 0u;
            res_nz |= sv.r[y] != pv.r[y];
        }
        int m[16];
        uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
        Rows4 rec = pv;
        bool coded = false;
        if (res_nz) {
            fast_fwd_transform(sv, pv, m);
            fast_quant(m, K.luma);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) any |= (uint32_t)m[i];
            coded = any != 0;
            if (coded) {
                // kZigzag = 0 1 4 8 | 5 2 3 6 | 9 12 13 10 | 7 11 14 15
                o0 = make_uint4(pair16(m[0], m[1]), pair16(m[4], m[8]), pair16(m[5], m[2]), pair16(m[3], m[6]));
                o1 = make_uint4(pair16(m[9], m[12]), pair16(m[13], m[10]), pair16(m[7], m[11]), pair16(m[14], m[15]));
                fast_dequant_inverse(m, K.luma, false);
                rec = fast_recon_clip(pv, m);
            }
        }
        uint2 P[4] = {make_uint2(o0.x, o0.y), make_uint2(o0.z, o0.w), make_uint2(o1.x, o1.y), make_uint2(o1.z, o1.w)};
        transpose4_u2(P, lane);   // all 32 lanes take part (nmb * 16 is a multiple of 4: a group of four lanes is valid or invalid as a whole)
        if (valid) {
#pragma unroll
            for (int y = 0; y < 4; ++y) *reinterpret_cast<uint32_t*>(rec_y + off + (size_t)y * W) = rec.r[y];
            uint2* o = reinterpret_cast<uint2*>(coeffs[mb].luma_level[b & ~3]) + (b & 3);   // the structure is 792 bytes: 8-byte alignment is all it guarantees
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k * 4] = P[k];   // piece (b & 3) of block (b & ~3) + k
        }
        const unsigned lb = __ballot_sync(full, valid && coded);
        if (valid && b == 0) coeffs[mb].cbp_luma4x4 = (uint16_t)((lb >> (lane & 16)) & 0xffffu);
        return;
    }
    // ---------------- chroma: t = macroblock * 8 + plane * 4 + block (raster) ----------------
    const int t = ((int)blockIdx.x - luma_ctas) * blockDim.x + threadIdx.x;
    const bool valid = t < nmb * 8;
    const int mb = valid ? t >> 3 : 0, plane = (t >> 2) & 1, cblk = t & 3, mby = div_rcp(mb, mbw, rcp_mbw), mbx = mb - mby * mbw, Wc = W >> 1;
    const size_t off = po + (size_t)(mby * 8 + (cblk >> 1) * 4) * Wc + mbx * 8 + (cblk & 1) * 4;
    const uint8_t* s = plane ? src_v : src_u;
    const uint8_t* p = plane ? pred_v : pred_u;
    uint8_t* r = plane ? rec_v : rec_u;
    Rows4 sv, pv;
    bool res_nz = false;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        sv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(s + off + (size_t)y * Wc)) : 0u;
        pv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(p + off + (size_t)y * Wc)) : 0u;
        res_nz |= sv.r[y] != pv.r[y];
    }
    int m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 0;
    int dc_coef = 0;   // pre-quantisation W00 (rdo.c:2591)
    if (res_nz) {
        fast_fwd_transform(sv, pv, m);
        dc_coef = m[0];
        fast_quant(m, K.chroma);
    }
    // single-coefficient elimination over the plane's four blocks (rdo.c:2599-2625): AC levels only
    int nnz_ac = 0, big = 0;
#pragma unroll
    for (int i = 1; i < 16; ++i) { nnz_ac += (m[i] != 0); big |= (m[i] > 1) | (m[i] < -1); }
    const int base = lane & ~3;
    int tot = 0, anybig = 0, dcl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        tot += __shfl_sync(full, nnz_ac, base + k);
        anybig |= __shfl_sync(full, big, base + k);
        dcl[k] = __shfl_sync(full, dc_coef, base + k);
    }
    const unsigned ac_ballot = __ballot_sync(full, nnz_ac != 0);
    unsigned ac_mask = (ac_ballot >> base) & 15u;   // CodedBlockPatternChromaAC4x4 of my plane
    if (tot == 1 && !anybig) ac_mask = 0;           // exactly one +-1 AC coefficient in the plane
    unsigned dc_mask = 0;
    int mydc = 0;
    if ((dcl[0] | dcl[1] | dcl[2] | dcl[3]) != 0) {
        hadamard2x2(dcl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int w = dcl[k], z = (iabs(w) * K.dc_mf + K.dc_f2) >> K.dc_qbits1;
            dcl[k] = w >= 0 ? z : -z;
            dc_mask |= (unsigned)(dcl[k] != 0) << k;
        }
        if (dc_mask) {   // transf.c:612: f = H.c.H ; dcC = ((f*LS00) << (qP/6)) >> 5
            int dcr[4] = {dcl[0], dcl[1], dcl[2], dcl[3]};
            hadamard2x2(dcr);
            const int f = cblk == 0 ? dcr[0] : (cblk == 1 ? dcr[1] : (cblk == 2 ? dcr[2] : dcr[3]));
            mydc = ((f * K.dc_ls) << K.dc_q6) >> 5;
        }
    } else { dcl[0] = dcl[1] = dcl[2] = dcl[3] = 0; }
    // levels as the reference leaves them (before the elimination decision; position 0 of the 15-entry AC list = zig-zag index 1)
    const uint4 o0 = make_uint4(pair16(m[1], m[4]), pair16(m[8], m[5]), pair16(m[2], m[3]), pair16(m[6], m[9]));
    const uint4 o1 = make_uint4(pair16(m[12], m[13]), pair16(m[10], m[7]), pair16(m[11], m[14]), pair16(m[15], 0));
    Rows4 rec = pv;
    if (mydc != 0 || ((ac_mask >> cblk) & 1)) {   // transf.c:236: AC levels are used whenever the DC is non-zero
        m[0] = mydc;
        fast_dequant_inverse(m, K.chroma, /*keep_dc*/ true);
        rec = fast_recon_clip(pv, m);
    }
    uint2 P[4] = {make_uint2(o0.x, o0.y), make_uint2(o0.z, o0.w), make_uint2(o1.x, o1.y), make_uint2(o1.z, o1.w)};
    transpose4_u2(P, lane);
    if (valid) {
#pragma unroll
        for (int y = 0; y < 4; ++y) *reinterpret_cast<uint32_t*>(r + off + (size_t)y * Wc) = rec.r[y];
        uint2* o = reinterpret_cast<uint2*>(coeffs[mb].chroma_ac_level[plane][0]) + cblk;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k * 4] = P[k];
        if (cblk == 0) {
            *reinterpret_cast<uint2*>(coeffs[mb].chroma_dc_level[plane]) = make_uint2(pair16(dcl[0], dcl[1]), pair16(dcl[2], dcl[3]));
            coeffs[mb].cbp_chroma_dc4x4[plane] = (uint8_t)dc_mask;
            coeffs[mb].cbp_chroma_ac4x4[plane] = (uint8_t)ac_mask;
        }
    }
}

// ---------------- SVC enhancement-layer macroblocks (base mode inter / I_BL): prediction + residual coding + reconstruction ------------------------------
// Same organisation as k_tq_recon: the luma blocks (16 per macroblock) and the chroma blocks (8 per macroblock) of a picture are two index spaces served by
// different CTAs of one launch, one thread per 4x4 block, the block in four packed registers, chroma planes exchanging through 4-lane shuffles.  On top of it:
// the PREDICTION is formed in the kernel (BL = false: from the layer's reference picture with the inferred partitions / vectors, luma by the staged six-tap form
// of hlb_fast.cuh, chroma by byte dot products; BL = true: P.ref_* are prediction planes) and the per-macroblock state the reference carries from picture to
// picture (ChromaACLevel of blocks without residual, ChromaDCLevel of planes without DC) travels in / out through `state`.  The formulation checked against the
// reference's trace on the CPU is the per-lane one of hlb_svc.cuh (tools/emu/svc_emu.cpp); this kernel is checked against the oracle and the golden
// fixtures on the GPU (tests/test_svc_inter.py).  Round 1's kernel (one warp per macroblock on hlb_svc.cuh's phases) measured 0.056 of the HBM peak.
#ifndef HLB_SVC_MINB
#define HLB_SVC_MINB 10   // 48 registers: +1.4 %
#endif
template <bool BL>
__global__ void __launch_bounds__(128, HLB_SVC_MINB) k_svc_inter_recon(SvcPlanes P, int mbw, int nmb, int qp, int qpc, const hlb200_mb_motion_t* __restrict__ motion,
                                                         hlb200_svc_mb_state_t* __restrict__ state, hlb200_mb_coeffs_t* __restrict__ coeffs, size_t stride, int luma_ctas, uint32_t rcp_mbw)
{
    __shared__ TqConst K;
    if (threadIdx.x < 2) quantk_device(threadIdx.x ? K.chroma : K.luma, threadIdx.x ? qpc : qp, true);   // luma too takes the intra offset here (rdo.c:1468)
    if (threadIdx.x == 2) { K.dc_qbits1 = 16 + qpc / 6; K.dc_f2 = 0; K.dc_mf = kQuantMF[qpc % 6][0]; K.dc_ls = 16 * kNormAdjust[qpc % 6][0]; K.dc_q6 = qpc / 6; }
    __syncthreads();
    const size_t po = blockIdx.y * stride;
    coeffs += (size_t)blockIdx.y * nmb; state += (size_t)blockIdx.y * nmb;
    if (!BL) motion += (size_t)blockIdx.y * nmb;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, W = P.W, H = P.H;
    if ((int)blockIdx.x < luma_ctas) {
        // ---------------- luma: t = macroblock * 16 + luma4x4BlkIdx ----------------
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = t < nmb * 16;
        const int mb = valid ? t >> 4 : 0, b = t & 15, mby = div_rcp(mb, mbw, rcp_mbw), mbx = mb - mby * mbw, bx = blk_x(b), by = blk_y(b);
        const size_t off = po + (size_t)(mby * 16 + by) * W + mbx * 16 + bx;
        Rows4 sv, pv;
        if (BL) {
#pragma unroll
            for (int y = 0; y < 4; ++y) pv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(P.ref_y + off + (size_t)y * W)) : 0u;
        } else {
            const SvcPredSrc ps = svc_pred_src(motion, mb, mbw, nmb);
            const SvcPart g = svc_part_of(ps.m->part_mode, ps.m->sub_mode, bx, by);
            const int mvx = ps.m->mv[g.part][g.sub][0], mvy = ps.m->mv[g.part][g.sub][1];
            // the origin clip applies to the PARTITION origin (pred_inter.c:395-396)
            const int X = clip3(-17, W + 17, ps.mbx * 16 + g.ox + (mvx >> 2)) + (bx - g.ox), Y = clip3(-17, H + 17, ps.mby * 16 + g.oy + (mvy >> 2)) + (by - g.oy);
            const uint8_t* ref = P.ref_y + po;
            if (X >= 2 && Y >= 2 && X + 7 <= W && Y + 7 <= H) pv = fast_pred_luma_staged<LdReadOnly>(reinterpret_cast<const uint32_t*>(ref), W >> 2, X, Y, mvx & 3, mvy & 3);
            else {   // the window touches the picture edge: staged with the reference's per-sample clamp (interpol.c:108-131)
                uint32_t win[9 * 3];
#pragma unroll 1
                for (int r = 0; r < 9; ++r) {
                    const uint8_t* row = ref + (size_t)clip3(0, H - 1, Y - 2 + r) * W;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        uint32_t v = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) v |= (uint32_t)__ldg(row + clip3(0, W - 1, X - 2 + q * 4 + c)) << (8 * c);
                        win[r * 3 + q] = v;
                    }
                }
                pv = fast_pred_luma_staged<LdPlain>(win, 3, 2, 2, mvx & 3, mvy & 3);
            }
        }
        bool res_nz = false;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            sv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(P.src_y + off + (size_t)y * W)) : 0u;
            if (!valid) pv.r[y] = 0;
            res_nz |= sv.r[y] != pv.r[y];
        }
        int m[16];
        uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;   // LumaLevel is cleared first (rdo.c:1453,1462)
        Rows4 rec = pv;
        bool coded = false;
        if (res_nz) {
            fast_fwd_transform(sv, pv, m);
            fast_quant(m, K.luma);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) any |= (uint32_t)m[i];
            coded = any != 0;
            if (coded) {
                o0 = make_uint4(pair16(m[0], m[1]), pair16(m[4], m[8]), pair16(m[5], m[2]), pair16(m[3], m[6]));
                o1 = make_uint4(pair16(m[9], m[12]), pair16(m[13], m[10]), pair16(m[7], m[11]), pair16(m[14], m[15]));
                fast_dequant_inverse(m, K.luma, false);
                rec = fast_recon_clip(pv, m);
            }
        }
        uint2 Pq[4] = {make_uint2(o0.x, o0.y), make_uint2(o0.z, o0.w), make_uint2(o1.x, o1.y), make_uint2(o1.z, o1.w)};
        transpose4_u2(Pq, lane);
        if (valid) {
#pragma unroll
            for (int y = 0; y < 4; ++y) *reinterpret_cast<uint32_t*>(P.rec_y + off + (size_t)y * W) = rec.r[y];
            uint2* o = reinterpret_cast<uint2*>(coeffs[mb].luma_level[b & ~3]) + (b & 3);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k * 4] = Pq[k];
        }
        const unsigned lb = __ballot_sync(full, valid && coded);
        if (valid && b == 0) coeffs[mb].cbp_luma4x4 = (uint16_t)((lb >> (lane & 16)) & 0xffffu);
        return;
    }
    // ---------------- chroma: t = macroblock * 8 + plane * 4 + block (raster) ----------------
    const int t = ((int)blockIdx.x - luma_ctas) * blockDim.x + threadIdx.x;
    const bool valid = t < nmb * 8;
    const int mb = valid ? t >> 3 : 0, plane = (t >> 2) & 1, cblk = t & 3, mby = div_rcp(mb, mbw, rcp_mbw), mbx = mb - mby * mbw, Wc = W >> 1, Hc = H >> 1;
    const int bx = (cblk & 1) * 4, by = (cblk >> 1) * 4;
    const size_t off = po + (size_t)(mby * 8 + by) * Wc + mbx * 8 + bx;
    const uint8_t* s = plane ? P.src_v : P.src_u;
    const uint8_t* refp = (plane ? P.ref_v : P.ref_u) + po;
    uint8_t* r = plane ? P.rec_v : P.rec_u;
    Rows4 sv, pv;
    bool mb_intra = BL;   // the macroblock counts as intra in the 2x2 DC quantisation (rdo.c:2660): I_BL, or a macroblock with an inherited prediction
    if (BL) {
#pragma unroll
        for (int y = 0; y < 4; ++y) pv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(refp - po + off + (size_t)y * Wc)) : 0u;
    } else {
        const SvcPredSrc ps = svc_pred_src(motion, mb, mbw, nmb);
        mb_intra = ps.inherited;
#pragma unroll 1
        for (int y = 0; y < 4; ++y) {
            uint32_t row = 0;
#pragma unroll
            for (int x = 0; x < 4; x += 2) {   // a 2x2 chroma area is the smallest one with its own motion vector
                const SvcPart g = svc_part_of(ps.m->part_mode, ps.m->sub_mode, (bx + x) * 2, (by + y) * 2);
                const int mvx = ps.m->mv[g.part][g.sub][0], mvy = ps.m->mv[g.part][g.sub][1];
                row |= fast_chroma_two(refp, Wc, Hc, ps.mbx * 8 + bx + x + (mvx >> 3), ps.mby * 8 + by + y + (mvy >> 3), mvx & 7, mvy & 7) << (8 * x);
            }
            pv.r[y] = row;
        }
    }
    bool res_nz = false;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        sv.r[y] = valid ? __ldg(reinterpret_cast<const uint32_t*>(s + off + (size_t)y * Wc)) : 0u;
        if (!valid) pv.r[y] = 0;
        res_nz |= sv.r[y] != pv.r[y];
    }
    // ChromaACLevel as the macroblock object holds it: 15 AC levels + the never-written [15]; kept from earlier pictures when the residual is zero
    hlb200_svc_mb_state_t& st = state[mb];
    uint32_t lvw[8];
    {
        const uint2* sp = reinterpret_cast<const uint2*>(st.chroma_ac_level[plane][cblk]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint2 v = valid ? sp[k] : make_uint2(0, 0); lvw[2 * k] = v.x; lvw[2 * k + 1] = v.y; }
    }
    int m[16];
    int dc_coef = 0;
    if (res_nz) {
        fast_fwd_transform(sv, pv, m);
        dc_coef = m[0];
        fast_quant(m, K.chroma);
        // Scan4x4_AC_C (utils.h:183): zig-zag positions 1..15 -> elements 0..14; element 15 keeps what it held
        lvw[0] = pair16(m[1], m[4]); lvw[1] = pair16(m[8], m[5]); lvw[2] = pair16(m[2], m[3]); lvw[3] = pair16(m[6], m[9]);
        lvw[4] = pair16(m[12], m[13]); lvw[5] = pair16(m[10], m[7]); lvw[6] = pair16(m[11], m[14]); lvw[7] = (lvw[7] & 0xffff0000u) | (uint32_t)(uint16_t)m[15];
    }
    // coded / counts look at all 16 elements of the list as it now stands (rdo.c:2599-2625 through the shared chroma function)
    int nnz = 0, big = 0;
    if (res_nz) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int lo = (int)(int16_t)(lvw[k] & 0xffffu), hi = (int)lvw[k] >> 16;
            nnz += (lo != 0) + (hi != 0); big |= (lo > 1) | (lo < -1) | (hi > 1) | (hi < -1);
        }
    }
    const bool coded = nnz != 0;
    const int base = lane & ~3;
    int tot = 0, anybig = 0, dcl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        tot += __shfl_sync(full, nnz, base + k);
        anybig |= __shfl_sync(full, big, base + k);
        dcl[k] = __shfl_sync(full, dc_coef, base + k);
    }
    const unsigned ac_ballot = __ballot_sync(full, coded);
    unsigned ac_mask = (ac_ballot >> base) & 15u;
    if (tot == 1 && !anybig) ac_mask = 0;   // exactly one +-1 AC coefficient in the plane: Single_ctr < 7 && TotalCoeffs == 1, rdo.c:2641-2649
    unsigned dc_mask = 0;
    int mydc = 0;
    if ((dcl[0] | dcl[1] | dcl[2] | dcl[3]) != 0) {
        hadamard2x2(dcl);
        const int f2 = ((1 << (K.dc_qbits1 - 1)) / (mb_intra ? 3 : 6)) << 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int w = dcl[k], z = (iabs(w) * K.dc_mf + f2) >> K.dc_qbits1;
            dcl[k] = w >= 0 ? z : -z;
            dc_mask |= (unsigned)(dcl[k] != 0) << k;
        }
        if (dc_mask) {
            int dcr[4] = {dcl[0], dcl[1], dcl[2], dcl[3]};
            hadamard2x2(dcr);
            const int f = cblk == 0 ? dcr[0] : (cblk == 1 ? dcr[1] : (cblk == 2 ? dcr[2] : dcr[3]));
            mydc = ((f * K.dc_ls) << K.dc_q6) >> 5;
        }
    } else if (valid) {   // ChromaDCLevel keeps its old content (rdo.c:2653: not entered)
        const uint2 v = *reinterpret_cast<const uint2*>(st.chroma_dc_level[plane]);
        dcl[0] = (int)(int16_t)(v.x & 0xffffu); dcl[1] = (int)v.x >> 16; dcl[2] = (int)(int16_t)(v.y & 0xffffu); dcl[3] = (int)v.y >> 16;
    }
    Rows4 rec = pv;
    if (mydc != 0 || ((ac_mask >> cblk) & 1)) {   // AC levels (possibly the stale ones) are used whenever the DC is non-zero (transf.c:236)
        const int e0 = (int)(int16_t)(lvw[0] & 0xffffu), e1 = (int)lvw[0] >> 16, e2 = (int)(int16_t)(lvw[1] & 0xffffu), e3 = (int)lvw[1] >> 16;
        const int e4 = (int)(int16_t)(lvw[2] & 0xffffu), e5 = (int)lvw[2] >> 16, e6 = (int)(int16_t)(lvw[3] & 0xffffu), e7 = (int)lvw[3] >> 16;
        const int e8 = (int)(int16_t)(lvw[4] & 0xffffu), e9 = (int)lvw[4] >> 16, e10 = (int)(int16_t)(lvw[5] & 0xffffu), e11 = (int)lvw[5] >> 16;
        const int e12 = (int)(int16_t)(lvw[6] & 0xffffu), e13 = (int)lvw[6] >> 16, e14 = (int)(int16_t)(lvw[7] & 0xffffu);
        int c[16] = {mydc, e0, e4, e5, e1, e3, e6, e11, e2, e7, e10, e12, e8, e9, e13, e14};
        fast_dequant_inverse(c, K.chroma, /*keep_dc*/ true);
        rec = fast_recon_clip(pv, c);
    }
    uint2 Pq[4] = {make_uint2(lvw[0], lvw[1]), make_uint2(lvw[2], lvw[3]), make_uint2(lvw[4], lvw[5]), make_uint2(lvw[6], lvw[7])};
    if (valid && res_nz) {   // state: the same fields, carried to the next picture of the layer (unchanged when the residual is zero)
        uint2* sp = reinterpret_cast<uint2*>(st.chroma_ac_level[plane][cblk]);
#pragma unroll
        for (int k = 0; k < 4; ++k) sp[k] = Pq[k];
    }
    transpose4_u2(Pq, lane);
    if (valid) {
#pragma unroll
        for (int y = 0; y < 4; ++y) *reinterpret_cast<uint32_t*>(r + off + (size_t)y * Wc) = rec.r[y];
        uint2* o = reinterpret_cast<uint2*>(coeffs[mb].chroma_ac_level[plane][0]) + cblk;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k * 4] = Pq[k];
        if (cblk == 0) {
            const uint2 d = make_uint2(pair16(dcl[0], dcl[1]), pair16(dcl[2], dcl[3]));
            *reinterpret_cast<uint2*>(coeffs[mb].chroma_dc_level[plane]) = d;
            *reinterpret_cast<uint2*>(st.chroma_dc_level[plane]) = d;
            coeffs[mb].cbp_chroma_dc4x4[plane] = (uint8_t)dc_mask;
            coeffs[mb].cbp_chroma_ac4x4[plane] = (uint8_t)ac_mask;
        }
    }
}

// ---------------- SAD / SATD of every 4x4 block of two planes -------------------------------------------------------
__global__ void __launch_bounds__(256) k_sad4x4(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int W, int H, int use_satd, int32_t* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int bw = W >> 2, nb = bw * (H >> 2);
    if (t >= nb) return;
    const int bx = (t % bw) * 4, by = (t / bw) * 4;
    uint8_t x[16], y[16];
    load4x4(a + by * W + bx, W, x);
    load4x4(b + by * W + bx, W, y);
    if (use_satd == 2) {   // hl_math_ssd4x4_u8, hl_math.c:360
        int s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int d = (int)x[i] - (int)y[i]; s += d * d; }
        out[t] = s;
    } else out[t] = use_satd ? satd16(x, y) : sad16(x, y);
}

// ---------------- edge-map homogeneity of every 8x8 block of a plane (hl_math_homogeneousity8x8_u8, hl_math.c:470-486) ----------------------------
// sum over the block of |dx| + |dy| (Sobel pair).  The 3x3 support of a block on the plane's border leaves the plane: such blocks get -1 (the encoder never
// asks for them, rdo.c:894-895 shifts its window inwards).  One warp per block: lane = two horizontally adjacent samples of four rows... kept simple: one
// thread per sample pair, 32 lanes x 2 samples = 64 samples, reduced by __reduce_add_sync.
__global__ void __launch_bounds__(256) k_homogeneity8x8(const uint8_t* __restrict__ p, int W, int H, int32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31, blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int bw = W >> 3, nb = bw * (H >> 3);
    if (blk >= nb) return;
    const int bx = (blk % bw) * 8, by = (blk / bw) * 8;
    if (bx == 0 || by == 0 || bx + 8 >= W || by + 8 >= H) { if (lane == 0) out[blk] = -1; return; }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = (lane & 3) * 2 + k, j = lane >> 2;
        const uint8_t* c = p + (size_t)(by + j) * W + bx + i;
        const uint8_t *up = c - W, *dn = c + W;
        const int dx = dn[-1] + 2 * dn[0] + dn[1] - up[-1] - 2 * up[0] - up[1], dy = up[1] + 2 * c[1] + dn[1] - up[-1] - 2 * c[-1] - dn[-1];
        s += iabs(dx) + iabs(dy);
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) out[blk] = s;
}

// ---------------- independent ME candidate cost (me_ds.c:527-688): 16 lanes per candidate ---------------------------
__global__ void __launch_bounds__(128) k_me_cost(const uint8_t* __restrict__ src, const uint8_t* __restrict__ ref, int W, int H, int qp,
                                                 const hlb200_me_cand_t* __restrict__ cands, int n, hlb200_me_cost_t* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci = t >> 4, k = t & 15;
    const bool valid = ci < n;
    const hlb200_me_cand_t cd = cands[valid ? ci : 0];
    const int bw = cd.part_w >> 2, bh = cd.part_h >> 2;
    const bool active = valid && k < bw * bh;
    const int bx = cd.part_x + (k % bw) * 4, by = cd.part_y + (k / bw) * 4;  // raster inside the partition
    int dist = 0, bits = 0, sctr = 0, tc = 0, t1 = 0, nzb = 0;
    if (active) {
        const int X = clip3(-17, W + 17, cd.mb_x * 16 + cd.part_x + (cd.mv_x >> 2)) + (bx - cd.part_x);
        const int Y = clip3(-17, H + 17, cd.mb_y * 16 + cd.part_y + (cd.mv_y >> 2)) + (by - cd.part_y);
        uint8_t win[81], pv[16], sv[16];
        load_win9(ref, W, H, X, Y, win);
        interp_luma_4x4(win + 20, 9, cd.mv_x & 3, cd.mv_y & 3, pv);
        load4x4(src + (cd.mb_y * 16 + by) * W + cd.mb_x * 16 + bx, W, sv);
        int m[16], lv[16];
        bool nz = false;
#pragma unroll
        for (int i = 0; i < 16; ++i) { m[i] = (int)sv[i] - (int)pv[i]; nz |= (m[i] != 0); }
        if (nz) {
            fwd_transform4x4(m);
            quant4x4_ac(m, qp, false);
            zigzag4x4(m, lv);
            nz = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) nz |= (lv[i] != 0);
        }
        if (nz) {
            const CavlcInfo ci2 = cavlc_block_info(lv, 16, false);
            bits = ci2.bits_rest; sctr = ci2.single_ctr; tc = ci2.total_coeff; t1 = ci2.trailing_ones; nzb = 1;
            int c[16];
            inv_zigzag4x4(lv, c);
            dequant4x4(c, qp, false);
            inv_transform4x4(c);
            uint8_t rec[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) rec[i] = (uint8_t)((int)pv[i] + c[i]);  // wraps (hl_math.h:261)
            dist = sad16(sv, rec);
        } else {
            dist = sad16(sv, pv);
        }
    }
    const int blk = active ? blk_idx_from_xy(bx, by) : 0;
    unsigned cbp = nzb ? (1u << blk) : 0u;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        dist += __shfl_xor_sync(0xffffffffu, dist, o, 16);
        bits += __shfl_xor_sync(0xffffffffu, bits, o, 16);
        sctr += __shfl_xor_sync(0xffffffffu, sctr, o, 16);
        cbp |= __shfl_xor_sync(0xffffffffu, cbp, o, 16);
    }
    if (valid) {
        hlb200_me_cost_t* o = out + ci;
        if (k == 0) { o->dist = dist; o->bits_rest = bits; o->single_ctr = sctr; o->cbp_luma4x4 = (uint16_t)cbp; o->pad = 0; }
    }
    // per-block counts: every lane of the candidate clears its slot, then active non-zero lanes write theirs
    if (valid) { out[ci].total_coeff[k] = 0; out[ci].trailing_ones[k] = 0; }
    __syncwarp();
    if (active && nzb) { out[ci].total_coeff[blk] = (uint8_t)tc; out[ci].trailing_ones[blk] = (uint8_t)t1; }
}

}  // namespace hlb

namespace hlb {
int launch_me_cost(const uint8_t* d_src, const uint8_t* d_ref, int W, int H, int qp, const hlb200_me_cand_t* d_cands, int n, hlb200_me_cost_t* d_out, cudaStream_t st);
}
using namespace hlb;

extern "C" {

// A batch = n_pics pictures whose planes lie frame_stride bytes apart (e.g. consecutive tight Y|U|V frames: stride = W*H*3/2) and whose
// per-macroblock arrays are contiguous (n_pics x nmb).  One launch for the whole batch: a single 1080p picture is only 3-16 MB, far too
// little to load HBM3e.
int hlb200_dev_interp_luma_batch(const uint8_t* d_ref_y, int width, int height, int n_pics, size_t frame_stride, const hlb200_mb_motion_t* d_motion, uint8_t* d_pred_y,
                                 void* cuda_stream)
{
    if (!d_ref_y || !d_motion || !d_pred_y || (width & 15) || (height & 15) || n_pics < 1 || n_pics > 65535) return HLB200_ERR_INVALID_PARAMETER;
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    k_interp_luma<<<dim3((nmb * 16 + 127) / 128, n_pics), 128, 0, (cudaStream_t)cuda_stream>>>(d_ref_y, width, height, mbw, nmb, d_motion, d_pred_y, frame_stride, host_rcp32(mbw, width, height));
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}
int hlb200_dev_interp_luma(const uint8_t* d_ref_y, int width, int height, const hlb200_mb_motion_t* d_motion, uint8_t* d_pred_y, void* cuda_stream)
{
    return hlb200_dev_interp_luma_batch(d_ref_y, width, height, 1, 0, d_motion, d_pred_y, cuda_stream);
}

int hlb200_dev_interp_chroma_batch(const uint8_t* d_ref_u, const uint8_t* d_ref_v, int width, int height, int n_pics, size_t frame_stride, const hlb200_mb_motion_t* d_motion,
                                   uint8_t* d_pred_u, uint8_t* d_pred_v, void* cuda_stream)
{
    if (!d_ref_u || !d_ref_v || !d_motion || !d_pred_u || !d_pred_v || (width & 15) || (height & 15) || n_pics < 1 || n_pics > 65535) return HLB200_ERR_INVALID_PARAMETER;
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    k_interp_chroma<<<dim3((nmb * 32 + 255) / 256, n_pics), 256, 0, (cudaStream_t)cuda_stream>>>(d_ref_u, d_ref_v, width, height, mbw, nmb, d_motion, d_pred_u, d_pred_v, frame_stride, host_rcp32(width >> 3, width, height));
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}
int hlb200_dev_interp_chroma(const uint8_t* d_ref_u, const uint8_t* d_ref_v, int width, int height, const hlb200_mb_motion_t* d_motion, uint8_t* d_pred_u,
                             uint8_t* d_pred_v, void* cuda_stream)
{
    return hlb200_dev_interp_chroma_batch(d_ref_u, d_ref_v, width, height, 1, 0, d_motion, d_pred_u, d_pred_v, cuda_stream);
}

int hlb200_dev_tq_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                              const uint8_t* d_pred_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                              hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream)
{
    if (!d_src_y || !d_pred_y || !d_coeffs || !d_recon_y || (width & 15) || (height & 15) || qp < 0 || qp > 51 || n_pics < 1 || n_pics > 65535) return HLB200_ERR_INVALID_PARAMETER;
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    if (!d_src_u || !d_src_v || !d_pred_u || !d_pred_v || !d_recon_u || !d_recon_v) return HLB200_ERR_INVALID_PARAMETER;
    const int luma_ctas = (nmb * 16 + 127) / 128, chroma_ctas = (nmb * 8 + 127) / 128;
    k_tq_recon<<<dim3(luma_ctas + chroma_ctas, n_pics), 128, 0, (cudaStream_t)cuda_stream>>>(d_src_y, d_src_u, d_src_v, d_pred_y, d_pred_u, d_pred_v, width, height, mbw, nmb, qp,
                                                                                             host_chroma_qp(qp, chroma_qp_index_offset), d_coeffs, d_recon_y, d_recon_u, d_recon_v,
                                                                                             frame_stride, luma_ctas, host_rcp32(mbw, width, height));
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}
int hlb200_dev_tq_recon(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                        const uint8_t* d_pred_v, int width, int height, int qp, int chroma_qp_index_offset, hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y,
                        uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream)
{
    return hlb200_dev_tq_recon_batch(d_src_y, d_src_u, d_src_v, d_pred_y, d_pred_u, d_pred_v, width, height, 1, 0, qp, chroma_qp_index_offset, d_coeffs, d_recon_y, d_recon_u,
                                     d_recon_v, cuda_stream);
}

static int launch_svc(bool bl, const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_ref_y, const uint8_t* d_ref_u, const uint8_t* d_ref_v,
                      int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* d_motion,
                      hlb200_svc_mb_state_t* d_state, hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream)
{
    if (!d_src_y || !d_src_u || !d_src_v || !d_ref_y || !d_ref_u || !d_ref_v || (!bl && !d_motion) || !d_state || !d_coeffs || !d_recon_y || !d_recon_u || !d_recon_v ||
        width < 16 || height < 16 || (width & 15) || (height & 15) || qp < 0 || qp > 51 || n_pics < 1 || n_pics > 65535)
        return HLB200_ERR_INVALID_PARAMETER;
    const int mbw = width >> 4, nmb = mbw * (height >> 4), qpc = host_chroma_qp(qp, chroma_qp_index_offset);
    SvcPlanes P;
    P.src_y = d_src_y; P.src_u = d_src_u; P.src_v = d_src_v; P.ref_y = d_ref_y; P.ref_u = d_ref_u; P.ref_v = d_ref_v;
    P.rec_y = d_recon_y; P.rec_u = d_recon_u; P.rec_v = d_recon_v; P.W = width; P.H = height;
    const int luma_ctas = (nmb * 16 + 127) / 128, chroma_ctas = (nmb * 8 + 127) / 128;
    const dim3 grid(luma_ctas + chroma_ctas, n_pics);
    if (bl) k_svc_inter_recon<true><<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(P, mbw, nmb, qp, qpc, nullptr, d_state, d_coeffs, frame_stride, luma_ctas, host_rcp32(mbw, width, height));
    else k_svc_inter_recon<false><<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(P, mbw, nmb, qp, qpc, d_motion, d_state, d_coeffs, frame_stride, luma_ctas, host_rcp32(mbw, width, height));
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}
int hlb200_dev_svc_inter_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_ref_y, const uint8_t* d_ref_u,
                                     const uint8_t* d_ref_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                                     const hlb200_mb_motion_t* d_motion, hlb200_svc_mb_state_t* d_state, hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y,
                                     uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream)
{
    return launch_svc(false, d_src_y, d_src_u, d_src_v, d_ref_y, d_ref_u, d_ref_v, width, height, n_pics, frame_stride, qp, chroma_qp_index_offset, d_motion, d_state, d_coeffs,
                      d_recon_y, d_recon_u, d_recon_v, cuda_stream);
}
int hlb200_dev_svc_bl_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                                  const uint8_t* d_pred_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                                  hlb200_svc_mb_state_t* d_state, hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream)
{
    return launch_svc(true, d_src_y, d_src_u, d_src_v, d_pred_y, d_pred_u, d_pred_v, width, height, n_pics, frame_stride, qp, chroma_qp_index_offset, nullptr, d_state, d_coeffs,
                      d_recon_y, d_recon_u, d_recon_v, cuda_stream);
}

// ---------------- SVC inter-layer motion derivation (hlb_svc_derive.cuh; SURVEY 8f-4, second half) ------------------------------------------------------------------------
// One thread per enhancement-layer macroblock, blockIdx.y = picture.  The reference layer's records are 84 bytes per macroblock and four (dyadic) enhancement macroblocks
// read the same one: served by L1/L2, the kernel is a few microseconds per 1080p picture.  Pass 2 resolves the macroblocks that inherit a prediction (base macroblock
// intra) by walking back over the kinds pass 1 wrote; it is a second launch because the walk crosses CTAs.
__global__ void __launch_bounds__(128) k_svc_derive(const hlb200_svc_base_mb_t* __restrict__ base, SvcDeriveGeom g, int mbw, int nmb, uint8_t* __restrict__ flags,
                                                    hlb200_mb_motion_t* __restrict__ motion, int32_t* __restrict__ status)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= nmb) return;
    const size_t pic = blockIdx.y;
    const int st = svc_derive_pass1(base + pic * g.nref, g, mb, mbw, flags + pic * nmb, motion + pic * nmb);
    if (st) atomicOr(status + pic, st);
}
__global__ void __launch_bounds__(128) k_svc_derive_inherit(int nmb, const uint8_t* __restrict__ flags, hlb200_mb_motion_t* __restrict__ motion, int32_t* __restrict__ status)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= nmb) return;
    const size_t pic = blockIdx.y;
    const int st = svc_derive_pass2(mb, flags + pic * nmb, motion + pic * nmb);
    if (st) atomicOr(status + pic, st);
}
int hlb200_dev_svc_derive_motion_batch(const hlb200_svc_base_mb_t* d_base, const hlb200_svc_layer_geom_t* geom, int width, int height, int n_pics, uint8_t* d_had_parts,
                                       hlb200_mb_motion_t* d_motion, int32_t* d_status, void* cuda_stream)
{
    if (!d_base || !geom || !d_had_parts || !d_motion || !d_status || width < 16 || height < 16 || (width & 15) || (height & 15) || n_pics < 1 || n_pics > 65535)
        return HLB200_ERR_INVALID_PARAMETER;
    SvcDeriveGeom g;
    if (geom->cropping_change ||
        !svc_derive_geom(geom->ref_width, geom->ref_height, geom->scaled_width, geom->scaled_height, geom->left_offset, geom->top_offset, geom->level_idc, geom->restricted, g)) {
        snprintf(g_err, sizeof(g_err), "hlb200: inter-layer motion derivation with CroppingChangeFlag = 1 (or a fixed-point precision the reference itself overflows) is not implemented");
        return HLB200_ERR_NOT_IMPLEMENTED;
    }
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    const dim3 grid((nmb + 127) / 128, n_pics);
    k_svc_derive<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(d_base, g, mbw, nmb, d_had_parts, d_motion, d_status);
    k_svc_derive_inherit<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(nmb, d_had_parts, d_motion, d_status);
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}

// Intra_Base resampling (hlb_svc.cuh: svc_resample_px): one thread = four horizontally adjacent output samples of one plane, stored as one word;
// blockIdx.y = picture, blockIdx.z = plane (Y, Cb, Cr).  Reads of the (four times smaller) reference plane go through the read-only path.
#ifndef HLB_RS_MINB
#define HLB_RS_MINB 6   // 40 registers: +1.3 %
#endif
// One launch per kind of plane (CHROMA = false: the luma planes of all pictures, true: their Cb and Cr planes): block = 64 x 4 threads = 256 x 4 output samples,
// blockIdx.x / .y walk the plane, blockIdx.z = picture (luma) or picture * 2 + plane (chroma) -- no index division anywhere, no empty chroma blocks.
extern "C++" {
template <bool CHROMA>
__global__ void __launch_bounds__(256, HLB_RS_MINB) k_svc_resample_intra(const uint8_t* __restrict__ ref_a, const uint8_t* __restrict__ ref_b, int rw, int rh, uint8_t* __restrict__ out_a,
                                                                         uint8_t* __restrict__ out_b, int w, int h, size_t ref_frame_stride, size_t frame_stride, SvcRsAxis ax, SvcRsAxis ay)
{
    const int x0 = (blockIdx.x * 64 + threadIdx.x) * 4, y = blockIdx.y * 4 + threadIdx.y;
    if (x0 >= w || y >= h) return;
    const int pic = CHROMA ? blockIdx.z >> 1 : blockIdx.z, second = CHROMA ? (blockIdx.z & 1) : 0;
    const uint8_t* ref = (second ? ref_b : ref_a) + (size_t)pic * ref_frame_stride;
    uint8_t* out = (second ? out_b : out_a) + (size_t)pic * frame_stride;
    *reinterpret_cast<uint32_t*>(out + (size_t)y * w + x0) = svc_resample_row4(ref, rw, rh, ax, ay, x0, y, CHROMA);
}
}   // extern "C++"

// host twin of svc_rs_axis (hlb_svc.cuh; the device functions are not callable from host code)
static SvcRsAxis host_rs_axis(int refDim, int scaledDim, int level_idc)
{
    SvcRsAxis a;
    int lg = 0;
    while ((1 << lg) < refDim) ++lg;
    const int shift = level_idc <= 30 ? 16 : 31 - lg;
    a.scale = ((refDim << shift) + (scaledDim >> 1)) / scaledDim;
    a.add = (((refDim * 2) << (shift - 2)) + (scaledDim >> 1)) / scaledDim + (1 << (shift - 5));
    a.shift4 = shift - 4;
    return a;
}
static bool rs_ok(int dim, int level_idc) { return level_idc <= 30 || (dim & (dim - 1)) != 0; }   // host twin of svc_rs_precision_ok
int hlb200_dev_svc_resample_intra_batch(const uint8_t* d_ref_y, const uint8_t* d_ref_u, const uint8_t* d_ref_v, int ref_width, int ref_height, uint8_t* d_pred_y,
                                        uint8_t* d_pred_u, uint8_t* d_pred_v, int width, int height, int n_pics, size_t ref_frame_stride, size_t frame_stride, int level_idc,
                                        void* cuda_stream)
{
    // level_idc > 30 selects another fixed-point precision ((G-43): shift = 31 - ceil(log2(refDim))); when a reference dimension (luma or chroma) is a power of two the
    // reference's int32 `refDim << shift` overflows: no reference behaviour there, refused
    if (!d_ref_y || !d_ref_u || !d_ref_v || !d_pred_y || !d_pred_u || !d_pred_v || ref_width < 16 || ref_height < 16 || (ref_width & 15) || (ref_height & 15) || width < ref_width ||
        height < ref_height || (width & 15) || (height & 15) || width > 8 * ref_width || height > 8 * ref_height || width > 8192 || height > 8192 || n_pics < 1 || n_pics > 65535 ||
        (ref_frame_stride & 3) || (((uintptr_t)d_ref_y | (uintptr_t)d_ref_u | (uintptr_t)d_ref_v) & 3) || level_idc < 0 || !rs_ok(ref_width, level_idc) || !rs_ok(ref_height, level_idc) || !rs_ok(ref_width >> 1, level_idc) || !rs_ok(ref_height >> 1, level_idc) || (frame_stride & 3) || (((uintptr_t)d_pred_y | (uintptr_t)d_pred_u | (uintptr_t)d_pred_v) & 3))
        return HLB200_ERR_INVALID_PARAMETER;
    // per-launch constants ((G-43)..(G-48) hold two integer divisions per axis): computed once on the host
    const SvcRsAxis lx = host_rs_axis(ref_width, width, level_idc), ly = host_rs_axis(ref_height, height, level_idc);
    const SvcRsAxis cx = host_rs_axis(ref_width >> 1, width >> 1, level_idc), cy = host_rs_axis(ref_height >> 1, height >> 1, level_idc);
    const dim3 blk(64, 4);
    for (int p0 = 0; p0 < n_pics; p0 += 32767) {   // gridDim.z <= 65535
        const int np = n_pics - p0 < 32767 ? n_pics - p0 : 32767;
        const size_t ro = (size_t)p0 * ref_frame_stride, oo = (size_t)p0 * frame_stride;
        k_svc_resample_intra<false><<<dim3((width / 4 + 63) / 64, (height + 3) / 4, np), blk, 0, (cudaStream_t)cuda_stream>>>(d_ref_y + ro, d_ref_y + ro, ref_width, ref_height, d_pred_y + oo,
                                                                                                                        d_pred_y + oo, width, height, ref_frame_stride, frame_stride, lx, ly);
        k_svc_resample_intra<true><<<dim3((width / 8 + 63) / 64, (height / 2 + 3) / 4, 2 * np), blk, 0, (cudaStream_t)cuda_stream>>>(d_ref_u + ro, d_ref_v + ro, ref_width >> 1, ref_height >> 1,
                                                                                                                               d_pred_u + oo, d_pred_v + oo, width >> 1, height >> 1,
                                                                                                                               ref_frame_stride, frame_stride, cx, cy);
    }
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}

int hlb200_dev_homogeneity8x8(const uint8_t* d_plane, int width, int height, int32_t* d_out, void* cuda_stream)
{
    if (!d_plane || !d_out || width < 8 || height < 8 || (width & 7) || (height & 7)) return HLB200_ERR_INVALID_PARAMETER;
    const int nb = (width >> 3) * (height >> 3);
    k_homogeneity8x8<<<(nb * 32 + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(d_plane, width, height, d_out);
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}

int hlb200_dev_sad4x4(const uint8_t* d_a, const uint8_t* d_b, int width, int height, int use_satd, int32_t* d_out, void* cuda_stream)
{
    if (!d_a || !d_b || !d_out || (width & 3) || (height & 3) || use_satd < 0 || use_satd > 2) return HLB200_ERR_INVALID_PARAMETER;
    const int nb = (width >> 2) * (height >> 2);
    k_sad4x4<<<(nb + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(d_a, d_b, width, height, use_satd, d_out);
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}

// dependency-free integer work: 8 independent accumulator chains x 4 ops (IADD3 + LOP3 pairs), fully unrolled
__global__ void __launch_bounds__(256) k_int_alu_probe(int iters, uint32_t* __restrict__ sink)
{
    uint32_t a0 = threadIdx.x, a1 = blockIdx.x, a2 = 3, a3 = 5, a4 = 7, a5 = 11, a6 = 13, a7 = 17;
    const uint32_t k = blockDim.x + 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a0 = (a0 + k) ^ a4; a1 = (a1 + k) ^ a5; a2 = (a2 + k) ^ a6; a3 = (a3 + k) ^ a7;
            a4 = (a4 + k) ^ a0; a5 = (a5 + k) ^ a1; a6 = (a6 + k) ^ a2; a7 = (a7 + k) ^ a3;
        }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

int hlb200_dev_int_alu_probe(int blocks, int iters, uint32_t* d_sink, void* cuda_stream, uint64_t* ops_out)
{
    if (blocks <= 0 || iters <= 0 || !d_sink) return HLB200_ERR_INVALID_PARAMETER;
    k_int_alu_probe<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(iters, d_sink);
    HLB_CUDA(cudaGetLastError());
    if (ops_out) *ops_out = (uint64_t)blocks * 256u * (uint64_t)iters * 32u;  // 16 adds + 16 xors per iteration
    return HLB200_OK;
}

// ---- self-test of the packed formulations (hlb_fast.cuh) against the plain ones (hlb_prims.cuh) ON THE DEVICE: checks what the CPU check cannot, i.e. that the
// hardware instructions behind the wrappers (IDP.4A.U8.S8, I2IP.SAT, VIADDMNMX.S16x2.RELU, PRMT, SHF.R.W, VABSDIFF4) do what their C++ twins do ----
__global__ void __launch_bounds__(64) k_selftest_fast(int qp, unsigned seed, int* __restrict__ bad)
{
    __shared__ alignas(16) uint8_t tile[48 * 48 + 16];
    __shared__ hlb::QuantK qk;
    unsigned s = seed * 2654435761u + blockIdx.x * 40503u + 1u;
    const int kind = blockIdx.x & 3;
    for (int i = threadIdx.x; i < 48 * 48 + 16; i += blockDim.x) {
        unsigned r = (s + (unsigned)i) * 1664525u + 1013904223u; r ^= r >> 13; r *= 2246822519u; r ^= r >> 16;
        tile[i] = kind == 0 ? (uint8_t)r : kind == 1 ? (uint8_t)((r & 256) ? (r >> 9) % 34 : r) : kind == 2 ? (uint8_t)((r & 256) ? 255 : 0) : (uint8_t)(((r >> 9) & 7) == 0 ? ((r & 1) ? 255 : 0) : (r & 255));
    }
    if (threadIdx.x == 0) {
        // the constants are computed on the host in the product (frame_ctx_derive); repeated here from the device tables as an independent derivation
        const int r6 = qp % 6, q6 = qp / 6;
        qk.qbits = 15 + q6; qk.f_pos = (1 << qk.qbits) / 6; qk.f_neg = (1 << qk.qbits) - 1 - qk.f_pos;
        for (int c = 0; c < 3; ++c) { qk.mf[c] = hlb::kQuantMF[r6][c]; qk.dq_mul[c] = qp >= 24 ? (16 * hlb::kNormAdjust[r6][c]) << (q6 - 4) : 16 * hlb::kNormAdjust[r6][c]; }
        qk.dq_shift = qp >= 24 ? 0 : 4 - q6; qk.dq_round = qp >= 24 ? 0 : 1 << (3 - q6);
        const int lim = (1 << qk.qbits) - qk.f_pos - 1, t0 = lim / qk.mf[0], t1 = (lim / qk.mf[1]) / 4, t2 = (lim / qk.mf[2]) / 2;
        qk.zero_sad = t0 < t1 ? (t0 < t2 ? t0 : t2) : (t1 < t2 ? t1 : t2);
    }
    __syncthreads();
    unsigned r = (s ^ (threadIdx.x * 7919u)) * 1664525u + 1013904223u;
    int nbad = 0;
    for (int it = 0; it < 8; ++it) {
        r = r * 1664525u + 1013904223u;
        const int tx = 2 + (int)((r >> 8) % 39u), ty = 2 + (int)((r >> 16) % 39u);
        for (int pos = 0; pos < 16; ++pos) {
            const int xf = pos & 3, yf = pos >> 2;
            uint8_t a[16];
            hlb::interp_luma_4x4_unrolled(tile + ty * 48 + tx, 48, xf, yf, a);
            const hlb::Rows4 b = hlb::fast_pred_luma((const uint32_t*)tile, 12, tx, ty, xf, yf);
            bool ok = true;
            for (int i = 0; i < 16; ++i) ok = ok && a[i] == (uint8_t)(b.r[i >> 2] >> (8 * (i & 3)));
            nbad += !ok;
            // trial encode of a source block (another place of the tile, or the prediction plus small noise) against this prediction
            r = r * 1664525u + 1013904223u;
            uint8_t sv[16];
            const int sx = (int)((r >> 8) % 44u), sy = (int)((r >> 16) % 44u), mode = (r >> 28) & 3;
            for (int i = 0; i < 16; ++i) {
                unsigned q = (r + i * 2654435761u); q ^= q >> 15; q *= 2246822519u; q ^= q >> 13;
                sv[i] = mode == 0 ? tile[(sy + (i >> 2)) * 48 + sx + (i & 3)] : (uint8_t)hlb::clip255((int)a[i] + (mode == 1 ? (int)(q % 5u) - 2 : mode == 2 ? (int)(q % 41u) - 20 : ((q & 15) == 0 ? 1 : 0)));
            }
            int m[16], lv[16];
            uint32_t any = 0, mask = 0, ref;
            for (int i = 0; i < 16; ++i) { m[i] = (int)sv[i] - (int)a[i]; any |= (uint32_t)m[i]; }
            if (any) { hlb::fwd_transform4x4(m); hlb::quant4x4_ac(m, qp, false); hlb::zigzag4x4(m, lv); mask = hlb::level_mask16(lv); }
            if (mask == 0) ref = (uint32_t)hlb::sad16(sv, a);
            else {
                const hlb::CavlcInfo ci = hlb::cavlc_block_info16(lv, mask);
                int cc[16];
                hlb::inv_zigzag4x4(lv, cc); hlb::dequant4x4(cc, qp, false); hlb::inv_transform4x4(cc);
                uint8_t rec[16];
                for (int i = 0; i < 16; ++i) rec[i] = (uint8_t)((int)a[i] + cc[i]);
                ref = (uint32_t)hlb::sad16(sv, rec) | ((uint32_t)ci.bits_rest << 12) | ((uint32_t)ci.total_coeff << 22) | ((uint32_t)ci.trailing_ones << 27) | ((uint32_t)(ci.single_ctr & 3) << 29);
            }
            hlb::Rows4 sr;
            for (int q = 0; q < 4; ++q) sr.r[q] = sv[4 * q] | (sv[4 * q + 1] << 8) | (sv[4 * q + 2] << 16) | ((uint32_t)sv[4 * q + 3] << 24);
            const uint32_t got = hlb::fast_trial(sr, b, qk, false), cnt = hlb::fast_trial(sr, b, qk, true);
            nbad += got != ref;
            const int tc = (ref >> 22) & 31, t1 = (ref >> 27) & 3;
            bool cok = (int)((cnt >> 22) & 31) == tc;
            if (tc == 1) cok = cok && (int)((cnt >> 27) & 3) == t1 && (t1 != 1 || ((cnt >> 29) & 3) == ((ref >> 29) & 3));
            nbad += !cok;
        }
    }
    if (nbad) atomicAdd(bad, nbad);
}

// ---- TMA probe: one 64x40 tile of a plane fetched the way the slice kernel does it (descriptor in global memory, mbarrier completion, border fix-up) against plain clamped loads ----
__device__ __forceinline__ void tma_probe_load(uint8_t* tile, unsigned long long* bar, const void* tmap, int x0, int y0, int fence_mode)
{
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(tile);
    if (fence_mode == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(64 * 40) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(tmap), "r"(x0), "r"(y0), "r"(b) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(0u) : "memory");
}
__global__ void __launch_bounds__(32) k_tma_probe(const uint8_t* __restrict__ plane, int W, int H, int x0, int y0, const void* tmap_g, const __grid_constant__ CUtensorMap tmap_p, int mode, int* __restrict__ bad)
{
    __shared__ alignas(128) uint8_t tile[64 * 40];
    __shared__ alignas(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    x0 &= ~15;   // the innermost start coordinate of a TMA tile load must sit on a 16-byte boundary
    if (threadIdx.x == 0) tma_probe_load(tile, &bar, mode == 2 ? (const void*)&tmap_p : tmap_g, x0, y0, mode);
    __syncwarp();
    int nbad = 0;
    for (int i = threadIdx.x; i < 64 * 40; i += 32) {
        const int ty = i / 64, tx = i % 64, y = y0 + ty, x = x0 + tx;
        const int want = (x >= 0 && x < W && y >= 0 && y < H) ? plane[y * W + x] : 0;   // out-of-picture samples arrive as zeros
        nbad += tile[i] != want;
    }
    if (nbad) atomicAdd(bad, nbad);
}

}  // extern "C"
namespace hlb { int encode_tile_map(void* out128, const uint8_t* d_plane, int width, int height); }
extern "C" {
int hlb200_dev_tma_probe(const uint8_t* d_plane, int width, int height, int x0, int y0, int mode, int* mismatches_out)
{
    if (!d_plane || !mismatches_out || (width & 15) || width <= 0 || height <= 0 || mode < 0 || mode > 2) return HLB200_ERR_INVALID_PARAMETER;
    alignas(64) CUtensorMap m;
    int rc = hlb::encode_tile_map(&m, d_plane, width, height);
    if (rc) return rc;
    void* d_map = nullptr;
    int* d_bad = nullptr;
    HLB_CUDA(cudaMalloc(&d_map, sizeof(m)));
    HLB_CUDA(cudaMemcpy(d_map, &m, sizeof(m), cudaMemcpyHostToDevice));
    HLB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    HLB_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    k_tma_probe<<<1, 32>>>(d_plane, width, height, x0, y0, d_map, m, mode, d_bad);
    HLB_CUDA(cudaGetLastError());
    HLB_CUDA(cudaMemcpy(mismatches_out, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    HLB_CUDA(cudaFree(d_bad));
    HLB_CUDA(cudaFree(d_map));
    return HLB200_OK;
}

int hlb200_dev_selftest(int blocks, unsigned seed, int* mismatches_out)
{
    if (blocks <= 0 || !mismatches_out) return HLB200_ERR_INVALID_PARAMETER;
    int* d_bad = nullptr;
    HLB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    HLB_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    for (int qp = 12; qp <= 51; ++qp) k_selftest_fast<<<blocks, 64>>>(qp, seed + (unsigned)qp, d_bad);
    HLB_CUDA(cudaGetLastError());
    HLB_CUDA(cudaMemcpy(mismatches_out, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    HLB_CUDA(cudaFree(d_bad));
    return HLB200_OK;
}

int hlb200_dev_me_cost(const uint8_t* d_src_y, const uint8_t* d_ref_y, int width, int height, int qp, const hlb200_me_cand_t* d_cands, int n, hlb200_me_cost_t* d_out,
                       void* cuda_stream)
{
    if (!d_src_y || !d_ref_y || !d_cands || !d_out || n < 0 || (width & 15) || (height & 15) || qp < 0 || qp > 51) return HLB200_ERR_INVALID_PARAMETER;
    return launch_me_cost(d_src_y, d_ref_y, width, height, qp, d_cands, n, d_out, (cudaStream_t)cuda_stream);
}

}  // extern "C"

namespace hlb {
int launch_me_cost(const uint8_t* d_src, const uint8_t* d_ref, int W, int H, int qp, const hlb200_me_cand_t* d_cands, int n, hlb200_me_cost_t* d_out, cudaStream_t st)
{
    if (n <= 0) return HLB200_OK;
    k_me_cost<<<(n * 16 + 127) / 128, 128, 0, st>>>(d_src, d_ref, W, H, qp, d_cands, n, d_out);
    HLB_CUDA(cudaGetLastError());
    return HLB200_OK;
}
}  // namespace hlb
