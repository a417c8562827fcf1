// hlb_svc_derive.cuh -- SVC inter-layer motion derivation on the device (second half): what the reference computes per enhancement-layer
// macroblock of a P picture with base_mode_flag = 1 before it predicts anything --
//   G.6.1 / G.6.2   reference-layer macroblock and partition of each 4x4 block     utils.c:966-1059, mb.h:313-339
//   G.8.6.1.1       refLayerPartIdc, intraILPredFlag                              utils.c:1677-1711
//   G.8.6.1.2       refIdxILPredL0, mvILPredL0 (vector scaling G-232..G-235)      utils.c:1779-1880
//   G.8.6.1.3       mbTypeILPred / subMbTypeILPred (Tables G-7, G-8, EP slices)   utils.c:1986-2218
//   G.8.4.1         refIdxL0 / mvL0 of every (sub-)macroblock partition           utils.c:1498-1650
// as one thread per enhancement-layer macroblock over a compact copy of the reference layer's macroblock fields (hlb200_svc_base_mb_t), so that a layer
// picture needs nothing from the host but those records.  Scope = what the reference can be run on here: frame macroblocks, P (EP) slices,
// CroppingChangeFlag = 0; both the restricted case (layer.c:143: same size or dyadic, macroblock-aligned offsets) and the general one with its replacement /
// merging steps (G-210..G-215, G-244..G-261), pinned against live runs of the reference with layers scaled 3:2.  The functions are HLB_HD: under nvcc the body of
// k_svc_derive (hlb_batch.cu), as plain C++ the CPU run of tools/emu/svc_emu.cpp that the CPU tier checks against the reference's trace (tag 11 / tag 6).
#pragma once
#include "../../include/hlb200.h"
#include "hlb_prims.cuh"

namespace hlb {

// per-picture constants (host side, once per launch)
struct SvcDeriveGeom {
    int ref_w, ref_h, ref_mbw, nref;   // reference layer: RefLayerPicWidthInSamplesL / HeightInSamplesL, its macroblock pitch and count
    int off_x, off_y;                  // ScaledRefLayerLeftOffset / TopOffset (G-5, G-6; frame pictures)
    int shift_x, shift_y;              // (G-7), (G-8)
    int scale_x, scale_y;              // (G-9), (G-10): position scaling
    int mv_scale_x, mv_scale_y;        // (G-232), (G-233) with dSW = dSH = 0 (CroppingChangeFlag = 0)
    int restricted;                    // RestrictedSpatialResolutionChangeFlag (layer.c:143): 0 = the replacement / merging steps of the general case run
};
// ceil(log2(v)), host side (utils.c:990-991 go through HL_MATH_CEIL(HL_MATH_LOG2()))
inline int svc_derive_ceil_log2(int v) { int n = 0; while ((1 << n) < v) ++n; return n; }
// false when the reference itself has no defined result: level_idc > 30 with a power-of-two reference dimension overflows `refW << shiftX` (as svc_rs_precision_ok)
inline bool svc_derive_geom(int ref_w, int ref_h, int scaled_w, int scaled_h, int off_x, int off_y, int level_idc, int restricted, SvcDeriveGeom& g)
{
    g.restricted = restricted != 0;
    if (ref_w < 16 || ref_h < 16 || (ref_w & 15) || (ref_h & 15) || scaled_w < 1 || scaled_h < 1 || ref_w > 16384 || ref_h > 16384 || scaled_w > 16384 || scaled_h > 16384) return false;
    if (level_idc > 30 && (!(ref_w & (ref_w - 1)) || !(ref_h & (ref_h - 1)))) return false;
    g.ref_w = ref_w; g.ref_h = ref_h; g.ref_mbw = ref_w >> 4; g.nref = (ref_w >> 4) * (ref_h >> 4);
    g.off_x = off_x; g.off_y = off_y;
    g.shift_x = level_idc <= 30 ? 16 : 31 - svc_derive_ceil_log2(ref_w);
    g.shift_y = level_idc <= 30 ? 16 : 31 - svc_derive_ceil_log2(ref_h);
    g.scale_x = (int)((((int64_t)ref_w << g.shift_x) + (scaled_w >> 1)) / scaled_w);
    g.scale_y = (int)((((int64_t)ref_h << g.shift_y) + (scaled_h >> 1)) / scaled_h);
    g.mv_scale_x = (int)((((int64_t)scaled_w << 16) + (ref_w >> 1)) / ref_w);
    g.mv_scale_y = (int)((((int64_t)scaled_h << 16) + (ref_h >> 1)) / ref_h);
    return true;
}

// status bits a picture's derivation can raise (any bit: the picture has no reproduced reference behaviour and is refused)
enum {
    SVC_DERIVE_BAD_REF = 1,        // a 4x4 block maps outside the reference layer, onto a macroblock object the reference would divide by zero on, or intra and inter
                                   // reference macroblocks mix inside one macroblock (cannot happen at the restricted ratios with aligned offsets)
    SVC_DERIVE_UNSUPPORTED = 2,    // a partition the fused kernel is not pinned for: refIdxL0 != 0 or predFlagL0 = 0 (the layer contexts hold ONE reference picture)
    SVC_DERIVE_STALE_PARTS = 4,    // base macroblock intra but the macroblock object still holds partitions of an earlier picture (NumSubMbPart[0] != 0): the reference predicts
                                   // from RefPicList0[-1]
    SVC_DERIVE_NO_PRED_SOURCE = 8  // base macroblock intra and no earlier macroblock of the picture has partitions (the reference codes it against scratch memory of an earlier picture)
};

HLB_HD int svc_derive_sar(int v, int s) { return v >> s; }   // arithmetic, as gcc compiles the reference's `>>` on int32_t

// One enhancement-layer macroblock.  Returns 0 (inter: `out` holds partition layout, refIdxL0, mvL0), 1 (intraILPredFlag = 1: the macroblock becomes I_BL, `out` zeroed)
// or -1 (SVC_DERIVE_BAD_REF).
HLB_HD int svc_derive_mb(const hlb200_svc_base_mb_t* __restrict__ base, const SvcDeriveGeom& g, int mbx, int mby, hlb200_mb_motion_t& out)
{
    int tref[16], mvx[16], mvy[16], n_intra = 0;
    unsigned intra = 0;   // bit b: refLayerPartIdc of 4x4 block b is -1 (reference-layer macroblock intra)
    out.part_mode = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { out.sub_mode[i] = 0; out.ref_idx[i] = 0; }
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int s = 0; s < 4; ++s) out.mv[p][s][0] = out.mv[p][s][1] = 0;
    // G.8.6.1.1 over the sixteen 4x4 blocks: luma location (4x + 1, 4y + 1) (utils.c:1692-1695) -> reference-layer sample (G-11..G-14) -> macroblock (G-15) -> partition
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const int x = b & 3, y = b >> 2;
        const int xC = mbx * 16 + 4 * x + 1, yC = mby * 16 + 4 * y + 1;
        int xr = svc_derive_sar((int)((unsigned)(xC - g.off_x) * (unsigned)g.scale_x + (1u << (g.shift_x - 1))), g.shift_x);
        int yr = svc_derive_sar((int)((unsigned)(yC - g.off_y) * (unsigned)g.scale_y + (1u << (g.shift_y - 1))), g.shift_y);
        xr = xr < g.ref_w - 1 ? xr : g.ref_w - 1;   // (G-13bis)
        yr = yr < g.ref_h - 1 ? yr : g.ref_h - 1;
        if (xr < 0 || yr < 0) return -1;
        const int addr = (yr >> 4) * g.ref_mbw + (xr >> 4);
        if (addr >= g.nref) return -1;
        const hlb200_svc_base_mb_t& B = base[addr];
        if (B.flags & 1) { ++n_intra; intra |= 1u << b; tref[b] = -1; mvx[b] = mvy[b] = 0; continue; }   // refLayerPartIdc = -1 (utils.c:1701-1703)
        const int xB = xr & 15, yB = yr & 15;
        // 6.4.12.4 as mb.h:313-339 evaluates it on the reference layer's macroblock object
        int p = 0, s = 0;
        if (!(B.flags & 2)) {
            if (!B.part_w || !B.part_h) return -1;
            p = (16 / B.part_w) * (yB / B.part_h) + xB / B.part_w;
        }
        if (p > 3) return -1;
        if (B.flags & 4) {
            const int sw = B.sub_w[p], sh = B.sub_h[p];
            if (!sw || !sh) return -1;
            s = (8 / sw) * ((yB & 7) / sh) + (xB & 7) / sw;
            if (s > 3) return -1;
        }
        if (B.pred_flag[p] == 0) { tref[b] = -1; mvx[b] = mvy[b] = 0; }   // (G-216..G-218)
        else {
            tref[b] = B.ref_idx[p];                                         // (G-222), frame macroblocks in both layers
            mvx[b] = svc_derive_sar(B.mv[p][s][0] * g.mv_scale_x + 32768, 16);   // (G-234)
            mvy[b] = svc_derive_sar(B.mv[p][s][1] * g.mv_scale_y + 32768, 16);   // (G-235)
        }
    }
    if (n_intra == 16) return 1;
    if (n_intra && g.restricted) return -1;
    int r00, r01, r10, r11;
    if (!g.restricted) {
        // ---- the general case (RestrictedSpatialResolutionChangeFlag = 0), statement by statement as the reference executes it ----
        // (1) intra 4x4 blocks take the partition of a neighbour inside their 8x8 block, utils.c:1713-1741.  Copying refLayerPartIdc = copying what follows from it
        // (reference index, scaled vector).  Two things differ from G.8.6.1.1 and are reproduced: procI4x4Blk is NOT reset between the 8x8 blocks, and the first test reads
        // `refLayerPartIdc[yO + yS][xO + 1] == -1` (fixed column, equality) where the clause has `[xO + 1 - xS] != -1`.
#define SVC_CP(dst, src) do { tref[dst] = tref[src]; mvx[dst] = mvx[src]; mvy[dst] = mvy[src]; intra = (intra & ~(1u << (dst))) | (((intra >> (src)) & 1u) << (dst)); } while (0)
        unsigned proc4 = 0;   // procI4x4Blk[yS][xS] at bit yS * 2 + xS
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int xO = (q & 1) << 1, yO = (q >> 1) << 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xS = k & 1, yS = k >> 1, b = (yO + yS) * 4 + xO + xS;
                if (!((intra >> b) & 1)) continue;
                proc4 |= 1u << k;
                const int bh = (yO + yS) * 4 + xO + 1 - xS, bv = (yO + 1 - yS) * 4 + xO + xS, bd = (yO + 1 - yS) * 4 + xO + 1 - xS;
                if (!((proc4 >> (yS * 2 + 1 - xS)) & 1) && ((intra >> ((yO + yS) * 4 + xO + 1)) & 1)) SVC_CP(b, bh);          // (G-210) as coded
                else if (!((proc4 >> ((1 - yS) * 2 + xS)) & 1) && !((intra >> bv) & 1)) SVC_CP(b, bv);                          // (G-211)
                else if (!((proc4 >> ((1 - yS) * 2 + 1 - xS)) & 1) && !((intra >> bd) & 1)) SVC_CP(b, bd);                      // (G-212)
            }
        }
        // (2) 8x8 blocks whose upper-left 4x4 block is still intra take a neighbouring 8x8 block's columns / rows / corner, utils.c:1743-1774
        unsigned proc8 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int xP = q & 1, yP = q >> 1;
            if (!((intra >> ((yP << 1) * 4 + (xP << 1))) & 1)) continue;
            proc8 |= 1u << q;
            int mode = 0;
            if (!((proc8 >> (yP * 2 + 1 - xP)) & 1) && !((intra >> ((yP << 1) * 4 + 2 - xP)) & 1)) mode = 1;                      // (G-213)
            else if (!((proc8 >> ((1 - yP) * 2 + xP)) & 1) && !((intra >> ((2 - yP) * 4 + (xP << 1))) & 1)) mode = 2;               // (G-214)
            else if (!((proc8 >> ((1 - yP) * 2 + 1 - xP)) & 1) && !((intra >> ((2 - yP) * 4 + 2 - xP)) & 1)) mode = 3;              // (G-215)
            if (!mode) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xS = k & 1, yS = k >> 1, b = ((yP << 1) + yS) * 4 + (xP << 1) + xS;
                const int src = mode == 1 ? ((yP << 1) + yS) * 4 + 2 - xP : (mode == 2 ? (2 - yP) * 4 + (xP << 1) + xS : (2 - yP) * 4 + 2 - xP);
                SVC_CP(b, src);
            }
        }
#undef SVC_CP
        if (intra) return -1;   // the reference would index its macroblock list with -1 (utils.c:1797-1803)
        // (3) reference index of an 8x8 block = smallest non-negative one of its 4x4 blocks, vectors of blocks with another index replaced, utils.c:1888-1912.  The
        // minimum is taken while the four blocks are walked, so a block is compared with the minimum SO FAR (the clause takes the minimum first).
        int rq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int xP = q & 1, yP = q >> 1;
            int r = tref[(yP << 1) * 4 + (xP << 1)];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xS = k & 1, yS = k >> 1, b = (2 * yP + yS) * 4 + 2 * xP + xS, t = tref[b];
                r = (r >= 0 && t >= 0) ? (r < t ? r : t) : (r > t ? r : t);   // HL_MATH_MIN_POSITIVE (G-245)
                if (t != r) {
                    const int bh = (2 * yP + yS) * 4 + 2 * xP + 1 - xS, bv = (2 * yP + 1 - yS) * 4 + 2 * xP + xS, bd = (2 * yP + 1 - yS) * 4 + 2 * xP + 1 - xS;
                    const int src = tref[bh] == r ? bh : (tref[bv] == r ? bv : bd);   // (G-246) / (G-247) / (G-248)
                    mvx[b] = mvx[src]; mvy[b] = mvy[src];
                }
            }
            rq[q] = r;
        }
        r00 = rq[0]; r01 = rq[1]; r10 = rq[2]; r11 = rq[3];
        // (4) vectors of an 8x8 block that differ by at most one quarter sample are merged, utils.c:1916-1979 (EP: list 0 only)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b0 = ((q >> 1) << 3) | ((q & 1) << 1), b1 = b0 + 1, b2 = b0 + 4, b3 = b0 + 5;
#define SVC_D(a, c) ((mvx[a] > mvx[c] ? mvx[a] - mvx[c] : mvx[c] - mvx[a]) + (mvy[a] > mvy[c] ? mvy[a] - mvy[c] : mvy[c] - mvy[a]))
            const bool d01 = SVC_D(b0, b1) <= 1, d02 = SVC_D(b0, b2) <= 1, d03 = SVC_D(b0, b3) <= 1, d23 = SVC_D(b2, b3) <= 1, d13 = SVC_D(b1, b3) <= 1;
#undef SVC_D
            if (d01 && d02 && d03) {
                const int ax = svc_derive_sar(mvx[b0] + mvx[b1] + mvx[b2] + mvx[b3] + 2, 2), ay = svc_derive_sar(mvy[b0] + mvy[b1] + mvy[b2] + mvy[b3] + 2, 2);   // (G-252)
                mvx[b0] = mvx[b1] = mvx[b2] = mvx[b3] = ax; mvy[b0] = mvy[b1] = mvy[b2] = mvy[b3] = ay;
            } else if (d01 && d23) {
                const int ax = svc_derive_sar(mvx[b0] + mvx[b1] + 1, 1), ay = svc_derive_sar(mvy[b0] + mvy[b1] + 1, 1);   // (G-253)
                const int bx = svc_derive_sar(mvx[b2] + mvx[b3] + 1, 1), by = svc_derive_sar(mvy[b2] + mvy[b3] + 1, 1);   // (G-254)
                mvx[b0] = mvx[b1] = ax; mvy[b0] = mvy[b1] = ay; mvx[b2] = mvx[b3] = bx; mvy[b2] = mvy[b3] = by;
            } else if (d02 && d13) {
                const int ax = svc_derive_sar(mvx[b0] + mvx[b2] + 1, 1), ay = svc_derive_sar(mvy[b0] + mvy[b2] + 1, 1);   // (G-255)
                const int bx = svc_derive_sar(mvx[b1] + mvx[b3] + 1, 1), by = svc_derive_sar(mvy[b1] + mvy[b3] + 1, 1);   // (G-256)
                mvx[b0] = mvx[b2] = ax; mvy[b0] = mvy[b2] = ay; mvx[b1] = mvx[b3] = bx; mvy[b1] = mvy[b3] = by;
            }
        }
    } else {
        // refIdxILPredL0 of an 8x8 block = that of its upper-left 4x4 block (utils.c:1888; nothing is merged in the restricted case)
        r00 = tref[0]; r01 = tref[2]; r10 = tref[8]; r11 = tref[10];
    }
    // G.8.6.1.3: partition size from equal reference indices and equal vectors, tested in the reference's order (utils.c:2006-2122)
    bool same_all = true, same_top = true, same_bot = true, same_left = true, same_right = true;
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const int x = b & 3, y = b >> 2;
        const bool e00 = mvx[b] == mvx[0] && mvy[b] == mvy[0];
        same_all = same_all && e00;
        if (y < 2) same_top = same_top && e00; else same_bot = same_bot && mvx[b] == mvx[8] && mvy[b] == mvy[8];
        if (x < 2) same_left = same_left && e00; else same_right = same_right && mvx[b] == mvx[2] && mvy[b] == mvy[2];
    }
    int mode;
    if (r00 == r01 && r00 == r10 && r00 == r11 && same_all) mode = 0;
    else if (r00 == r01 && r10 == r11 && same_top && same_bot) mode = 1;
    else if (r00 == r10 && r01 == r11 && same_left && same_right) mode = 2;
    else mode = 3;
    out.part_mode = (uint8_t)mode;
    if (mode == 3) {
        // sub-macroblock partition sizes, utils.c:2143-2186.  The reference initialises `subPartitionSize[4] = { 4X4 }`: only element 0 starts as 4x4, the other three as
        // enum value 0 = 8x8 -- that is what a block keeps when none of the three tests matches
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b0 = ((q >> 1) << 3) | ((q & 1) << 1), b1 = b0 + 1, b2 = b0 + 4, b3 = b0 + 5;
            const bool h0 = mvx[b0] == mvx[b1] && mvy[b0] == mvy[b1], h1 = mvx[b2] == mvx[b3] && mvy[b2] == mvy[b3];
            const bool v0 = mvx[b0] == mvx[b2] && mvy[b0] == mvy[b2], v1 = mvx[b1] == mvx[b3] && mvy[b1] == mvy[b3];
            int sm = q == 0 ? 3 : 0;
            if (h0 && v0 && mvx[b0] == mvx[b3] && mvy[b0] == mvy[b3]) sm = 0;
            else if (h0 && h1) sm = 1;
            else if (v0 && v1) sm = 2;
            out.sub_mode[q] = (uint8_t)sm;
        }
    }
    // G.8.4.1 with base_mode_flag = 1 (utils.c:1606-1621): a partition takes the predictors at its upper-left sample
    const int nparts = mode == 0 ? 1 : (mode == 3 ? 4 : 2);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (p >= nparts) break;
        const int xP = mode == 2 ? p * 8 : (mode == 3 ? (p & 1) * 8 : 0), yP = mode == 1 ? p * 8 : (mode == 3 ? (p >> 1) * 8 : 0);
        const int sm = mode == 3 ? out.sub_mode[p] : 0, nsub = sm == 0 ? 1 : (sm == 3 ? 4 : 2);
        const int qi = ((yP >> 3) << 1) | (xP >> 3);
        out.ref_idx[p] = (int8_t)(qi == 0 ? r00 : (qi == 1 ? r01 : (qi == 2 ? r10 : r11)));
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= nsub) break;
            const int xS = sm == 2 ? s * 4 : (sm == 3 ? (s & 1) * 4 : 0), yS = sm == 1 ? s * 4 : (sm == 3 ? (s >> 1) * 4 : 0);
            const int b = (((yP + yS) >> 2) << 2) | ((xP + xS) >> 2);
            out.mv[p][s][0] = (int16_t)mvx[b]; out.mv[p][s][1] = (int16_t)mvy[b];
        }
    }
    return 0;
}

// is the derived macroblock one the fused prediction + residual kernel is pinned for?  Every partition layout down to 4x4 sub-macroblock partitions (the general case
// produces them), as long as every partition predicts from RefPicList0[0] (refIdxL0 = 0, which includes predFlagL0 = 1): the layer contexts hold one reference picture.
HLB_HD bool svc_derive_supported(const hlb200_mb_motion_t& m)
{
    const int nparts = m.part_mode == 0 ? 1 : (m.part_mode == 3 ? 4 : 2);
    for (int p = 0; p < nparts; ++p)
        if (m.ref_idx[p] != 0) return false;
    return true;
}

// Pass 1 of a picture, one call per macroblock: derive, classify, keep the "object holds partitions" flag of the layer's macroblock (NumSubMbPart[0] of the reference's
// persistent macroblock object: written by every inter derivation, mb.c:137-207 / :209-244, never by an intra one).  flags[mb]: bit 0 = that carried flag, bits 1-2 = kind of the
// macroblock in THIS picture (0 inter, 1 base macroblock intra: resolved by pass 2, 2 not reproduced) -- pass 2 only reads them, so the walk needs no other scratch.
HLB_HD int svc_derive_pass1(const hlb200_svc_base_mb_t* __restrict__ base, const SvcDeriveGeom& g, int mb, int mbw, uint8_t* flags, hlb200_mb_motion_t* motion)
{
    hlb200_mb_motion_t m;
    const int r = svc_derive_mb(base, g, mb % mbw, mb / mbw, m);
    int status = 0, had = flags[mb] & 1, kind;
    if (r < 0) { status = SVC_DERIVE_BAD_REF; kind = 2; }
    else if (r == 1) { kind = 1; if (had) status = SVC_DERIVE_STALE_PARTS; }
    else {
        had = 1;
        if (svc_derive_supported(m)) kind = 0;
        else { kind = 2; status = SVC_DERIVE_UNSUPPORTED; }
    }
    flags[mb] = (uint8_t)(had | (kind << 1));
    motion[mb] = m;
    return status;
}
// Pass 2: a macroblock whose base macroblock is intra inherits the prediction of the last macroblock before it (raster order) that has partitions
// (hlb_svc.cuh: SvcPredSrc; host/hlb200_glue.c did this serially while it derived)
HLB_HD int svc_derive_pass2(int mb, const uint8_t* flags, hlb200_mb_motion_t* motion)
{
    if ((flags[mb] >> 1) != 1) return 0;
    int a = mb - 1;
    while (a >= 0 && (flags[a] >> 1) != 0) --a;
    if (a < 0 || a >= 65536) return SVC_DERIVE_NO_PRED_SOURCE;
    motion[mb].pad[0] = 1; motion[mb].pad[1] = (uint8_t)(a & 255); motion[mb].pad[2] = (uint8_t)(a >> 8);
    return 0;
}

}  // namespace hlb
