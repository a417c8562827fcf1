// hlb_cavlc.cuh -- CAVLC serialisation of a macroblock from its decision record: the device-side counterpart of
//   _hl_codec_264_mb_write_no_pcm            source/h264/hl_codec_264_mb.c:543-860      (macroblock layer, 7.3.5)
//   hl_codec_264_residual_write              source/h264/hl_codec_264_residual.c:903-1094 (residual(), 7.3.5.3)
//   hl_codec_264_residual_write_block_cavlc  residual.c:587-901                           (residual_block_cavlc(), 7.3.5.3.2 / 9.2)
//   VLC writers                              source/h264/hl_codec_264_cavlc.c:59-104 (level codes), :652-836 (coeff_token, total_zeros, run_before)
//   Exp-Golomb / me(v)                       include/hartallo/h264/hl_codec_264_bits.h:739-886
// for the syntax the reference's encoder emits: Baseline, CAVLC, frame macroblocks, P and I slices, num_ref_idx_l0_active_minus1 = 0 (no ref_idx is written,
// mb.c:707,800), no 8x8 transform, mb_qp_delta as recorded.  Written against a "sink" (put(value, nbits), MSB first) so that the same code counts bits
// (pass 1) and writes them (pass 2); it also compiles as plain C++ (tools/emu: CPU check against the reference's bitstreams).
// Tables are the values of H.264 Tables 9-4, 9-5, 9-7, 9-8, 9-9 and 9-10 in layouts of our own.
#pragma once
#include "../../include/hlb200.h"
#include "hlb_prims.cuh"

namespace hlb {

// coeff_token code words [vlc class 0..2][TotalCoeff 0..16][TrailingOnes 0..3] (Table 9-5; lengths: kCoeffTokenLen); class 3 (nC >= 8) is 6 bits fixed
HLB_TABLE static const uint8_t kCoeffTokenCode[3][17][4] = {
    {{1, 0, 0, 0}, {5, 1, 0, 0}, {7, 4, 1, 0}, {7, 6, 5, 3}, {7, 6, 5, 3}, {7, 6, 5, 4}, {15, 6, 5, 4}, {11, 14, 5, 4}, {8, 10, 13, 4}, {15, 14, 9, 4}, {11, 10, 13, 12}, {15, 14, 9, 12}, {11, 10, 13, 8}, {15, 1, 9, 12}, {11, 14, 13, 8}, {7, 10, 9, 12}, {4, 6, 5, 8}},
    {{3, 0, 0, 0}, {11, 2, 0, 0}, {7, 7, 3, 0}, {7, 10, 9, 5}, {7, 6, 5, 4}, {4, 6, 5, 6}, {7, 6, 5, 8}, {15, 6, 5, 4}, {11, 14, 13, 4}, {15, 10, 9, 4}, {11, 14, 13, 12}, {8, 10, 9, 8}, {15, 14, 13, 12}, {11, 10, 9, 12}, {7, 11, 6, 8}, {9, 8, 10, 1}, {7, 6, 5, 4}},
    {{15, 0, 0, 0}, {15, 14, 0, 0}, {11, 15, 13, 0}, {8, 12, 14, 12}, {15, 10, 11, 11}, {11, 8, 9, 10}, {9, 14, 13, 9}, {8, 10, 9, 8}, {15, 14, 13, 13}, {11, 14, 10, 12}, {15, 10, 13, 12}, {11, 14, 9, 12}, {8, 10, 13, 8}, {13, 7, 9, 12}, {9, 12, 11, 10}, {5, 8, 7, 6}, {1, 4, 3, 2}}};
// chroma DC coeff_token [TotalCoeff 0..4][TrailingOnes 0..3] (Table 9-5, nC == -1): length, code
HLB_TABLE static const uint8_t kCoeffTokenChromaDC[5][4][2] = {{{2, 1}, {0, 0}, {0, 0}, {0, 0}}, {{6, 7}, {1, 1}, {0, 0}, {0, 0}}, {{6, 4}, {6, 6}, {3, 1}, {0, 0}},
                                                               {{6, 3}, {7, 3}, {7, 2}, {6, 5}}, {{6, 2}, {8, 3}, {8, 2}, {7, 0}}};
// total_zeros code words [TotalCoeff-1][total_zeros] (Tables 9-7, 9-8; lengths: kTotalZerosLen)
HLB_TABLE static const uint8_t kTotalZerosCode[15][16] = {
    {1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1}, {7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0, 0}, {5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0, 0, 0},
    {3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0, 0, 0, 0}, {5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0, 0, 0, 0, 0}, {1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0},
    {1, 1, 5, 4, 3, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {1, 1, 1, 3, 3, 2, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0}, {1, 0, 1, 3, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0},
    {1, 0, 1, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 1, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
HLB_TABLE static const uint8_t kTotalZerosCodeChromaDC[3][4] = {{1, 1, 1, 0}, {1, 1, 0, 0}, {1, 0, 0, 0}};   // Table 9-9 (a); lengths: kTotalZerosLenChromaDC
// run_before code words [min(zerosLeft,7)-1][run_before] (Table 9-10; lengths: kRunBeforeLen)
HLB_TABLE static const uint8_t kRunBeforeCode[7][16] = {{1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                        {3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {3, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                        {3, 2, 3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {3, 0, 1, 3, 2, 5, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                        {7, 6, 5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0}};
// me(v): codeNum of coded_block_pattern for ChromaArrayType 1 (Table 9-4), [cbp][0: Intra_4x4, 1: Inter]
HLB_TABLE static const uint8_t kCbpCodeNum[48][2] = {
    {3, 0},   {29, 2},  {30, 3},  {17, 7},  {31, 4},  {18, 8},  {37, 17}, {8, 13},  {32, 5},  {38, 18}, {19, 9},  {9, 14},  {20, 10}, {10, 15}, {11, 16}, {2, 11},
    {16, 1},  {33, 32}, {34, 33}, {21, 36}, {35, 34}, {22, 37}, {39, 44}, {4, 40},  {36, 35}, {40, 45}, {23, 38}, {5, 41},  {24, 39}, {6, 42},  {7, 43},  {1, 19},
    {41, 6},  {42, 24}, {43, 25}, {25, 20}, {44, 26}, {26, 21}, {46, 46}, {12, 28}, {45, 27}, {47, 47}, {27, 22}, {13, 29}, {28, 23}, {14, 30}, {15, 31}, {0, 12}};

// ---- sinks ----
struct BitCount {   // pass 1: only the length
    uint32_t n;
    HLB_HD void put(uint32_t, int nbits) { n += (uint32_t)nbits; }
};
// pass 2: MSB-first bits appended at bit position `pos` of a zero-initialised word buffer shared with other writers (neighbouring macroblocks end / start inside
// the same word): whole words are OR-ed in.  The buffer is big-endian per 32-bit word; the host (or the final kernel) byte-swaps.
struct BitWriter {
    uint32_t* buf;
    unsigned long long pos;   // absolute bit position
    uint64_t acc;             // pending bits, right-aligned
    int nacc;
    HLB_HD void flush_word(uint32_t word, unsigned long long at_bit)
    {
        // `word` holds 32 bits starting at absolute bit at_bit (which may not be word aligned)
        const unsigned long long wi = at_bit >> 5;
        const int sh = (int)(at_bit & 31);
#if defined(__CUDA_ARCH__)
        atomicOr(buf + wi, word >> sh);
        if (sh) atomicOr(buf + wi + 1, word << (32 - sh));
#else
        buf[wi] |= word >> sh;
        if (sh) buf[wi + 1] |= word << (32 - sh);
#endif
    }
    HLB_HD void put(uint32_t v, int nbits)
    {
        if (nbits <= 0) return;
        acc = (acc << nbits) | (uint64_t)(nbits < 32 ? (v & ((1u << nbits) - 1u)) : v);
        nacc += nbits;
        if (nacc >= 32) {
            nacc -= 32;
            flush_word((uint32_t)(acc >> nacc), pos);
            pos += 32;
        }
    }
    HLB_HD void finish()
    {
        if (nacc > 0) { flush_word((uint32_t)(acc << (32 - nacc)), pos); pos += (unsigned)nacc; nacc = 0; }
    }
};

template <class S> HLB_HD void put_ue(S& s, uint32_t v)
{
    const uint32_t c = v + 1;
    const int n = 31 - hlb_clz(c);
    if (2 * n + 1 <= 32) s.put(c, 2 * n + 1);   // n zeros, then the n+1 bits of v+1
    else { s.put(0, n); s.put(c, n + 1); }
}
template <class S> HLB_HD void put_se(S& s, int v) { put_ue(s, v <= 0 ? (uint32_t)(-v) << 1 : ((uint32_t)v << 1) - 1); }

// residual_block_cavlc (9.2): lv = levels in scan order (n = maxNumCoeff of the call: 16, 15 or 4), nC >= 0 or -1 (chroma DC).  Returns TotalCoeff.
template <class S> HLB_HD int cavlc_put_block(S& s, const int16_t* lv, int n, int nC)
{
    int nz[16], run[16];
    int tc = 0, t1 = 0, tz = 0, k = -1;
    bool count_t1 = true, seen = false;
#pragma unroll 1
    for (int i = 0; i < 16; ++i) run[i] = 0;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {  // reverse scan (residual.c:757-786)
        const int c = lv[n - 1 - j];
        if (c) {
            nz[tc++] = c;
            seen = true;
            ++k;
            if (count_t1) {
                if (c == 1 || c == -1) { ++t1; count_t1 = (t1 < 3); }
                else count_t1 = false;
            }
        } else if (seen) {
            ++run[k];
            ++tz;
        }
    }
    if (nC >= 8) s.put(tc ? (uint32_t)(((tc - 1) << 2) | t1) : 3u, 6);
    else if (nC >= 0) { const int vlc = nC < 2 ? 0 : (nC < 4 ? 1 : 2); s.put(kCoeffTokenCode[vlc][tc][t1], kCoeffTokenLen[vlc][tc][t1]); }
    else s.put(kCoeffTokenChromaDC[tc][t1][1], kCoeffTokenChromaDC[tc][t1][0]);
    if (tc == 0) return 0;
    int sl = (tc > 10 && t1 < 3) ? 1 : 0;
#pragma unroll 1
    for (int j = 0; j < tc; ++j) {
        const int v = nz[j];
        if (j < t1) { s.put((uint32_t)((1 - v) >> 1), 1); continue; }   // trailing_ones_sign_flag
        int lc = v > 0 ? (v << 1) - 2 : -(v << 1) - 1;
        if (j == t1 && t1 < 3 && lc >= 2) lc -= 2;
        // level_prefix zeros, a one, level_suffix (cavlc.c:59-104: prefix <= 15)
        if (sl == 0) {
            if (lc < 14) s.put(1, lc + 1);
            else if (lc < 30) { s.put(1, 15); s.put((uint32_t)(lc - 14), 4); }
            else { s.put(1, 16); s.put((uint32_t)(lc - 30), 12); }
        } else {
            const int prefix = lc >> sl;
            if (prefix < 15) { s.put(1, prefix + 1); s.put((uint32_t)(lc & ((1 << sl) - 1)), sl); }
            else { s.put(1, 16); s.put((uint32_t)(lc - (15 << sl)), 12); }
        }
        if (sl == 0) sl = 1;
        if (iabs(v) > (3 << (sl - 1)) && sl < 6) ++sl;
    }
    int zl = 0;
    if (tc < n) {
        if (nC >= 0) s.put(kTotalZerosCode[tc - 1][tz], kTotalZerosLen[tc - 1][tz]);
        else s.put(kTotalZerosCodeChromaDC[tc - 1][tz], kTotalZerosLenChromaDC[tc - 1][tz]);
        zl = tz;
    }
#pragma unroll 1
    for (int q = 0; q < tc - 1 && zl > 0; ++q) {
        const int row = (zl > 7 ? 7 : zl) - 1;
        s.put(kRunBeforeCode[row][run[q]], kRunBeforeLen[row][run[q]]);
        zl -= run[q];
    }
    return tc;
}

// What the nC derivation needs to know of a neighbouring macroblock (9.2.1; residual.c:698-755): written by the slice kernel as part of MbState
struct CavlcNb {
    bool avail;
    uint8_t kind, cbp_luma, cbp_chroma;
    const uint8_t* tc_luma;      // [16]
    const uint8_t* tc_cac;       // [2][4]
};
HLB_HD int cavlc_nb_luma(const CavlcNb& nb, int blk)
{
    if (!nb.avail) return -1;
    if (nb.kind == HLB200_MB_P_SKIP || ((nb.cbp_luma >> (blk >> 2)) & 1) == 0) return 0;
    return nb.tc_luma[blk];
}
HLB_HD int cavlc_nb_chroma(const CavlcNb& nb, int c, int blk)
{
    if (!nb.avail) return -1;
    if (nb.kind == HLB200_MB_P_SKIP || (nb.cbp_chroma & 2) == 0) return 0;
    return nb.tc_cac[c * 4 + blk];
}
HLB_HD int cavlc_nc(int nA, int nB)
{
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    if (nA >= 0) return nA;
    if (nB >= 0) return nB;
    return 0;
}

// macroblock_layer() + residual() of one non-skipped macroblock (mb.c:620-860, residual.c:903-1094); the mb_skip_run / slice-level part is the caller's.
// A, B: left / top neighbouring macroblocks as they stand after their own serialisation.
template <class S> HLB_HD void cavlc_put_mb(S& s, const hlb200_mb_record_t& r, const CavlcNb& A, const CavlcNb& B)
{
    const bool i16 = r.mb_class == HLB200_MB_I16x16, i4 = r.mb_class == HLB200_MB_I4x4;
    put_ue(s, r.mb_type);
    if (i4) {
#pragma unroll 1
        for (int b = 0; b < 16; ++b) {
            s.put(r.prev_intra4x4_pred_mode_flag[b] ? 1u : 0u, 1);
            if (!r.prev_intra4x4_pred_mode_flag[b]) s.put(r.rem_intra4x4_pred_mode[b], 3);
        }
    }
    if (i4 || i16) put_ue(s, r.intra_chroma_pred_mode);
    else if (r.part_mode == 3) {   // P_8x8ref0: sub_mb_type x 4, then the vectors of every sub-partition (no ref_idx)
#pragma unroll 1
        for (int p = 0; p < 4; ++p) put_ue(s, r.sub_mode[p]);
#pragma unroll 1
        for (int p = 0; p < 4; ++p) {
            const int ns = r.sub_mode[p] == 0 ? 1 : (r.sub_mode[p] == 3 ? 4 : 2);
#pragma unroll 1
            for (int q = 0; q < ns; ++q) { put_se(s, r.mvd[p][q][0]); put_se(s, r.mvd[p][q][1]); }
        }
    } else {
        const int np = r.part_mode == 0 ? 1 : 2;
#pragma unroll 1
        for (int p = 0; p < np; ++p) { put_se(s, r.mvd[p][0][0]); put_se(s, r.mvd[p][0][1]); }
    }
    if (!i16) put_ue(s, kCbpCodeNum[r.coded_block_pattern][i4 ? 0 : 1]);
    if (!(r.cbp_luma > 0 || r.cbp_chroma > 0 || i16)) return;
    put_se(s, r.mb_qp_delta);
    // ---- residual(0, 15) ----
    uint8_t tc[16];   // TotalCoeffsLuma of this macroblock as the writer leaves them block by block
#pragma unroll 1
    for (int b = 0; b < 16; ++b) tc[b] = 0;
    if (i16) {
        const int nC = cavlc_nc(cavlc_nb_luma(A, 5), cavlc_nb_luma(B, 10));
        tc[0] = (uint8_t)cavlc_put_block(s, r.i16_dc_level, 16, nC);
    }
#pragma unroll 1
    for (int b = 0; b < 16; ++b) {
        if (!((r.cbp_luma >> (b >> 2)) & 1)) continue;
        const int x = blk_x(b), y = blk_y(b);
        // in-macroblock neighbours are gated by this macroblock's own CodedBlockPatternLuma (utils.h:10-20)
        int nA, nB;
        if (x > 0) { const int a = blk_idx_from_xy(x - 4, y); nA = ((r.cbp_luma >> (a >> 2)) & 1) ? tc[a] : 0; }
        else nA = cavlc_nb_luma(A, blk_idx_from_xy(12, y));
        if (y > 0) { const int t = blk_idx_from_xy(x, y - 4); nB = ((r.cbp_luma >> (t >> 2)) & 1) ? tc[t] : 0; }
        else nB = cavlc_nb_luma(B, blk_idx_from_xy(x, 12));
        tc[b] = (uint8_t)(i16 ? cavlc_put_block(s, r.i16_ac_level[b], 15, cavlc_nc(nA, nB)) : cavlc_put_block(s, r.luma_level[b], 16, cavlc_nc(nA, nB)));
    }
    if (r.cbp_chroma & 3) {
        int16_t z[16];
#pragma unroll 1
        for (int i = 0; i < 16; ++i) z[i] = 0;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) cavlc_put_block(s, r.cbp_chroma_dc4x4[c] ? r.chroma_dc_level[c] : z, 4, -1);
        if (r.cbp_chroma & 2) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint8_t tcc[4] = {0, 0, 0, 0};
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    // 6.4.10.5: left / top 4x4 chroma blocks; inside the macroblock the gate is this macroblock's CodedBlockPatternChroma & 2 (set here)
                    const int nA = (b & 1) ? tcc[b - 1] : cavlc_nb_chroma(A, c, b + 1), nB = (b & 2) ? tcc[b - 2] : cavlc_nb_chroma(B, c, b + 2);
                    tcc[b] = (uint8_t)cavlc_put_block(s, ((r.cbp_chroma_ac4x4[c] >> b) & 1) ? r.chroma_ac_level[c][b] : z, 15, cavlc_nc(nA, nB));
                }
            }
        }
    }
}

}  // namespace hlb
