// hlb_prims.cuh -- per-4x4-block integer primitives of the H.264 pixel hot path, written for one CUDA thread per
// 4x4 block with the whole block in registers (the natural unit of the reference, see below), bit-exact with the
// reference's C path.  Every function is `HLB_HD` (device-only under nvcc) and the same source also compiles as plain C++ for the CPU
// emulation harness under tests/emu (a debugging aid -- never part of the shipped library).
//
// Reference behaviour followed (file:line under the reference tree):
//   luma interpolation   source/h264/hl_codec_264_pred_inter.c:339-885, include/hartallo/h264/hl_codec_264_interpol.h:41-923
//   chroma interpolation source/h264/hl_codec_264_interpol.c:337-385, pred_inter.c:888-940
//   forward transform    source/h264/hl_codec_264_transf.c:716-768     Hadamards transf.c:774-868
//   quantisation         source/h264/hl_codec_264_quant.c:116-189      dequant quant.c:68-111
//   inverse transform    source/h264/hl_codec_264_transf.c:420-456     DC scaling transf.c:498-608, 612-700
//   SAD/SATD             source/hl_math.c:239, :283                    add+clip include/hartallo/hl_math.h:261-323
//   CAVLC bit length     source/h264/hl_codec_264_residual.c:757-898, source/h264/hl_codec_264_cavlc.c:59-104,652-836
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HLB_HD __device__ __forceinline__
#define HLB_FN __device__ __noinline__   /* phase / control functions of the slice kernel: one copy each */
#ifndef HLB_CAVLC_FN
#define HLB_CAVLC_FN __device__ __forceinline__   /* hlb_slice.cu overrides: one out-of-line copy */
#endif
#ifndef HLB_INTERP_FN
#define HLB_INTERP_FN __device__ __forceinline__
#define HLB_INTERP_SRC(g) ((void)0)
#endif
#define HLB_TABLE __constant__   /* small read-only tables: constant cache (plain global loads are compiled .cg and would go to L2) */
#define HLB_LDG(p) __ldg(p)   /* read-only for the whole kernel (reference / source planes): non-coherent path, L1-cacheable */
#define HLB_LDCG(p) __ldcg(p)   /* written by other CTAs of the same launch (macroblock state, reconstruction): served from L2, never from a stale L1 line */
#if defined(__CUDA_ARCH__)
#define HLB_IN_SHARED(ref) __builtin_assume(__isShared(&(ref)))   /* lets nvcc emit LDS/STS instead of generic accesses in non-inlined functions */
#else
#define HLB_IN_SHARED(ref) ((void)0)
#endif
#else
#define HLB_HD inline
#define HLB_FN inline
#define HLB_CAVLC_FN inline
#define HLB_INTERP_FN inline
#define HLB_INTERP_SRC(g) ((void)0)
#define HLB_TABLE
#define HLB_LDG(p) (*(p))
#define HLB_LDCG(p) (*(p))
#define HLB_IN_SHARED(ref) ((void)0)
#endif

namespace hlb {

// ------------------------------------------------------------------------------------------------------------------
// Constant tables (H.264 spec values; layouts are ours)
// ------------------------------------------------------------------------------------------------------------------
// forward-quant multiplier by (QP%6, position class): class 0 = (even,even), 1 = (odd,odd), 2 = mixed
HLB_TABLE static const int32_t kQuantMF[6][3] = {
    {13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554}, {9362, 3647, 5825}, {8192, 3355, 5243}, {7282, 2893, 4559}};
// dequant normAdjust4x4 by (QP%6, class); LevelScale4x4 = 16 (flat weight) * this
HLB_TABLE static const int32_t kNormAdjust[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
// zig-zag (frame) scan: raster index (y*4+x) of the k-th coefficient
HLB_TABLE static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
// chroma QP from qPI (Table 8-15)
HLB_TABLE static const uint8_t kQpc[52] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                                            18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33,
                                            34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
// coeff_token length [vlc class 0..2][TotalCoeff 0..16][TrailingOnes 0..3] (Table 9-5); class 3 (nC>=8) is 6 bits flat
HLB_TABLE static const uint8_t kCoeffTokenLen[3][17][4] = {
    {{1, 0, 0, 0},   {6, 2, 0, 0},    {8, 6, 3, 0},    {9, 8, 7, 5},    {10, 9, 8, 6},   {11, 10, 9, 7},
     {13, 11, 10, 8}, {13, 13, 11, 9}, {13, 13, 13, 10}, {14, 14, 13, 11}, {14, 14, 14, 13}, {15, 15, 14, 14},
     {15, 15, 15, 14}, {16, 15, 15, 15}, {16, 16, 16, 15}, {16, 16, 16, 16}, {16, 16, 16, 16}},
    {{2, 0, 0, 0},   {6, 2, 0, 0},    {6, 5, 3, 0},    {7, 6, 6, 4},    {8, 6, 6, 4},    {8, 7, 7, 5},
     {9, 8, 8, 6},   {11, 9, 9, 6},   {11, 11, 11, 7}, {12, 11, 11, 9}, {12, 12, 12, 11}, {12, 12, 12, 11},
     {13, 13, 13, 12}, {13, 13, 13, 13}, {13, 14, 13, 13}, {14, 14, 14, 13}, {14, 14, 14, 14}},
    {{4, 0, 0, 0},   {6, 4, 0, 0},    {6, 5, 4, 0},    {6, 5, 5, 4},    {7, 5, 5, 4},    {7, 5, 5, 4},
     {7, 6, 6, 4},   {7, 6, 6, 4},    {8, 7, 7, 5},    {8, 8, 7, 6},    {9, 8, 8, 7},    {9, 9, 8, 8},
     {9, 9, 9, 8},   {10, 9, 9, 9},   {10, 10, 10, 10}, {10, 10, 10, 10}, {10, 10, 10, 10}}};
// chroma DC coeff_token length [TotalCoeff 0..4][TrailingOnes 0..3]
HLB_TABLE static const uint8_t kCoeffTokenLenChromaDC[5][4] = {{2, 0, 0, 0}, {6, 1, 0, 0}, {6, 6, 3, 0}, {6, 7, 7, 6}, {6, 8, 8, 7}};
// total_zeros length [TotalCoeff-1][total_zeros] (Tables 9-7, 9-8)
HLB_TABLE static const uint8_t kTotalZerosLen[15][16] = {
    {1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9}, {3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6, 0},
    {4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6, 0, 0}, {5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5, 0, 0, 0},
    {4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5, 0, 0, 0, 0}, {6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6, 0, 0, 0, 0, 0},
    {6, 5, 3, 3, 3, 2, 3, 4, 3, 6, 0, 0, 0, 0, 0, 0}, {6, 4, 5, 3, 2, 2, 3, 3, 6, 0, 0, 0, 0, 0, 0, 0},
    {6, 6, 4, 2, 2, 3, 2, 5, 0, 0, 0, 0, 0, 0, 0, 0}, {5, 5, 3, 2, 2, 2, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {4, 4, 3, 3, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {4, 4, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {3, 3, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {2, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
HLB_TABLE static const uint8_t kTotalZerosLenChromaDC[3][4] = {{1, 2, 3, 3}, {1, 2, 2, 0}, {1, 1, 0, 0}};
// run_before length [min(zerosLeft,7)-1][run_before] (Table 9-10)
HLB_TABLE static const uint8_t kRunBeforeLen[7][16] = {{1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {1, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                       {2, 2, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {2, 2, 2, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                       {2, 2, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {2, 3, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                                       {3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0}};

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
HLB_HD int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
HLB_HD int iabs(int v) { return v < 0 ? -v : v; }
HLB_HD int tap6(int E, int F, int G, int H, int I, int J) { return E - 5 * (F + I) + 20 * (G + H) + J; }
HLB_HD int pos_class(int i, int j) { return ((i | j) & 1) == 0 ? 0 : (((i & j) & 1) ? 1 : 2); }
// upper-left luma sample of 4x4 block blkIdx inside the MB (6.4.3)
HLB_HD int blk_x(int b) { return ((b >> 2) & 1) * 8 + (b & 1) * 4; }
HLB_HD int blk_y(int b) { return (b >> 3) * 8 + ((b >> 1) & 1) * 4; }
HLB_HD int blk_idx_from_xy(int x, int y) { return ((y >> 3) << 3) | ((x >> 3) << 2) | (((y >> 2) & 1) << 1) | ((x >> 2) & 1); }
// se(v)/ue(v) Exp-Golomb lengths (include/hartallo/h264/hl_codec_264_bits.h:739-803)
HLB_HD int ue_len(uint32_t k)
{
    const uint32_t v = k + 1;   // 2 * floor(log2(k + 1)) + 1
#if defined(__CUDA_ARCH__)
    return 2 * (31 - __clz((int)v)) + 1;
#else
    return 2 * (31 - (v ? __builtin_clz(v) : 32)) + 1;
#endif
}
HLB_HD int se_len(int v) { return ue_len(v <= 0 ? (uint32_t)(-v) << 1 : ((uint32_t)v << 1) - 1); }

// ------------------------------------------------------------------------------------------------------------------
// Luma fractional-sample interpolation of one 4x4 block (8.4.2.2.1, Table 8-12).
// `g` points at integer sample G of output pixel (0,0) inside a u8 tile with row pitch `pitch`; the tile must hold
// rows/cols -2..+6 around g (the caller has already applied the reference's origin clip + per-sample clamp when it
// filled the tile).  out[16] raster.
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int hl_h(const uint8_t* p) { return tap6(p[-2], p[-1], p[0], p[1], p[2], p[3]); }                       // unrounded b
HLB_HD int hl_v(const uint8_t* p, int s) { return tap6(p[-2 * s], p[-s], p[0], p[s], p[2 * s], p[3 * s]); }  // unrounded h
HLB_HD int rnd5(int v) { return clip255((v + 16) >> 5); }

// Compact formulation: three paths only (integer copy; positions built from the horizontal / vertical half samples b, h;
// positions involving the centre half sample j), rows produced by small loops that are NOT unrolled.  The slice kernel runs
// hundreds of CTAs that sit in different places of the code: short loops keep the instruction-cache footprint (and the
// divergence between candidates with different fractional positions) small.  Same arithmetic as interpol.h:162-923.
struct Rows4 { uint32_t r[4]; };   // four rows of four samples, one byte each
HLB_INTERP_FN Rows4 interp_luma_4x4_rows(const uint8_t* g, int pitch, int xf, int yf)
{
    HLB_INTERP_SRC(g);
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;   // rows, shifted in as they are produced (no dynamic register indexing)
    if ((xf | yf) == 0) {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const uint8_t* p = g + r * pitch;
            o0 = o1; o1 = o2; o2 = o3;
            o3 = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        }
    } else if (xf != 2 && yf != 2 ? true : (xf == 0 || yf == 0)) {
        // a b c d h n (one half sample, possibly averaged with an integer sample) and e g p r (average of b and h)
        const bool needB = xf != 0, needH = yf != 0;
        const int rowB = yf == 3 ? pitch : 0, colH = xf == 3 ? 1 : 0;
        const int gofs = needB ? (xf == 3 ? 1 : 0) : (yf == 3 ? pitch : 0);   // integer sample averaged with a lone half sample
        const bool lone_half = needB != needH && (needB ? xf == 2 : yf == 2);
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const uint8_t* p = g + r * pitch;
            uint32_t row = 0;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int bb = rnd5(hl_h(p + rowB + x)), hh = rnd5(hl_v(p + x + colH, pitch));
                int v;
                if (needB && needH) v = (bb + hh + 1) >> 1;
                else {
                    const int hs = needB ? bb : hh;
                    v = lone_half ? hs : ((p[x + gofs] + hs + 1) >> 1);
                }
                row |= (uint32_t)v << (8 * x);
            }
            o0 = o1; o1 = o2; o2 = o3; o3 = row;
        }
    } else {
        // j and f q (xf == 2: partner b of row y / y+1) or i k (yf == 2: partner h of column x / x+1).  Streaming over the nine
        // rows -2..+6 of unrounded horizontal half samples with a six-row window.
        int w0[4], w1[4], w2[4], w3[4], w4[4], w5[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) w0[x] = w1[x] = w2[x] = w3[x] = w4[x] = w5[x] = 0;
        const int colH = xf == 3 ? 1 : 0;
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            const uint8_t* p = g + (r - 2) * pitch;
#pragma unroll
            for (int x = 0; x < 4; ++x) { w0[x] = w1[x]; w1[x] = w2[x]; w2[x] = w3[x]; w3[x] = w4[x]; w4[x] = w5[x]; w5[x] = hl_h(p + x); }
            if (r >= 5) {   // window = rows y-2 .. y+3 of output row y = r - 5
                const uint8_t* q = g + (r - 5) * pitch;
                uint32_t row = 0;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int j = clip255((tap6(w0[x], w1[x], w2[x], w3[x], w4[x], w5[x]) + 512) >> 10);
                    int v = j;
                    if (xf == 2) { if (yf != 2) v = (rnd5(yf == 3 ? w3[x] : w2[x]) + j + 1) >> 1; }
                    else v = (rnd5(hl_v(q + x + colH, pitch)) + j + 1) >> 1;
                    row |= (uint32_t)v << (8 * x);
                }
                o0 = o1; o1 = o2; o2 = o3; o3 = row;
            }
        }
    }
    Rows4 o;
    o.r[0] = o0; o.r[1] = o1; o.r[2] = o2; o.r[3] = o3;
    return o;
}
HLB_HD void interp_luma_4x4(const uint8_t* g, int pitch, int xf, int yf, uint8_t out[16])
{
    const Rows4 o = interp_luma_4x4_rows(g, pitch, xf, yf);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        out[x] = (uint8_t)(o.r[0] >> (8 * x)); out[4 + x] = (uint8_t)(o.r[1] >> (8 * x)); out[8 + x] = (uint8_t)(o.r[2] >> (8 * x)); out[12 + x] = (uint8_t)(o.r[3] >> (8 * x));
    }
}

// the original per-position formulation (kept as the cross-check of the compact one in tools/emu/check_interp.cpp)
HLB_HD void interp_luma_4x4_unrolled(const uint8_t* g, int pitch, int xf, int yf, uint8_t out[16])
{
    if (xf == 0 && yf == 0) {
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) out[y * 4 + x] = g[y * pitch + x];
        return;
    }
    if (yf == 0) {  // a, b, c
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const uint8_t* p = g + y * pitch + x;
                int b = rnd5(hl_h(p));
                out[y * 4 + x] = (uint8_t)(xf == 2 ? b : ((p[xf == 1 ? 0 : 1] + b + 1) >> 1));
            }
        return;
    }
    if (xf == 0) {  // d, h, n
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const uint8_t* p = g + y * pitch + x;
                int h = rnd5(hl_v(p, pitch));
                out[y * 4 + x] = (uint8_t)(yf == 2 ? h : ((p[yf == 1 ? 0 : pitch] + h + 1) >> 1));
            }
        return;
    }
    if ((xf & 1) && (yf & 1)) {  // e, g, p, r: average of a horizontal half (row y or y+1) and a vertical half (col x or x+1)
        const int hy = (yf == 3) ? 1 : 0, vx = (xf == 3) ? 1 : 0;
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const uint8_t* p = g + y * pitch + x;
                int b = rnd5(hl_h(p + hy * pitch));
                int h = rnd5(hl_v(p + vx, pitch));
                out[y * 4 + x] = (uint8_t)((b + h + 1) >> 1);
            }
        return;
    }
    // j and its neighbours f, q (xf==2) / i, k (yf==2): j = clip((tap6 over unrounded horizontal tap6 rows + 512) >> 10)
    int b1[9][4];  // unrounded horizontal half samples for rows -2..6
#pragma unroll
    for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int x = 0; x < 4; ++x) b1[r][x] = hl_h(g + (r - 2) * pitch + x);
    if (xf == 2) {
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                int j = clip255((tap6(b1[y][x], b1[y + 1][x], b1[y + 2][x], b1[y + 3][x], b1[y + 4][x], b1[y + 5][x]) + 512) >> 10);
                if (yf == 2) out[y * 4 + x] = (uint8_t)j;
                else {
                    int b = rnd5(b1[y + 2 + (yf == 3 ? 1 : 0)][x]);  // b (row y) for f, s (row y+1) for q
                    out[y * 4 + x] = (uint8_t)((b + j + 1) >> 1);
                }
            }
    } else {  // yf == 2, xf in {1,3}: i = (h + j + 1)>>1, k = (j + m + 1)>>1 with m = vertical half at column x+1
        const int vx = (xf == 3) ? 1 : 0;
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                int j = clip255((tap6(b1[y][x], b1[y + 1][x], b1[y + 2][x], b1[y + 3][x], b1[y + 4][x], b1[y + 5][x]) + 512) >> 10);
                int h = rnd5(hl_v(g + y * pitch + x + vx, pitch));
                out[y * 4 + x] = (uint8_t)((h + j + 1) >> 1);
            }
    }
}

// Chroma 1/8-pel bilinear sample (8.4.2.2.2): A,B,C,D are the four neighbours (already clamped by the caller)
HLB_HD int interp_chroma_px(int A, int B, int C, int D, int xf, int yf)
{
    return ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6;
}

// ------------------------------------------------------------------------------------------------------------------
// 4x4 transforms, in place on raster int[16]
// ------------------------------------------------------------------------------------------------------------------
HLB_HD void fwd_transform4x4(int m[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // columns: Cf * X
        int a = m[j], b = m[4 + j], c = m[8 + j], d = m[12 + j];
        m[j] = a + b + c + d;
        m[4 + j] = 2 * a + b - c - 2 * d;
        m[8 + j] = a - b - c + d;
        m[12 + j] = a - 2 * b + 2 * c - d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // rows: (.) * Cf^T
        int a = m[4 * i], b = m[4 * i + 1], c = m[4 * i + 2], d = m[4 * i + 3];
        m[4 * i] = a + b + c + d;
        m[4 * i + 1] = 2 * a + b - c - 2 * d;
        m[4 * i + 2] = a - b - c + d;
        m[4 * i + 3] = a - 2 * b + 2 * c - d;
    }
}

// 8.5.12.2 (transf.c:420): rows first, then columns, then (x + 32) >> 6
HLB_HD void inv_transform4x4(int d[16])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int d0 = d[4 * i], d1 = d[4 * i + 1], d2 = d[4 * i + 2], d3 = d[4 * i + 3];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        d[4 * i] = e0 + e3; d[4 * i + 1] = e1 + e2; d[4 * i + 2] = e1 - e2; d[4 * i + 3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int f0 = d[j], f1 = d[4 + j], f2 = d[8 + j], f3 = d[12 + j];
        int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        d[j] = (g0 + g3 + 32) >> 6; d[4 + j] = (g1 + g2 + 32) >> 6; d[8 + j] = (g1 - g2 + 32) >> 6; d[12 + j] = (g0 - g3 + 32) >> 6;
    }
}

// 4x4 Hadamard, no scaling (used by the luma-DC forward (then >>1), the luma-DC inverse and SATD)
HLB_HD void hadamard4x4(int m[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int a = m[j], b = m[4 + j], c = m[8 + j], d = m[12 + j];
        m[j] = a + b + c + d; m[4 + j] = a + b - c - d; m[8 + j] = a - b - c + d; m[12 + j] = a - b + c - d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int a = m[4 * i], b = m[4 * i + 1], c = m[4 * i + 2], d = m[4 * i + 3];
        m[4 * i] = a + b + c + d; m[4 * i + 1] = a + b - c - d; m[4 * i + 2] = a - b - c + d; m[4 * i + 3] = a - b + c - d;
    }
}
HLB_HD void hadamard2x2(int m[4])
{
    int a = m[0] + m[2], b = m[1] + m[3], c = m[0] - m[2], d = m[1] - m[3];
    m[0] = a + b; m[1] = a - b; m[2] = c + d; m[3] = c - d;
}

// ------------------------------------------------------------------------------------------------------------------
// Quantisation (quant.c:116-189).  qbits = 15 + QP/6, f = (1<<qbits)/6 (inter) or /3 (intra)
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int quant_f(int qp, bool intra) { return (1 << (15 + qp / 6)) / (intra ? 3 : 6); }

HLB_HD void quant4x4_ac(int m[16], int qp, bool intra)
{
    const int qbits = 15 + qp / 6, f = quant_f(qp, intra), r = qp % 6;
    const int mf0 = kQuantMF[r][0], mf1 = kQuantMF[r][1], mf2 = kQuantMF[r][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = pos_class(i, j);
            const int mf = c == 0 ? mf0 : (c == 1 ? mf1 : mf2);
            int w = m[i * 4 + j];
            int z = (iabs(w) * mf + f) >> qbits;
            m[i * 4 + j] = w >= 0 ? z : -z;
        }
}
// luma DC (n = 16) and chroma DC (n = 4): (|Y| * MF00 + 2f) >> (qbits + 1)
HLB_HD void quant_dc(int* m, int n, int qp, bool intra)
{
    const int qbits1 = 16 + qp / 6, f2 = quant_f(qp, intra) << 1, mf = kQuantMF[qp % 6][0];
    for (int i = 0; i < n; ++i) {
        int w = m[i];
        int z = (iabs(w) * mf + f2) >> qbits1;
        m[i] = w >= 0 ? z : -z;
    }
}

// 8.5.12.1 (quant.c:68): flat scaling lists => LevelScale = 16 * normAdjust.  keep_dc: d00 = c00 (Intra16x16 / chroma)
HLB_HD void dequant4x4(int c[16], int qp, bool keep_dc)
{
    const int r = qp % 6, q6 = qp / 6;
    const int c00 = c[0];
    const int ls0 = 16 * kNormAdjust[r][0], ls1 = 16 * kNormAdjust[r][1], ls2 = 16 * kNormAdjust[r][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = pos_class(i, j);
            const int ls = k == 0 ? ls0 : (k == 1 ? ls1 : ls2);
            int v = c[i * 4 + j] * ls;
            c[i * 4 + j] = qp >= 24 ? (v << (q6 - 4)) : ((v + (1 << (3 - q6))) >> (4 - q6));
        }
    if (keep_dc) c[0] = c00;
}

// zig-zag scan / inverse
HLB_HD void zigzag4x4(const int m[16], int lv[16])
{
#pragma unroll
    for (int k = 0; k < 16; ++k) lv[k] = m[kZigzag[k]];
}
HLB_HD void inv_zigzag4x4(const int lv[16], int m[16])
{
#pragma unroll
    for (int k = 0; k < 16; ++k) m[kZigzag[k]] = lv[k];
}

// ------------------------------------------------------------------------------------------------------------------
// Distortion
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int sad16(const uint8_t a[16], const uint8_t b[16])
{
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += iabs((int)a[i] - (int)b[i]);
    return s;
}
HLB_HD int satd16(const uint8_t a[16], const uint8_t b[16])
{
    int d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = (int)a[i] - (int)b[i];
    hadamard4x4(d);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += iabs(d[i]);
    return s >> 1;
}

// ------------------------------------------------------------------------------------------------------------------
// CAVLC bit length of one residual block (residual.c:757-898).  The coeff_token length depends on nC, which depends
// on encoder history, so it is returned for all four nC classes and resolved by the caller.
// ------------------------------------------------------------------------------------------------------------------
struct CavlcInfo {
    uint8_t total_coeff;   // TotalCoeff(coeff_token)
    uint8_t trailing_ones; // TrailingOnes(coeff_token)
    uint8_t single_ctr;    // JVT-O079 2.3 contribution of this block (9 unless a lone +-1)
    uint16_t bits_rest;    // every bit except coeff_token
};

HLB_HD int coeff_token_len(int nC, int total_coeff, int trailing_ones)
{
    if (nC >= 8) return 6;
    const int vlc = nC < 2 ? 0 : (nC < 4 ? 1 : 2);
    return kCoeffTokenLen[vlc][total_coeff][trailing_ones];
}

// level_prefix/suffix length for levelCode under suffixLength (cavlc.c:59-104, level_prefix <= 15)
HLB_HD int level_code_len(int suffix_length, int level_code)
{
    if (suffix_length == 0) {
        if (level_code < 14) return level_code + 1;
        if (level_code < 30) return 15 + 4;
        return 16 + 12;
    }
    const int prefix = level_code >> suffix_length;
    if (prefix < 15) return prefix + 1 + suffix_length;
    return 16 + 12;
}

// lv: coefficients in scan order, n = maxNumCoeff handed to the reference (16 for every RDO call; 4 for chroma DC)
HLB_HD CavlcInfo cavlc_block_info_ref(const int* lv, int n, bool chroma_dc)
{
    int nz[16];
    int run[16];
    int tc = 0, t1 = 0, tz = 0, k = -1;
    bool count_t1 = true, seen = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) run[i] = 0;
    for (int j = 0; j < n; ++j) {  // reverse scan
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
    CavlcInfo r;
    r.total_coeff = (uint8_t)tc;
    r.trailing_ones = (uint8_t)t1;
    r.single_ctr = 9;
    int bits = 0;
    if (tc > 0) {
        int sl = (tc > 10 && t1 < 3) ? 1 : 0;
        for (int j = 0; j < tc; ++j) {
            if (j < t1) { bits += 1; continue; }
            const int v = nz[j];
            int lc = v > 0 ? (v << 1) - 2 : -(v << 1) - 1;
            if (j == t1 && t1 < 3 && lc >= 2) lc -= 2;
            bits += level_code_len(sl, lc);
            if (sl == 0) sl = 1;
            if (iabs(v) > (3 << (sl - 1)) && sl < 6) ++sl;
        }
        int zeros_left = 0;
        if (tc < n) {
            bits += chroma_dc ? kTotalZerosLenChromaDC[tc - 1][tz] : kTotalZerosLen[tc - 1][tz];
            zeros_left = tz;
        }
        for (int q = 0; q < tc - 1 && zeros_left > 0; ++q) {
            bits += kRunBeforeLen[(zeros_left > 7 ? 7 : zeros_left) - 1][run[q]];
            zeros_left -= run[q];
        }
        if (tc == 1 && (nz[0] == 1 || nz[0] == -1)) {
            const int rn = zeros_left > 0 ? run[0] : 0;
            r.single_ctr = (uint8_t)(rn < 6 ? (rn == 0 ? 3 : (rn < 3 ? 2 : 1)) : 0);
        }
    }
    r.bits_rest = (uint16_t)bits;
    return r;
}


// Register-only formulation for the only shape the kernels use (16 coefficients, luma/chroma-AC tables).  The loop visits only the
// NON-ZERO coefficients, from the high-frequency end, located with count-leading-zeros on the significance mask; a coefficient is
// fetched from the eight packed registers with a 3-level select tree (no local-memory array, no dynamic register indexing).  Run
// lengths come from the distance between consecutive set bits.  Typical residual blocks of the search hold 1-4 coefficients, so
// this executes ~4x fewer instructions than a scan over all 16 positions (r01c profile: the scan form was 1/3 of the slice kernel).
HLB_HD int hlb_clz(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
HLB_HD int hlb_popc(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// significance mask of 16 levels in scan order (bit i = lv[i] != 0)
HLB_HD uint32_t level_mask16(const int* lv)
{
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) mask |= (uint32_t)(lv[i] != 0) << i;
    return mask;
}
// core on the eight packed level pairs (r_k = lv[2k] | lv[2k+1] << 16) and the significance mask: nine register arguments, so the slice
// kernel can keep ONE out-of-line copy of it for all its callers (HLB_CAVLC_FN; its hot loop is instruction-fetch bound)
HLB_CAVLC_FN CavlcInfo cavlc_block_info_packed(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t r4, uint32_t r5, uint32_t r6, uint32_t r7, uint32_t mask)
{
    CavlcInfo out;
    out.total_coeff = 0; out.trailing_ones = 0; out.single_ctr = 9; out.bits_rest = 0;
    if (mask == 0) return out;
    const int tc = hlb_popc(mask), hb = 31 - hlb_clz(mask);
    const int tz = hb + 1 - tc;
    int bits = 0, zl = 0;
    if (tc < 16) { bits += kTotalZerosLen[tc - 1][tz]; zl = tz; }
    int t1 = 0, sl = 0, first_level = 1, first_v = 0, prev_p = hb;
    bool cnt = true;
    uint32_t m = mask;
#pragma unroll 1
    while (m) {
        const int p = 31 - hlb_clz(m);
        m &= ~(1u << p);
        const uint32_t q0 = (p & 2) ? r1 : r0, q1 = (p & 2) ? r3 : r2, q2 = (p & 2) ? r5 : r4, q3 = (p & 2) ? r7 : r6;
        const uint32_t h0 = (p & 4) ? q1 : q0, h1 = (p & 4) ? q3 : q2;
        const uint32_t wv = (p & 8) ? h1 : h0;
        const int c = (int)(int16_t)(uint16_t)(wv >> ((p & 1) << 4));
        if (p != hb) {   // run_before of the previous (higher-frequency) coefficient = zeros between it and this one
            const int run = prev_p - p - 1;
            if (zl > 0) { bits += kRunBeforeLen[(zl > 7 ? 7 : zl) - 1][run]; zl -= run; }
        } else first_v = c;
        prev_p = p;
        if (cnt && (c == 1 || c == -1)) { ++t1; bits += 1; cnt = t1 < 3; }
        else {
            cnt = false;
            int lc = c > 0 ? (c << 1) - 2 : -(c << 1) - 1;
            if (first_level) { sl = (tc > 10 && t1 < 3) ? 1 : 0; if (t1 < 3 && lc >= 2) lc -= 2; first_level = 0; }
            bits += level_code_len(sl, lc);
            if (sl == 0) sl = 1;
            if (iabs(c) > (3 << (sl - 1)) && sl < 6) ++sl;
        }
    }
    out.total_coeff = (uint8_t)tc; out.trailing_ones = (uint8_t)t1; out.bits_rest = (uint16_t)bits;
    if (tc == 1 && (first_v == 1 || first_v == -1)) out.single_ctr = (uint8_t)(hb < 6 ? (hb == 0 ? 3 : (hb < 3 ? 2 : 1)) : 0);
    return out;
}
HLB_HD CavlcInfo cavlc_block_info16(const int* lv, uint32_t mask)
{
    const uint32_t r0 = ((uint32_t)lv[0] & 0xffffu) | ((uint32_t)lv[1] << 16), r1 = ((uint32_t)lv[2] & 0xffffu) | ((uint32_t)lv[3] << 16);
    const uint32_t r2 = ((uint32_t)lv[4] & 0xffffu) | ((uint32_t)lv[5] << 16), r3 = ((uint32_t)lv[6] & 0xffffu) | ((uint32_t)lv[7] << 16);
    const uint32_t r4 = ((uint32_t)lv[8] & 0xffffu) | ((uint32_t)lv[9] << 16), r5 = ((uint32_t)lv[10] & 0xffffu) | ((uint32_t)lv[11] << 16);
    const uint32_t r6 = ((uint32_t)lv[12] & 0xffffu) | ((uint32_t)lv[13] << 16), r7 = ((uint32_t)lv[14] & 0xffffu) | ((uint32_t)lv[15] << 16);
    return cavlc_block_info_packed(r0, r1, r2, r3, r4, r5, r6, r7, mask);
}
HLB_HD CavlcInfo cavlc_block_info(const int* lv, int n, bool chroma_dc)
{
    if (n != 16 || chroma_dc) return cavlc_block_info_ref(lv, n, chroma_dc);
    return cavlc_block_info16(lv, level_mask16(lv));
}

}  // namespace hlb
