// hlb_fast.cuh -- packed-byte formulations of the per-4x4-block primitives of the search (prediction, residual transform, quantisation,
// trial reconstruction, SAD), written for the B200 integer pipes: rows of four samples stay packed in one register from the shared-memory
// reference tile to the distortion, so that
//   * windows are fetched as aligned 32-bit words + funnel shifts (SHF.R.W) instead of byte loads,
//   * the 6-tap filter and the forward transform's row pass are byte dot products (IDP.4A.U8.S8: four multiply-adds per instruction),
//   * vertical filters work on two samples per register (16-bit halves, VIADDMNMX.S16x2.RELU for the round-and-clip),
//   * rounding + clipping + packing is one I2IP.U8.S32.SAT per two samples, averages are byte-parallel,
//   * SAD is VABSDIFF4.U8.ACC (one instruction per row).
// Arithmetic is the reference's, bit for bit (interpol.h:41-923, transf.c:716-768, quant.c:116-137, quant.c:68-111, transf.c:420-456,
// hl_math.h:261,303-323, hl_math.c:239); tools/emu/check_fast.cpp checks every function against the plain formulations of hlb_prims.cuh
// (which the oracle pins against the reference) on the CPU, hlb200_dev_selftest does the same on the device.
#pragma once
#include "hlb_prims.cuh"

namespace hlb {

// ------------------------------------------------------------------------------------------------------------------
// portable wrappers of the packed instructions (plain C++ when not compiling device code)
// ------------------------------------------------------------------------------------------------------------------
HLB_HD uint32_t p_shf_r(uint32_t lo, uint32_t hi, uint32_t sh)   // low word of (hi:lo) >> (sh & 31)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
HLB_HD uint32_t p_prmt(uint32_t a, uint32_t b, uint32_t sel)   // byte k of the result = byte (sel >> 4k) & 7 of (b:a)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xff) << (8 * k);
    return r;
#endif
}
HLB_HD int p_dp4a_us(uint32_t a_u8, uint32_t b_s8, int c)   // c + sum_k a.u8[k] * b.s8[k]
{
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8), "r"(b_s8), "r"(c));
    return d;
#else
    for (int k = 0; k < 4; ++k) c += (int)((a_u8 >> (8 * k)) & 0xff) * (int)(int8_t)((b_s8 >> (8 * k)) & 0xff);
    return c;
#endif
}
HLB_HD uint32_t p_sad4(uint32_t a, uint32_t b, uint32_t c)   // c + sum_k |a.u8[k] - b.u8[k]|
{
#if defined(__CUDA_ARCH__)
    return __vsadu4(a, b) + c;
#else
    for (int k = 0; k < 4; ++k) { const int d = (int)((a >> (8 * k)) & 0xff) - (int)((b >> (8 * k)) & 0xff); c += (uint32_t)(d < 0 ? -d : d); }
    return c;
#endif
}
HLB_HD uint32_t p_pack_sat_u8(int hi, int lo, uint32_t c)   // (sat_u8(hi) << 8 | sat_u8(lo)) | c << 16
{
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    return d;
#else
    return ((uint32_t)clip255(hi) << 8) | (uint32_t)clip255(lo) | (c << 16);
#endif
}
HLB_HD uint32_t p_addmin_relu_s16x2(uint32_t a, uint32_t b, uint32_t c)   // per signed 16-bit half: max(min(a + b, c), 0)
{
#if defined(__CUDA_ARCH__)
    return __viaddmin_s16x2_relu(a, b, c);
#else
    uint32_t r = 0;
    for (int k = 0; k < 2; ++k) {
        int v = (int)(int16_t)(uint16_t)((a >> (16 * k)) + (b >> (16 * k)));   // wraps like the instruction
        const int m = (int16_t)(uint16_t)(c >> (16 * k));
        v = v < m ? v : m;
        v = v < 0 ? 0 : v;
        r |= (uint32_t)(uint16_t)v << (16 * k);
    }
    return r;
#endif
}
HLB_HD uint32_t pack16(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }   // two 16-bit levels per word
HLB_HD uint32_t p_avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu); }                       // per byte (a + b + 1) >> 1
HLB_HD uint32_t p_add4_wrap(uint32_t a, uint32_t b) { return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u); }   // per byte (a + b) mod 256
HLB_HD uint32_t pack4_sat(int v0, int v1, int v2, int v3) { return p_pack_sat_u8(v1, v0, p_pack_sat_u8(v3, v2, 0)); }         // clip255 of each, v0 in the low byte

// ------------------------------------------------------------------------------------------------------------------
// window rows out of a byte tile addressed as 32-bit words (`t` 4-byte aligned, row pitch `pw` words)
// ------------------------------------------------------------------------------------------------------------------
// how a word of the tile is read: plain (shared memory tile, local staging, host) or through the read-only path (a picture plane in global memory)
struct LdPlain { static HLB_HD uint32_t ld(const uint32_t* p) { return *p; } };
struct LdReadOnly { static HLB_HD uint32_t ld(const uint32_t* p) { return HLB_LDG(p); } };
// four samples starting at byte column c of row y
template <class LD = LdPlain> HLB_HD uint32_t tile_row4(const uint32_t* t, int pw, int y, int c)
{
    const uint32_t* p = t + y * pw + (c >> 2);
    return p_shf_r(LD::ld(p), LD::ld(p + 1), (uint32_t)(c & 3) * 8);
}
// nine samples starting at byte column c of row y: R0 = samples 0..3, R1 = 4..7, R2 = sample 8 in its low byte (higher bytes unspecified)
template <class LD = LdPlain> HLB_HD void tile_row9(const uint32_t* t, int pw, int y, int c, uint32_t& R0, uint32_t& R1, uint32_t& R2)
{
    const uint32_t* p = t + y * pw + (c >> 2);
    const uint32_t sh = (uint32_t)(c & 3) * 8, w0 = LD::ld(p), w1 = LD::ld(p + 1), w2 = LD::ld(p + 2);
    R0 = p_shf_r(w0, w1, sh); R1 = p_shf_r(w1, w2, sh); R2 = w2 >> sh;
}

// unrounded horizontal 6-tap values (+ acc) at the four positions whose window starts at samples 0..3 of (R0,R1,R2): tap6(b[x..x+5])
#define HLB_TAP_A 0x1414FB01u   /* ( 1, -5, 20, 20) on window samples x   .. x+3 */
#define HLB_TAP_B 0x000001FBu   /* (-5,  1,  0,  0) on window samples x+4 .. x+7 */
HLB_HD void hrow_taps(uint32_t R0, uint32_t R1, uint32_t R2, int acc, int& v0, int& v1, int& v2, int& v3)
{
    v0 = p_dp4a_us(R1, HLB_TAP_B, p_dp4a_us(R0, HLB_TAP_A, acc));
    v1 = p_dp4a_us(p_shf_r(R1, R2, 8), HLB_TAP_B, p_dp4a_us(p_shf_r(R0, R1, 8), HLB_TAP_A, acc));
    v2 = p_dp4a_us(p_shf_r(R1, R2, 16), HLB_TAP_B, p_dp4a_us(p_shf_r(R0, R1, 16), HLB_TAP_A, acc));
    v3 = p_dp4a_us(p_shf_r(R1, R2, 24), HLB_TAP_B, p_dp4a_us(p_shf_r(R0, R1, 24), HLB_TAP_A, acc));
}
// rounded horizontal half samples b of a row: clip255((tap6 + 16) >> 5), packed
HLB_HD uint32_t half_h_row(uint32_t R0, uint32_t R1, uint32_t R2)
{
    int v0, v1, v2, v3;
    hrow_taps(R0, R1, R2, 16, v0, v1, v2, v3);
    return pack4_sat(v0 >> 5, v1 >> 5, v2 >> 5, v3 >> 5);
}
// rounded vertical half samples h of one row of four positions from the six packed rows e..j above / below it.  Two samples per register:
// t = (E + J + K) + 20 (G + H) - 5 (F + I) with K = 2560 = 80 * 32 keeps both 16-bit halves non-negative (so plain 32-bit arithmetic never
// borrows across them) and commutes with the >> 5; ((t + 16) >> 5) - 80 is then clamped to 0..255 by one VIADDMNMX.S16x2.RELU.
HLB_HD uint32_t vtap_pair(uint32_t E, uint32_t F, uint32_t G, uint32_t H, uint32_t I, uint32_t J)
{
    const uint32_t t = (E + J + 0x0A000A00u) + 20u * (G + H) - 5u * (F + I);
    const uint32_t r = ((t + 0x00100010u) >> 5) & 0x07FF07FFu;
    return p_addmin_relu_s16x2(r, 0xFFB0FFB0u, 0x00FF00FFu);
}
HLB_HD uint32_t exp_lo(uint32_t r) { return p_prmt(r, 0, 0x4140); }   // samples 0,1 as 16-bit halves
HLB_HD uint32_t exp_hi(uint32_t r) { return p_prmt(r, 0, 0x4342); }   // samples 2,3

// ------------------------------------------------------------------------------------------------------------------
// Luma prediction of one 4x4 block out of the reference tile (8.4.2.2.1): (tx,ty) = tile coordinates of integer sample G of the block's
// pixel (0,0); the tile holds columns tx-2..tx+6 and rows ty-2..ty+6.  Returns the four rows packed.
// ------------------------------------------------------------------------------------------------------------------
#ifndef HLB_FASTPRED_FN   /* hlb_slice.cu overrides: one out-of-line copy for all callers, the tile known to live in shared memory */
#if defined(__CUDACC__)
#define HLB_FASTPRED_FN __device__ __forceinline__
#else
#define HLB_FASTPRED_FN inline
#endif
#define HLB_FASTPRED_SRC(t) ((void)0)
#endif
template <class LD> HLB_HD Rows4 fast_pred_luma_t(const uint32_t* t, int pw, int tx, int ty, int xf, int yf)
{
    Rows4 o;
    if ((xf | yf) == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) o.r[r] = tile_row4<LD>(t, pw, ty + r, tx);
    } else if (yf == 0) {   // a b c
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t R0, R1, R2;
            tile_row9<LD>(t, pw, ty + r, tx - 2, R0, R1, R2);
            const uint32_t b = half_h_row(R0, R1, R2);
            o.r[r] = xf == 2 ? b : p_avg4(p_shf_r(R0, R1, xf == 1 ? 16 : 24), b);
        }
    } else if (xf == 0) {   // d h n
        uint32_t lo[9], hi[9], g[5];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint32_t v = tile_row4<LD>(t, pw, ty - 2 + r, tx);
            lo[r] = exp_lo(v); hi[r] = exp_hi(v);
            if (r >= 2 && r < 7) g[r - 2] = v;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t h = p_prmt(vtap_pair(lo[r], lo[r + 1], lo[r + 2], lo[r + 3], lo[r + 4], lo[r + 5]), vtap_pair(hi[r], hi[r + 1], hi[r + 2], hi[r + 3], hi[r + 4], hi[r + 5]), 0x6420);
            o.r[r] = yf == 2 ? h : p_avg4(g[yf == 1 ? r : r + 1], h);
        }
    } else if ((xf & yf) & 1) {   // e g p r: average of b (row y or y+1) and h (column x or x+1)
        const int cx = tx + (xf == 3 ? 1 : 0), ry = yf == 3 ? 1 : 0;
        uint32_t lo[9], hi[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint32_t v = tile_row4<LD>(t, pw, ty - 2 + r, cx);
            lo[r] = exp_lo(v); hi[r] = exp_hi(v);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t R0, R1, R2;
            tile_row9<LD>(t, pw, ty + r + ry, tx - 2, R0, R1, R2);
            const uint32_t b = half_h_row(R0, R1, R2);
            const uint32_t h = p_prmt(vtap_pair(lo[r], lo[r + 1], lo[r + 2], lo[r + 3], lo[r + 4], lo[r + 5]), vtap_pair(hi[r], hi[r + 1], hi[r + 2], hi[r + 3], hi[r + 4], hi[r + 5]), 0x6420);
            o.r[r] = p_avg4(b, h);
        }
    } else {
        // j, and f q (xf == 2: averaged with b of row y / y+1) or i k (yf == 2: averaged with h of column x / x+1): the nine rows of unrounded
        // horizontal half samples stream through a six-row window
        int w0[4], w1[4], w2[4], w3[4], w4[4], w5[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) w0[x] = w1[x] = w2[x] = w3[x] = w4[x] = w5[x] = 0;
        uint32_t j0 = 0, j1 = 0, j2 = 0, j3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        const int rb = yf == 3 ? 3 : 2;   // window row holding the horizontal half samples that f / q average with (row y / y+1)
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            uint32_t R0, R1, R2;
            tile_row9<LD>(t, pw, ty - 2 + r, tx - 2, R0, R1, R2);
#pragma unroll
            for (int x = 0; x < 4; ++x) { w0[x] = w1[x]; w1[x] = w2[x]; w2[x] = w3[x]; w3[x] = w4[x]; w4[x] = w5[x]; }
            hrow_taps(R0, R1, R2, 0, w5[0], w5[1], w5[2], w5[3]);
            if (r >= 5) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = (w0[x] + w5[x] + 512 - 5 * (w1[x] + w4[x]) + 20 * (w2[x] + w3[x])) >> 10;
                j0 = j1; j1 = j2; j2 = j3; j3 = pack4_sat(v[0], v[1], v[2], v[3]);
                if (xf == 2 && yf != 2) {
                    b0 = b1; b1 = b2; b2 = b3;
                    b3 = rb == 2 ? pack4_sat((w2[0] + 16) >> 5, (w2[1] + 16) >> 5, (w2[2] + 16) >> 5, (w2[3] + 16) >> 5)
                                 : pack4_sat((w3[0] + 16) >> 5, (w3[1] + 16) >> 5, (w3[2] + 16) >> 5, (w3[3] + 16) >> 5);
                }
            }
        }
        o.r[0] = j0; o.r[1] = j1; o.r[2] = j2; o.r[3] = j3;
        if (xf == 2) {
            if (yf != 2) { o.r[0] = p_avg4(b0, j0); o.r[1] = p_avg4(b1, j1); o.r[2] = p_avg4(b2, j2); o.r[3] = p_avg4(b3, j3); }
        } else {
            const int cx = tx + (xf == 3 ? 1 : 0);
            uint32_t lo[9], hi[9];
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                const uint32_t v = tile_row4<LD>(t, pw, ty - 2 + r, cx);
                lo[r] = exp_lo(v); hi[r] = exp_hi(v);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t h = p_prmt(vtap_pair(lo[r], lo[r + 1], lo[r + 2], lo[r + 3], lo[r + 4], lo[r + 5]), vtap_pair(hi[r], hi[r + 1], hi[r + 2], hi[r + 3], hi[r + 4], hi[r + 5]), 0x6420);
                o.r[r] = p_avg4(h, o.r[r]);
            }
        }
    }
    return o;
}
// The same prediction written as predicated STAGES instead of one branch per fractional class, for callers whose lanes hold blocks of different classes
// (the whole-picture kernel k_interp_luma: every 4x4 block of a warp may carry its own vector).  Every class is avg(X, Y) of two of {G integer samples,
// b horizontal half samples of row y / y+1, h vertical half samples of column x / x+1, j centre half samples} (8.4.2.2.1, Table 8-12; X = Y for the
// positions that are one of them), so a warp runs at most: the horizontal tap rows (4, or all 9 when a lane needs j), the vertical pass over them, the
// vertical half samples, the integer rows and one select + average -- about 550 instructions however mixed its lanes are, against the sum of all
// sixteen branches of fast_pred_luma_t (measured 1,930 per warp on a random motion field).  Bit-exact with fast_pred_luma_t (tools/emu/check_fast.cpp).
template <class LD> HLB_HD Rows4 fast_pred_luma_staged(const uint32_t* t, int pw, int tx, int ty, int xf, int yf)
{
    const bool sB = xf != 0 && yf != 2, sH = yf != 0 && xf != 2, sJ = (xf == 2 && yf != 0) || (yf == 2 && xf != 0);
    const bool isG = (xf == 0 || yf == 0) && (((xf | yf) & 1) != 0 || (xf | yf) == 0);
    Rows4 B, H, J, G;
#pragma unroll
    for (int r = 0; r < 4; ++r) B.r[r] = H.r[r] = J.r[r] = G.r[r] = 0;
    if (sB || sJ) {
        const bool lower = yf == 3;   // b of row y+1 (positions p q r) instead of row y
        int w[9][4];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            w[r][0] = w[r][1] = w[r][2] = w[r][3] = 0;
            if (sJ || (r >= 2 && r <= 6 && (r != 2 || !lower) && (r != 6 || lower))) {
                uint32_t R0, R1, R2;
                tile_row9<LD>(t, pw, ty - 2 + r, tx - 2, R0, R1, R2);
                hrow_taps(R0, R1, R2, 0, w[r][0], w[r][1], w[r][2], w[r][3]);
            }
        }
        if (sB) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = ((lower ? w[3 + k][x] : w[2 + k][x]) + 16) >> 5;
                B.r[k] = pack4_sat(v[0], v[1], v[2], v[3]);
            }
        }
        if (sJ) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = (w[k][x] + w[k + 5][x] + 512 - 5 * (w[k + 1][x] + w[k + 4][x]) + 20 * (w[k + 2][x] + w[k + 3][x])) >> 10;
                J.r[k] = pack4_sat(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if (sH) {
        const int cx = tx + (xf == 3 ? 1 : 0);
        uint32_t lo[9], hi[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint32_t v = tile_row4<LD>(t, pw, ty - 2 + r, cx);
            lo[r] = exp_lo(v); hi[r] = exp_hi(v);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            H.r[r] = p_prmt(vtap_pair(lo[r], lo[r + 1], lo[r + 2], lo[r + 3], lo[r + 4], lo[r + 5]), vtap_pair(hi[r], hi[r + 1], hi[r + 2], hi[r + 3], hi[r + 4], hi[r + 5]), 0x6420);
    }
    if (isG) {
        const int gx = tx + (xf == 3 ? 1 : 0), gy = ty + (yf == 3 ? 1 : 0);
#pragma unroll
        for (int r = 0; r < 4; ++r) G.r[r] = tile_row4<LD>(t, pw, gy + r, gx);
    }
    Rows4 o;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t X = sB ? B.r[r] : (sH ? H.r[r] : (sJ ? J.r[r] : G.r[r]));
        const uint32_t Y = sJ ? J.r[r] : ((sB && sH) ? H.r[r] : (isG ? G.r[r] : X));
        o.r[r] = p_avg4(X, Y);
    }
    return o;
}
// the shared-memory tile of the slice kernel (one out-of-line copy there, see hlb_slice.cu)
HLB_FASTPRED_FN Rows4 fast_pred_luma(const uint32_t* t, int pw, int tx, int ty, int xf, int yf)
{
    HLB_FASTPRED_SRC(t);
#ifdef HLB_PRED_STAGED   /* measured in the slice kernel: 4.99 M against 5.14 M macroblocks/s -- most search steps are integer-pel, where the per-class form is four word loads */
    return fast_pred_luma_staged<LdPlain>(t, pw, tx, ty, xf, yf);
#else
    return fast_pred_luma_t<LdPlain>(t, pw, tx, ty, xf, yf);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// Chroma prediction (8.4.2.2.2, pred_inter.c:888-940, interpol.c:337-385): the two samples (x0, y0), (x0 + 1, y0) of a plane (pitch Wc, height Hc, a read-only
// picture plane in global memory) at eighth-sample fraction (xf, yf), returned in the two low bytes.  Interior: two rows fetched as aligned words + funnel shift;
// at the picture edge the reference's per-sample clamp.  The bilinear sample is two byte dot products (weights (8-xf)(8-yf), xf(8-yf) | (8-xf)yf, xf yf).
// ------------------------------------------------------------------------------------------------------------------
HLB_HD uint32_t fast_chroma_two(const uint8_t* rp, int Wc, int Hc, int x0, int y0, int xf, int yf)
{
    const uint32_t wa = (uint32_t)((8 - xf) * (8 - yf)) | ((uint32_t)(xf * (8 - yf)) << 8), wc = (uint32_t)((8 - xf) * yf) | ((uint32_t)(xf * yf) << 8);
    uint32_t a, c;   // three samples of row y0 / y0 + 1 starting at x0
    if (x0 >= 0 && y0 >= 0 && y0 + 1 < Hc && (x0 >> 2) + 1 < (Wc >> 2)) {
        const uint32_t* ra = reinterpret_cast<const uint32_t*>(rp + (size_t)y0 * Wc) + (x0 >> 2);
        const uint32_t* rc = ra + (Wc >> 2);
        const uint32_t sh = (uint32_t)(x0 & 3) * 8;
        a = p_shf_r(HLB_LDG(ra), HLB_LDG(ra + 1), sh); c = p_shf_r(HLB_LDG(rc), HLB_LDG(rc + 1), sh);
    } else {
        const uint8_t* ra = rp + (size_t)clip3(0, Hc - 1, y0) * Wc;
        const uint8_t* rc = rp + (size_t)clip3(0, Hc - 1, y0 + 1) * Wc;
        a = c = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int x = clip3(0, Wc - 1, x0 + i);
            a |= (uint32_t)HLB_LDG(ra + x) << (8 * i); c |= (uint32_t)HLB_LDG(rc + x) << (8 * i);
        }
    }
    const int v0 = p_dp4a_us(c, wc, p_dp4a_us(a, wa, 32)) >> 6, v1 = p_dp4a_us(c >> 8, wc, p_dp4a_us(a >> 8, wa, 32)) >> 6;
    return (uint32_t)v0 | ((uint32_t)v1 << 8);
}

// ------------------------------------------------------------------------------------------------------------------
// Quantiser constants of a picture (computed once per picture, host or device): the inter rounding offset of the search's trial encodes
// ------------------------------------------------------------------------------------------------------------------
struct QuantK {
    int32_t qbits;        // 15 + QP / 6
    int32_t f_pos;        // (1 << qbits) / 6: added to w * MF for w >= 0                      (quant.c:116-137)
    int32_t f_neg;        // (1 << qbits) - 1 - f_pos: added to w * MF for w < 0, then an arithmetic shift gives -((|w| MF + f) >> qbits)
    int32_t mf[3];        // by position class (even,even) / (odd,odd) / mixed
    int32_t dq_mul[3];    // LevelScale4x4 by class, pre-shifted by QP/6 - 4 when QP >= 24           (quant.c:68-111)
    int32_t dq_shift;     // QP < 24: (c * LevelScale + (1 << (3 - QP/6))) >> (4 - QP/6); else 0
    int32_t dq_round;
    int32_t zero_sad;     // a residual block whose SAD is <= this quantises to all-zero levels (see quantk_make); -1 = no shortcut
};
inline void quantk_make(QuantK& k, int qp, bool intra = false)
{
    static const int MF[6][3] = {{13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554}, {9362, 3647, 5825}, {8192, 3355, 5243}, {7282, 2893, 4559}};
    static const int NA[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
    const int r = qp % 6, q6 = qp / 6;
    k.qbits = 15 + q6;
    k.f_pos = (1 << k.qbits) / (intra ? 3 : 6);
    k.f_neg = (1 << k.qbits) - 1 - k.f_pos;
    for (int c = 0; c < 3; ++c) {
        k.mf[c] = MF[r][c];
        k.dq_mul[c] = qp >= 24 ? (16 * NA[r][c]) << (q6 - 4) : 16 * NA[r][c];
    }
    k.dq_shift = qp >= 24 ? 0 : 4 - q6;
    k.dq_round = qp >= 24 ? 0 : 1 << (3 - q6);
    // |W_ij| <= g_i g_j SAD with g = (1, 2, 1, 2) the largest magnitude in row i of the forward matrix; level_ij = 0 iff |W_ij| MF + f < 2^qbits.
    // zero_sad = the largest SAD for which that holds for every class.
    const int lim = (1 << k.qbits) - k.f_pos - 1;   // |W| * MF <= lim
    const int t0 = lim / MF[r][0], t1 = (lim / MF[r][1]) / 4, t2 = (lim / MF[r][2]) / 2;
    k.zero_sad = t0 < t1 ? (t0 < t2 ? t0 : t2) : (t1 < t2 ? t1 : t2);
}

// ------------------------------------------------------------------------------------------------------------------
// One trial encode of a 4x4 block (me_ds.c:527-688 for one block): residual -> forward transform -> quantisation -> [CAVLC counts ->
// de-quantisation -> inverse transform -> wrap-around reconstruction -> SAD].  Returns the memo word of me_phase_trial:
// dist:12 | bits_rest:10 | TotalCoeff:5 | TrailingOnes:2 | lone-coefficient Single_ctr:2.  `counts_only`: the caller needs TotalCoeff /
// TrailingOnes / Single_ctr but neither the distortion nor the bit count (searches that can no longer improve, see me_find_best_cost).
// ------------------------------------------------------------------------------------------------------------------
#define HLB_CF0 0x01010101u   /* ( 1,  1,  1,  1) */
#define HLB_CF1 0xFEFF0102u   /* ( 2,  1, -1, -2) */
#define HLB_CF2 0x01FFFF01u   /* ( 1, -1, -1,  1) */
#define HLB_CF3 0xFF02FE01u   /* ( 1, -2,  2, -1) */
#define HLB_NCF0 0xFFFFFFFFu
#define HLB_NCF1 0x0201FFFEu
#define HLB_NCF2 0xFF0101FFu
#define HLB_NCF3 0x01FE02FFu
// forward 4x4 transform of the residual (source - prediction) of packed rows, raster order: rows (S - P) Cf^T as byte dot products, then columns Cf (.)
HLB_HD void fast_fwd_transform(const Rows4& s, const Rows4& p, int m[16])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        m[r * 4 + 0] = p_dp4a_us(s.r[r], HLB_CF0, p_dp4a_us(p.r[r], HLB_NCF0, 0));
        m[r * 4 + 1] = p_dp4a_us(s.r[r], HLB_CF1, p_dp4a_us(p.r[r], HLB_NCF1, 0));
        m[r * 4 + 2] = p_dp4a_us(s.r[r], HLB_CF2, p_dp4a_us(p.r[r], HLB_NCF2, 0));
        m[r * 4 + 3] = p_dp4a_us(s.r[r], HLB_CF3, p_dp4a_us(p.r[r], HLB_NCF3, 0));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = m[j], b = m[4 + j], c = m[8 + j], d = m[12 + j];
        const int s0 = a + d, s1 = b + c, d0 = a - d, d1 = b - c;
        m[j] = s0 + s1; m[4 + j] = 2 * d0 + d1; m[8 + j] = s0 - s1; m[12 + j] = d0 - 2 * d1;
    }
}
// quantisation in place: sign(w) * ((|w| MF + f) >> qbits) == (w MF + (w < 0 ? f_neg : f_pos)) >> qbits (arithmetic shift)
HLB_HD void fast_quant(int m[16], const QuantK& q)
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = m[i * 4 + j] * q.mf[pos_class(i, j)];
            m[i * 4 + j] = (t + (t < 0 ? q.f_neg : q.f_pos)) >> q.qbits;
        }
}
// de-quantisation (flat scaling lists; keep_dc: coefficient (0,0) is already de-quantised) + inverse transform, in place; `round` = 32 for the final (x + 32) >> 6
HLB_HD void fast_dequant_inverse(int m[16], const QuantK& q, bool keep_dc)
{
    const int dc = m[0];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = m[i * 4 + j] * q.dq_mul[pos_class(i, j)];
            m[i * 4 + j] = (v + q.dq_round) >> q.dq_shift;
        }
    if (keep_dc) m[0] = dc;
    m[0] += 32;   // reaches every output sample once with weight 1: the rounding of the final >> 6
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d0 = m[4 * i], d1 = m[4 * i + 1], d2 = m[4 * i + 2], d3 = m[4 * i + 3];
        const int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        m[4 * i] = e0 + e3; m[4 * i + 1] = e1 + e2; m[4 * i + 2] = e1 - e2; m[4 * i + 3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int f0 = m[j], f1 = m[4 + j], f2 = m[8 + j], f3 = m[12 + j];
        const int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        m[j] = (g0 + g3) >> 6; m[4 + j] = (g1 + g2) >> 6; m[8 + j] = (g1 - g2) >> 6; m[12 + j] = (g0 - g3) >> 6;
    }
}
// final reconstruction of packed prediction rows + residual samples: clip255(p + r) (hl_math.h:278), packed
HLB_HD Rows4 fast_recon_clip(const Rows4& p, const int r[16])
{
    Rows4 o;
#pragma unroll
    for (int y = 0; y < 4; ++y)
        o.r[y] = pack4_sat((int)(p.r[y] & 255u) + r[y * 4], (int)((p.r[y] >> 8) & 255u) + r[y * 4 + 1], (int)((p.r[y] >> 16) & 255u) + r[y * 4 + 2], (int)(p.r[y] >> 24) + r[y * 4 + 3]);
    return o;
}
HLB_HD uint32_t fast_trial(const Rows4& s, const Rows4& p, const QuantK& q, bool counts_only)
{
    uint32_t sad0 = p_sad4(s.r[0], p.r[0], 0);
    sad0 = p_sad4(s.r[1], p.r[1], sad0); sad0 = p_sad4(s.r[2], p.r[2], sad0); sad0 = p_sad4(s.r[3], p.r[3], sad0);
    if ((int)sad0 <= q.zero_sad) return sad0;   // every level is zero (and so is the block's residual when sad0 == 0)
    int m[16];
    fast_fwd_transform(s, p, m);
    fast_quant(m, q);
    int lv[16];
    zigzag4x4(m, lv);
    const uint32_t mask = level_mask16(lv);
    if (mask == 0) return sad0;
    if (counts_only) {
        const int tc = hlb_popc(mask);
        uint32_t val = (uint32_t)tc << 22;
        if (tc == 1) {
            int sum = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) sum += lv[i];
            if (sum == 1 || sum == -1) {
                const int hb = 31 - hlb_clz(mask);
                val |= (1u << 27) | ((uint32_t)(hb < 6 ? (hb == 0 ? 3 : (hb < 3 ? 2 : 1)) : 0) << 29);
            }
        }
        return val;
    }
    const CavlcInfo ci = cavlc_block_info16(lv, mask);
    // de-quantisation (flat scaling lists) with the +32 of the final rounding folded into the DC coefficient, inverse transform
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = m[i * 4 + j] * q.dq_mul[pos_class(i, j)];
            m[i * 4 + j] = (v + q.dq_round) >> q.dq_shift;
        }
    m[0] += 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d0 = m[4 * i], d1 = m[4 * i + 1], d2 = m[4 * i + 2], d3 = m[4 * i + 3];
        const int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        m[4 * i] = e0 + e3; m[4 * i + 1] = e1 + e2; m[4 * i + 2] = e1 - e2; m[4 * i + 3] = e0 - e3;
    }
    uint32_t dist = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int f0 = m[j], f1 = m[4 + j], f2 = m[8 + j], f3 = m[12 + j];
        const int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        m[j] = (g0 + g3) >> 6; m[4 + j] = (g1 + g2) >> 6; m[8 + j] = (g1 - g2) >> 6; m[12 + j] = (g0 - g3) >> 6;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        // low bytes of the four residual samples of the row, added to the prediction modulo 256 (the reference's trial reconstruction wraps, hl_math.h:261)
        const uint32_t res = p_prmt(p_prmt((uint32_t)m[r * 4], (uint32_t)m[r * 4 + 1], 0x0040), p_prmt((uint32_t)m[r * 4 + 2], (uint32_t)m[r * 4 + 3], 0x0040), 0x5410);
        dist = p_sad4(s.r[r], p_add4_wrap(p.r[r], res), dist);
    }
    return dist | ((uint32_t)ci.bits_rest << 12) | ((uint32_t)ci.total_coeff << 22) | ((uint32_t)ci.trailing_ones << 27) | ((uint32_t)(ci.single_ctr & 3) << 29);
}

}  // namespace hlb
