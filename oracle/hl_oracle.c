/*
 * hl_oracle.c -- CPU restatement (plain C, scalar loops) of the reference's arithmetic for the encoder pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product (hartallo_b200/) never links, imports or executes it.
 *
 * PINNED: every function below is checked bit-for-bit against the reference's own function (called through
 * oracle/_ref/libref_kernels.so, built from /root/reference by oracle/build_ref.sh) by tests/test_oracle_pinned.py,
 * and against golden vectors generated from the reference and committed under tests/golden/.
 *
 * Each function cites the reference file:line it restates.  Written independently of the CUDA code (per-sample
 * formulas straight from the structure of the reference / H.264 clauses, no shared source).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HLO_API __attribute__((visibility("default")))

static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static int clip1(int v) { return clip3(0, 255, v); }

/* --- luma sample fetch ------------------------------------------------------------------------------------------
 * source/h264/hl_codec_264_pred_inter.c:384-396 + source/h264/hl_codec_264_interpol.c:74-131:
 * the partition ORIGIN is clipped to [-17, W+17] x [-17, H+17] (not each sample), then every sample is fetched through
 * an index table that clamps its coordinates to the picture. */
typedef struct { const uint8_t* p; int W, H, X, Y; } luma_src_t;
static int L(const luma_src_t* s, int dx, int dy)
{
    int x = clip3(0, s->W - 1, s->X + dx), y = clip3(0, s->H - 1, s->Y + dy);
    return s->p[y * s->W + x];
}
/* include/hartallo/h264/hl_codec_264_macros.h:73 Tap6Filter */
static int tap6(int E, int F, int G, int H, int I, int J) { return E - 5 * (F + I) + 20 * (G + H) + J; }
static int b1(const luma_src_t* s, int x, int y) { return tap6(L(s, x - 2, y), L(s, x - 1, y), L(s, x, y), L(s, x + 1, y), L(s, x + 2, y), L(s, x + 3, y)); }
static int h1(const luma_src_t* s, int x, int y) { return tap6(L(s, x, y - 2), L(s, x, y - 1), L(s, x, y), L(s, x, y + 1), L(s, x, y + 2), L(s, x, y + 3)); }
static int half_b(const luma_src_t* s, int x, int y) { return clip1((b1(s, x, y) + 16) >> 5); }
static int half_h(const luma_src_t* s, int x, int y) { return clip1((h1(s, x, y) + 16) >> 5); }
static int half_j(const luma_src_t* s, int x, int y)
{
    int j1 = tap6(b1(s, x, y - 2), b1(s, x, y - 1), b1(s, x, y), b1(s, x, y + 1), b1(s, x, y + 2), b1(s, x, y + 3));
    return clip1((j1 + 512) >> 10);
}

/* hl_codec_264_interpol_luma, pred_inter.c:339-885 (dispatch :401-845; kernels include/hartallo/h264/hl_codec_264_interpol.h:162-923).
 * Table 8-12 of the standard.  out is a 16x16 u8 array (stride 16); only partW x partH is written. */
HLO_API void hlo_interp_luma(const uint8_t* ref, int W, int H, int xL, int yL, int partW, int partH, int mvx, int mvy, uint8_t* out)
{
    luma_src_t s;
    int xf = mvx & 3, yf = mvy & 3, x, y;
    s.p = ref; s.W = W; s.H = H;
    s.X = clip3(-17, W + 17, xL + (mvx >> 2));
    s.Y = clip3(-17, H + 17, yL + (mvy >> 2));
    for (y = 0; y < partH; ++y) for (x = 0; x < partW; ++x) {
        int G = L(&s, x, y), v;
        switch (xf + 4 * yf) {
        case 0: v = G; break;
        case 1: v = (G + half_b(&s, x, y) + 1) >> 1; break;                              /* a */
        case 2: v = half_b(&s, x, y); break;                                               /* b */
        case 3: v = (L(&s, x + 1, y) + half_b(&s, x, y) + 1) >> 1; break;                  /* c */
        case 4: v = (G + half_h(&s, x, y) + 1) >> 1; break;                              /* d */
        case 5: v = (half_b(&s, x, y) + half_h(&s, x, y) + 1) >> 1; break;                 /* e */
        case 6: v = (half_b(&s, x, y) + half_j(&s, x, y) + 1) >> 1; break;                 /* f */
        case 7: v = (half_b(&s, x, y) + half_h(&s, x + 1, y) + 1) >> 1; break;             /* g */
        case 8: v = half_h(&s, x, y); break;                                               /* h */
        case 9: v = (half_h(&s, x, y) + half_j(&s, x, y) + 1) >> 1; break;                 /* i */
        case 10: v = half_j(&s, x, y); break;                                              /* j */
        case 11: v = (half_j(&s, x, y) + half_h(&s, x + 1, y) + 1) >> 1; break;            /* k */
        case 12: v = (L(&s, x, y + 1) + half_h(&s, x, y) + 1) >> 1; break;                 /* n */
        case 13: v = (half_h(&s, x, y) + half_b(&s, x, y + 1) + 1) >> 1; break;            /* p */
        case 14: v = (half_j(&s, x, y) + half_b(&s, x, y + 1) + 1) >> 1; break;            /* q */
        default: v = (half_h(&s, x + 1, y) + half_b(&s, x, y + 1) + 1) >> 1; break;        /* r */
        }
        out[y * 16 + x] = (uint8_t)v;
    }
}

/* hl_codec_264_interpol_chroma_cpp, pred_inter.c:888-940 -> hl_codec_264_interpol_chroma_cat1_u8_cpp, interpol.c:337-385
 * (sample loader interpol.c:250-334: per-sample clamp).  One plane; out is 8x8 (stride 8).  xL,yL = LUMA origin of
 * the partition, mv = chroma mv (= luma mv for 4:2:0 frame MBs, utils.c:834-851). */
HLO_API void hlo_interp_chroma(const uint8_t* refc, int Wc, int Hc, int xL, int yL, int partWc, int partHc, int mvx, int mvy, uint8_t* out)
{
    int x0 = (xL >> 1) + (mvx >> 3), y0 = (yL >> 1) + (mvy >> 3), xf = mvx & 7, yf = mvy & 7, x, y;
    for (y = 0; y < partHc; ++y) for (x = 0; x < partWc; ++x) {
        int xa = clip3(0, Wc - 1, x0 + x), xb = clip3(0, Wc - 1, x0 + x + 1);
        int ya = clip3(0, Hc - 1, y0 + y), yc = clip3(0, Hc - 1, y0 + y + 1);
        int A = refc[ya * Wc + xa], B = refc[ya * Wc + xb], C = refc[yc * Wc + xa], D = refc[yc * Wc + xb];
        out[y * 8 + x] = (uint8_t)(((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6);
    }
}

/* hl_codec_264_transf_frw_residual4x4_cpp, transf.c:716-768: W = Cf . X . Cf^T */
HLO_API void hlo_fwd4x4(const int32_t* in, int32_t* out)
{
    static const int Cf[4][4] = {{1, 1, 1, 1}, {2, 1, -1, -2}, {1, -1, -1, 1}, {1, -2, 2, -1}};
    int t[4][4], i, j, k;
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { t[i][j] = 0; for (k = 0; k < 4; ++k) t[i][j] += Cf[i][k] * in[k * 4 + j]; }
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { int a = 0; for (k = 0; k < 4; ++k) a += t[i][k] * Cf[j][k]; out[i * 4 + j] = a; }
}

/* HL_CODEC_264_QUANT_MF, source/h264/hl_codec_264_tables.c:10-17 (spec values) */
static int quant_mf(int qp, int i, int j)
{
    static const int a[6][3] = {{13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554}, {9362, 3647, 5825}, {8192, 3355, 5243}, {7282, 2893, 4559}};
    int k = (i % 2 == 0 && j % 2 == 0) ? 0 : ((i % 2 == 1 && j % 2 == 1) ? 1 : 2);
    return a[qp % 6][k];
}
/* normAdjust4x4, include/hartallo/h264/hl_codec_264_macros.h:69; LevelScale4x4 = flat16 * normAdjust, pps.c:38-80 */
static int level_scale(int qp, int i, int j)
{
    static const int v[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
    int k = (i % 2 == 0 && j % 2 == 0) ? 0 : ((i % 2 == 1 && j % 2 == 1) ? 1 : 2);
    return 16 * v[qp % 6][k];
}

/* hl_codec_264_quant_frw4x4_scale_ac_cpp, quant.c:116-137; tables.c:19-45 (qbits, f) */
HLO_API void hlo_quant4x4(int qp, int intra, const int32_t* in, int32_t* out)
{
    int qbits = 15 + qp / 6, f = (1 << qbits) / (intra ? 3 : 6), i, j;
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) {
        int w = in[i * 4 + j], z = (abs(w) * quant_mf(qp, i, j) + f) >> qbits;
        out[i * 4 + j] = w >= 0 ? z : -z;
    }
}
/* hl_codec_264_quant_frw4x4_scale_dc_luma_cpp quant.c:141-165 (n=16), hl_codec_264_quant_frw2x2_scale_dc_chroma_cpp quant.c:168-189 (n=4) */
HLO_API void hlo_quant_dc(int qp, int intra, const int32_t* in, int32_t* out, int n)
{
    int qbits = 15 + qp / 6, f = (1 << qbits) / (intra ? 3 : 6), i;
    for (i = 0; i < n; ++i) { int w = in[i], z = (abs(w) * quant_mf(qp, 0, 0) + 2 * f) >> (qbits + 1); out[i] = w >= 0 ? z : -z; }
}

/* hl_codec_264_quant_scale_residual4x4_cpp, quant.c:68-111 */
HLO_API void hlo_dequant4x4(int qp, int keep_dc, const int32_t* c, int32_t* d)
{
    int i, j;
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) {
        int v = c[i * 4 + j] * level_scale(qp, i, j);
        d[i * 4 + j] = qp >= 24 ? v << (qp / 6 - 4) : (v + (1 << (3 - qp / 6))) >> (4 - qp / 6);
    }
    if (keep_dc) d[0] = c[0];
}
/* hl_codec_264_transf_inverse_residual4x4_cpp, transf.c:420-456 */
HLO_API void hlo_inv4x4(const int32_t* d, int32_t* r)
{
    int e[4][4], f[4][4], g[4][4], h[4][4], i, j;
    for (i = 0; i < 4; ++i) {
        e[i][0] = d[i * 4] + d[i * 4 + 2]; e[i][1] = d[i * 4] - d[i * 4 + 2];
        e[i][2] = (d[i * 4 + 1] >> 1) - d[i * 4 + 3]; e[i][3] = d[i * 4 + 1] + (d[i * 4 + 3] >> 1);
        f[i][0] = e[i][0] + e[i][3]; f[i][1] = e[i][1] + e[i][2]; f[i][2] = e[i][1] - e[i][2]; f[i][3] = e[i][0] - e[i][3];
    }
    for (j = 0; j < 4; ++j) {
        g[0][j] = f[0][j] + f[2][j]; g[1][j] = f[0][j] - f[2][j]; g[2][j] = (f[1][j] >> 1) - f[3][j]; g[3][j] = f[1][j] + (f[3][j] >> 1);
        h[0][j] = g[0][j] + g[3][j]; h[1][j] = g[1][j] + g[2][j]; h[2][j] = g[1][j] - g[2][j]; h[3][j] = g[0][j] - g[3][j];
    }
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) r[i * 4 + j] = (h[i][j] + 32) >> 6;
}
/* hl_codec_264_transf_scale_residual4x4, transf.c:376-417 = dequant + inverse */
HLO_API void hlo_dequant_inv4x4(int qp, int keep_dc, const int32_t* c, int32_t* r)
{
    int32_t d[16];
    hlo_dequant4x4(qp, keep_dc, c, d);
    hlo_inv4x4(d, r);
}

static void hadamard4(const int32_t* in, int32_t* out)
{
    static const int Hm[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
    int t[4][4], i, j, k;
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { t[i][j] = 0; for (k = 0; k < 4; ++k) t[i][j] += Hm[i][k] * in[k * 4 + j]; }
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { int a = 0; for (k = 0; k < 4; ++k) a += t[i][k] * Hm[j][k]; out[i * 4 + j] = a; }
}
/* hl_codec_264_transf_frw_hadamard4x4_dc_luma_cpp, transf.c:774-841: H.X.H then >>1 */
HLO_API void hlo_hadamard4x4_dc_luma(const int32_t* in, int32_t* out)
{
    int i;
    hadamard4(in, out);
    for (i = 0; i < 16; ++i) out[i] >>= 1;
}
/* hl_codec_264_transf_scale_luma_dc_coeff_intra16x16_cpp, transf.c:498-608 */
HLO_API void hlo_scale_luma_dc(int qp, const int32_t* c, int32_t* dcY)
{
    int32_t f[16];
    int i, scale = level_scale(qp, 0, 0);
    hadamard4(c, f);
    for (i = 0; i < 16; ++i) dcY[i] = qp >= 36 ? (f[i] * scale) << (qp / 6 - 6) : (f[i] * scale + (1 << (5 - qp / 6))) >> (6 - qp / 6);
}
/* hl_codec_264_transf_frw_hadamard2x2_dc_chroma, transf.c:843-868 */
HLO_API void hlo_hadamard2x2(const int32_t* in, int32_t* out)
{
    int a = in[0] + in[2], b = in[1] + in[3], c = in[0] - in[2], d = in[1] - in[3];
    out[0] = a + b; out[1] = a - b; out[2] = c + d; out[3] = c - d;
}
/* hl_codec_264_transf_scale_chroma_dc_coeff (ChromaArrayType 1), transf.c:612-700 */
HLO_API void hlo_scale_chroma_dc(int qpc, const int32_t* c, int32_t* dc)
{
    int32_t f[4];
    int i, scale = level_scale(qpc, 0, 0);
    hlo_hadamard2x2(c, f);
    for (i = 0; i < 4; ++i) dc[i] = ((f[i] * scale) << (qpc / 6)) >> 5;
}

/* hl_math_sad4x4_u8_cpp, source/hl_math.c:239-257 */
HLO_API int hlo_sad4x4(const uint8_t* a, int sa, const uint8_t* b, int sb)
{
    int s = 0, x, y;
    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) s += abs((int)a[y * sa + x] - (int)b[y * sb + x]);
    return s;
}
/* hl_math_ssd4x4_u8_cpp, source/hl_math.c:360-375 */
HLO_API int hlo_ssd4x4(const uint8_t* a, int sa, const uint8_t* b, int sb)
{
    int s = 0, x, y;
    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) { const int d = (int)a[y * sa + x] - (int)b[y * sb + x]; s += d * d; }
    return s;
}
/* hl_math_homogeneousity8x8_u8_cpp, source/hl_math.c:470-486: JVT-O079 edge map (2-35); p = top-left sample of the 8x8 block, not on the plane's border */
HLO_API int hlo_homogeneity8x8(const uint8_t* p, int stride)
{
    int s = 0, i, j;
    for (j = 0; j < 8; ++j) for (i = 0; i < 8; ++i) {
        const uint8_t *c = p + j * stride + i, *up = c - stride, *dn = c + stride;
        s += abs(dn[-1] + 2 * dn[0] + dn[1] - up[-1] - 2 * up[0] - up[1]) + abs(up[1] + 2 * c[1] + dn[1] - up[-1] - 2 * c[-1] - dn[-1]);
    }
    return s;
}
/* hl_math_satd4x4_u8_cpp, source/hl_math.c:283-357: sum |H.D.H| >> 1 */
HLO_API int hlo_satd4x4(const uint8_t* a, int sa, const uint8_t* b, int sb)
{
    int32_t d[16], t[16];
    int s = 0, x, y;
    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) d[y * 4 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    hadamard4(d, t);
    for (x = 0; x < 16; ++x) s += abs(t[x]);
    return s >> 1;
}
/* hl_math_addclip_4x4_u8xi32_cpp, include/hartallo/hl_math.h:261,303-323: the sum is stored to a uint8_t BEFORE the
 * clip, so it wraps modulo 256. */
HLO_API void hlo_addclip_u8xi32(const uint8_t* pred, const int32_t* res, uint8_t* out)
{
    int i;
    for (i = 0; i < 16; ++i) out[i] = (uint8_t)((int)pred[i] + res[i]);
}
/* hl_math_addclip_4x4_cpp, hl_math.h:278-300: Clip3(0,255, pred + res) */
HLO_API void hlo_addclip_i32(const int32_t* pred, const int32_t* res, int32_t* out)
{
    int i;
    for (i = 0; i < 16; ++i) out[i] = clip1(pred[i] + res[i]);
}

/* zig-zag: Scan4x4_L / InverseScan4x4, include/hartallo/h264/hl_codec_264_utils.h:146-180 */
static const int kZZ[16][2] = {{0, 0}, {0, 1}, {1, 0}, {2, 0}, {1, 1}, {0, 2}, {0, 3}, {1, 2}, {2, 1}, {3, 0}, {3, 1}, {2, 2}, {1, 3}, {2, 3}, {3, 2}, {3, 3}};
HLO_API void hlo_zigzag(const int32_t* m, int32_t* lv) { int k; for (k = 0; k < 16; ++k) lv[k] = m[kZZ[k][0] * 4 + kZZ[k][1]]; }
HLO_API void hlo_inv_zigzag(const int32_t* lv, int32_t* m) { int k; for (k = 0; k < 16; ++k) m[kZZ[k][0] * 4 + kZZ[k][1]] = lv[k]; }

/* ---------------------------------------------------------------------------------------------------------------
 * CAVLC bit length of a residual block in RDO mode: hl_codec_264_residual_write_block_cavlc, residual.c:757-898,
 * writers cavlc.c:652-836 (tables = H.264 Tables 9-5, 9-7..9-10), level table cavlc.c:59-104.
 * nC is given by the caller (residual.c:698-755 computes it from neighbour state).
 * Returns total bits; *single_ctr per residual.c:881-897; *total_coeff as stored at residual.c:797-805. */
static const uint8_t kCT[3][4][17] = {
    {{1, 6, 8, 9, 10, 11, 13, 13, 13, 14, 14, 15, 15, 16, 16, 16, 16}, {0, 2, 6, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 15, 16, 16, 16},
     {0, 0, 3, 7, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 16, 16, 16}, {0, 0, 0, 5, 6, 7, 8, 9, 10, 11, 13, 14, 14, 15, 15, 16, 16}},
    {{2, 6, 6, 7, 8, 8, 9, 11, 11, 12, 12, 12, 13, 13, 13, 14, 14}, {0, 2, 5, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 14, 14, 14},
     {0, 0, 3, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 13, 14, 14}, {0, 0, 0, 4, 4, 5, 6, 6, 7, 9, 11, 11, 12, 13, 13, 13, 14}},
    {{4, 6, 6, 6, 7, 7, 7, 7, 8, 8, 9, 9, 9, 10, 10, 10, 10}, {0, 4, 5, 5, 5, 5, 6, 6, 7, 8, 8, 9, 9, 9, 10, 10, 10},
     {0, 0, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10}, {0, 0, 0, 4, 4, 4, 4, 4, 5, 6, 7, 8, 8, 9, 10, 10, 10}}};
static const uint8_t kTZ[15][16] = {
    {1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9}, {3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6}, {4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6},
    {5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5}, {4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5}, {6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6}, {6, 5, 3, 3, 3, 2, 3, 4, 3, 6},
    {6, 4, 5, 3, 2, 2, 3, 3, 6}, {6, 6, 4, 2, 2, 3, 2, 5}, {5, 5, 3, 2, 2, 2, 4}, {4, 4, 3, 3, 1, 3}, {4, 4, 2, 1, 3}, {3, 3, 1, 2}, {2, 2, 1}, {1, 1}};
static const uint8_t kRB[7][15] = {{1, 1}, {1, 2, 2}, {2, 2, 2, 2}, {2, 2, 2, 3, 3}, {2, 2, 3, 3, 3, 3}, {2, 3, 3, 3, 3, 3, 3}, {3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11}};

HLO_API int hlo_cavlc_bits(const int32_t* lv, int max_num_coeff, int nC, int32_t* single_ctr, int32_t* total_coeff)
{
    int nz[16], run[16] = {0}, tc = 0, t1 = 0, tz = 0, k = -1, j, bits = 0, cnt_t1 = 1, seen = 0, sl, zl = 0;
    for (j = 0; j < max_num_coeff; ++j) {          /* residual.c:761-784 (reverse scan) */
        int c = lv[max_num_coeff - 1 - j];
        if (c) {
            nz[tc++] = c; seen = 1; ++k;
            if (cnt_t1) { if (c == 1 || c == -1) { ++t1; cnt_t1 = t1 < 3; } else cnt_t1 = 0; }
        } else if (seen) { ++run[k]; ++tz; }
    }
    bits += nC >= 8 ? 6 : kCT[nC < 2 ? 0 : (nC < 4 ? 1 : 2)][t1][tc];     /* cavlc.c:652-706 */
    *total_coeff = tc;
    *single_ctr = 9;                                /* residual.c:883 (only meaningful when tc > 0) */
    if (tc > 0) {
        static const int thr[7] = {0, 3, 6, 12, 24, 48, 1 << 15};
        sl = (tc > 10 && t1 < 3) ? 1 : 0;
        for (j = 0; j < tc; ++j) {                  /* residual.c:819-859 */
            int lc, prefix, ssz;
            if (j < t1) { bits += 1; continue; }
            lc = nz[j] > 0 ? (nz[j] << 1) - 2 : -(nz[j] << 1) - 1;
            if (j == t1 && t1 < 3 && lc >= 2) lc -= 2;
            /* cavlc.c:59-104: prefix/suffix sizes */
            if (sl == 0) { if (lc < 14) { prefix = lc; ssz = 0; } else if (lc < 30) { prefix = 14; ssz = 4; } else { prefix = 15; ssz = 12; } }
            else { prefix = lc >> sl; if (prefix < 15) ssz = sl; else { prefix = 15; ssz = 12; } }
            bits += prefix + 1 + ssz;
            if (sl == 0) sl = 1;
            if (abs(nz[j]) > thr[sl]) ++sl;
        }
        if (tc < max_num_coeff) { bits += kTZ[tc - 1][tz]; zl = tz; }        /* residual.c:862-873 */
        for (j = 0; j < tc - 1 && zl > 0; ++j) { bits += kRB[(zl > 7 ? 7 : zl) - 1][run[j]]; zl -= run[j]; }
        if (tc == 1 && abs(nz[0]) == 1) {           /* residual.c:884-896 */
            static const int T[6] = {3, 2, 2, 1, 1, 1};
            int rn = zl > 0 ? run[0] : 0;
            *single_ctr = rn < 6 ? T[rn] : 0;
        }
    }
    return bits;
}

/* nC from neighbour counts (residual.c:742-754); nA/nB < 0 = not available */
HLO_API int hlo_nC(int nA, int nB)
{
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    if (nA >= 0) return nA;
    if (nB >= 0) return nB;
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * One luma 4x4 trial encode as done for every ME candidate: hl_codec_264_me_ds_mb_compute_cost_mode, me_ds.c:607-670
 * -> hl_codec_264_rdo_mb_compute_inter_luma4x4, rdo.c:2784-2830.
 * src/pred: 4x4 u8 with strides.  Returns distortion (SAD); *nonzero, levels[16] (zig-zag) filled when nonzero. */
HLO_API int hlo_trial_luma4x4(const uint8_t* src, int ss, const uint8_t* pred, int ps, int qp, int32_t* levels, int32_t* nonzero)
{
    int32_t res[16], w[16], z[16], c[16], r[16];
    uint8_t p4[16], rec[16];
    int x, y, all0 = 1;
    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) { res[y * 4 + x] = (int)src[y * ss + x] - (int)pred[y * ps + x]; p4[y * 4 + x] = pred[y * ps + x]; if (res[y * 4 + x]) all0 = 0; }
    if (!all0) {
        hlo_fwd4x4(res, w);
        hlo_quant4x4(qp, 0, w, z);
        hlo_zigzag(z, levels);
        all0 = 1;
        for (x = 0; x < 16; ++x) if (levels[x]) all0 = 0;
    }
    *nonzero = !all0;
    if (all0) return hlo_sad4x4(src, ss, pred, ps);
    hlo_inv_zigzag(levels, c);
    hlo_dequant_inv4x4(qp, 0, c, r);
    hlo_addclip_u8xi32(p4, r, rec);                /* wraps: me_ds.c:636-639 */
    return hlo_sad4x4(src, ss, rec, 4);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Residual coding + reconstruction of one INTER macroblock given its prediction:
 *   luma   _hl_codec_264_rdo_mb_reconstruct_inter, rdo.c:2428-2478 (the "Single_ctr_luma >= 6" branch)
 *   chroma _hl_codec_264_rdo_mb_reconstruct_chroma, rdo.c:2502-2700 and hl_codec_264_transf_decode_chroma, transf.c:161-296
 * Planes are tight (pitch W for luma, W/2 for chroma).  chroma_ac is IN/OUT: it is the persistent ChromaACLevel of
 * the macroblock object (the reference does not clear it for blocks whose residual is all zero, and uses it whenever
 * the de-quantised DC of the block is non-zero -- transf.c:236-245).  Pass zeros for a stateless call. */
static void recon_inter_mb_impl(const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* pred_y, const uint8_t* pred_u,
                                const uint8_t* pred_v, int W, int mbx, int mby, int qp, int qpc, int mb_is_intra, int luma_intra_f, int16_t* luma_level /*[16][16]*/,
                                int16_t* chroma_dc /*[2][4]*/, int16_t* chroma_ac /*[2][4][16]*/, int32_t* cbp_luma4x4, int32_t* cbp_dc /*[2]*/,
                                int32_t* cbp_ac /*[2]*/, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v);
HLO_API void hlo_recon_inter_mb(const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* pred_y, const uint8_t* pred_u,
                                const uint8_t* pred_v, int W, int mbx, int mby, int qp, int qpc, int mb_is_intra, int16_t* luma_level /*[16][16]*/,
                                int16_t* chroma_dc /*[2][4]*/, int16_t* chroma_ac /*[2][4][16]*/, int32_t* cbp_luma4x4, int32_t* cbp_dc /*[2]*/,
                                int32_t* cbp_ac /*[2]*/, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v)
{
    recon_inter_mb_impl(src_y, src_u, src_v, pred_y, pred_u, pred_v, W, mbx, mby, qp, qpc, mb_is_intra, 0, luma_level, chroma_dc, chroma_ac, cbp_luma4x4, cbp_dc, cbp_ac,
                        rec_y, rec_u, rec_v);
}
/* SVC enhancement-layer inter macroblock (base_mode_flag = 1), SURVEY 8a row a14: hl_codec_264_rdo_mb_guess_best_inter_pred_svc, rdo.c:1273-1521.
 * After the (host-side, serial) inter-layer derivation of partitions / motion vectors and the interpolation of the prediction, the residual
 * coding differs from the AVC inter macroblock in ONE thing: the luma coefficients are quantised with the INTRA rounding offset
 * (__isIntraBlockTrue, rdo.c:1468); de-quantisation takes the coefficient block before scanning (rdo.c:1483), which is the same data.
 * Chroma is the shared _hl_codec_264_rdo_mb_reconstruct_chroma (rdo.c:1500) with the macroblock flagged INFERRED = inter (mb.h:46). */
HLO_API void hlo_recon_svc_inter_mb(const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* pred_y, const uint8_t* pred_u,
                                    const uint8_t* pred_v, int W, int mbx, int mby, int qp, int qpc, int16_t* luma_level /*[16][16]*/,
                                    int16_t* chroma_dc /*[2][4]*/, int16_t* chroma_ac /*[2][4][16]*/, int32_t* cbp_luma4x4, int32_t* cbp_dc /*[2]*/,
                                    int32_t* cbp_ac /*[2]*/, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v)
{
    recon_inter_mb_impl(src_y, src_u, src_v, pred_y, pred_u, pred_v, W, mbx, mby, qp, qpc, 0, 1, luma_level, chroma_dc, chroma_ac, cbp_luma4x4, cbp_dc, cbp_ac,
                        rec_y, rec_u, rec_v);
}
/* SVC enhancement-layer I_BL macroblock (enhancement I pictures): hl_codec_264_rdo_mb_guess_best_intra_pred_svc, rdo.c:301-461.  pred_* = the base-layer
 * reconstruction resampled by G.8.6.2.1 (decode_svc.c:216, host side).  Luma as above (intra offset, rdo.c:404); in the shared chroma function the macroblock
 * now counts as INTRA (the SVC initialisation process flags an I_BL macroblock intra), so the 2x2 DC quantisation uses the intra offset too (rdo.c:2660). */
HLO_API void hlo_recon_svc_bl_mb(const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* pred_y, const uint8_t* pred_u,
                                 const uint8_t* pred_v, int W, int mbx, int mby, int qp, int qpc, int16_t* luma_level /*[16][16]*/,
                                 int16_t* chroma_dc /*[2][4]*/, int16_t* chroma_ac /*[2][4][16]*/, int32_t* cbp_luma4x4, int32_t* cbp_dc /*[2]*/,
                                 int32_t* cbp_ac /*[2]*/, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v)
{
    recon_inter_mb_impl(src_y, src_u, src_v, pred_y, pred_u, pred_v, W, mbx, mby, qp, qpc, 1, 1, luma_level, chroma_dc, chroma_ac, cbp_luma4x4, cbp_dc, cbp_ac,
                        rec_y, rec_u, rec_v);
}
static void recon_inter_mb_impl(const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* pred_y, const uint8_t* pred_u,
                                const uint8_t* pred_v, int W, int mbx, int mby, int qp, int qpc, int mb_is_intra, int luma_intra_f, int16_t* luma_level /*[16][16]*/,
                                int16_t* chroma_dc /*[2][4]*/, int16_t* chroma_ac /*[2][4][16]*/, int32_t* cbp_luma4x4, int32_t* cbp_dc /*[2]*/,
                                int32_t* cbp_ac /*[2]*/, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v)
{
    int b, i, x, y, c, Wc = W >> 1;
    *cbp_luma4x4 = 0;
    for (b = 0; b < 16; ++b) {                                        /* rdo.c:2428 */
        int xo = ((b >> 2) & 1) * 8 + (b & 1) * 4, yo = (b >> 3) * 8 + ((b >> 1) & 1) * 4;
        const uint8_t* s = src_y + (mby * 16 + yo) * W + mbx * 16 + xo;
        const uint8_t* p = pred_y + (mby * 16 + yo) * W + mbx * 16 + xo;
        uint8_t* r = rec_y + (mby * 16 + yo) * W + mbx * 16 + xo;
        int32_t res[16], w[16], z[16], lv[16], cc[16], rr[16];
        int all0 = 1;
        for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) { res[y * 4 + x] = (int)s[y * W + x] - (int)p[y * W + x]; if (res[y * 4 + x]) all0 = 0; }
        if (!all0) {
            hlo_fwd4x4(res, w);
            hlo_quant4x4(qp, luma_intra_f, w, z);
            all0 = 1;
            for (i = 0; i < 16; ++i) if (z[i]) all0 = 0;
            if (!all0) { hlo_zigzag(z, lv); *cbp_luma4x4 |= 1 << b; }
        }
        if (all0) for (i = 0; i < 16; ++i) lv[i] = 0;
        for (i = 0; i < 16; ++i) luma_level[b * 16 + i] = (int16_t)lv[i];
        if (*cbp_luma4x4 & (1 << b)) {
            hlo_inv_zigzag(lv, cc);
            hlo_dequant_inv4x4(qp, 0, cc, rr);
            for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) r[y * W + x] = (uint8_t)clip1((int)p[y * W + x] + rr[y * 4 + x]);
        } else {
            for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) r[y * W + x] = p[y * W + x];
        }
    }
    /* chroma, rdo.c:2561-2672 */
    {
        int32_t dccoef[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, single[2] = {0, 0}, totc[2] = {0, 0};
        cbp_ac[0] = cbp_ac[1] = cbp_dc[0] = cbp_dc[1] = 0;
        for (b = 0; b < 4; ++b) {
            int xo = (b & 1) * 4, yo = (b >> 1) * 4;
            for (c = 0; c < 2; ++c) {
                const uint8_t* s = (c ? src_v : src_u) + (mby * 8 + yo) * Wc + mbx * 8 + xo;
                const uint8_t* p = (c ? pred_v : pred_u) + (mby * 8 + yo) * Wc + mbx * 8 + xo;
                int16_t* ac = chroma_ac + (c * 4 + b) * 16;
                int32_t res[16], w[16], z[16], lv[16];
                int all0 = 1;
                for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) { res[y * 4 + x] = (int)s[y * Wc + x] - (int)p[y * Wc + x]; if (res[y * 4 + x]) all0 = 0; }
                if (!all0) {
                    hlo_fwd4x4(res, w);
                    hlo_quant4x4(qpc, 1, w, z);                      /* always the intra offset: rdo.c:2588,2618 */
                    hlo_zigzag(z, lv);
                    for (i = 1; i < 16; ++i) ac[i - 1] = (int16_t)lv[i];   /* Scan4x4_AC_C, utils.h:183 */
                    all0 = 1;
                    for (i = 0; i < 16; ++i) if (ac[i]) all0 = 0;    /* allzero16 over [0..15]; [15] is never written */
                    dccoef[c][b] = w[0];
                    cbp_ac[c] |= all0 ? 0 : (1 << b);
                    cbp_dc[c] |= w[0] ? (1 << b) : 0;
                } else dccoef[c][b] = 0;
                if (single[c] < 7 && (cbp_ac[c] & (1 << b))) {      /* rdo.c:2599-2607 */
                    int32_t l16[16], sc, tc;
                    for (i = 0; i < 16; ++i) l16[i] = ac[i];
                    hlo_cavlc_bits(l16, 16, 0, &sc, &tc);
                    single[c] += sc; totc[c] += tc;
                }
            }
        }
        for (c = 0; c < 2; ++c) if (single[c] < 7 && totc[c] == 1) cbp_ac[c] = 0;   /* rdo.c:2641-2649 */
        if (cbp_dc[0] || cbp_dc[1]) {
            for (c = 0; c < 2; ++c) if (cbp_dc[c]) {
                int32_t h4[4], q4[4];
                hlo_hadamard2x2(dccoef[c], h4);
                hlo_quant_dc(qpc, mb_is_intra, h4, q4, 4);           /* rdo.c:2660: isIntra of the MB */
                for (i = 0; i < 4; ++i) chroma_dc[c * 4 + i] = (int16_t)q4[i];
                cbp_dc[c] = (q4[0] ? 1 : 0) | (q4[1] ? 2 : 0) | (q4[2] ? 4 : 0) | (q4[3] ? 8 : 0);
            }
        }
        /* transf.c:161-296 */
        for (c = 0; c < 2; ++c) {
            const uint8_t* pp = (c ? pred_v : pred_u);
            uint8_t* rp = (c ? rec_v : rec_u);
            int32_t dcC[4] = {0, 0, 0, 0};
            if (cbp_dc[c]) { int32_t l4[4]; for (i = 0; i < 4; ++i) l4[i] = chroma_dc[c * 4 + i]; hlo_scale_chroma_dc(qpc, l4, dcC); }
            for (b = 0; b < 4; ++b) {
                int xo = (b & 1) * 4, yo = (b >> 1) * 4;
                const uint8_t* p = pp + (mby * 8 + yo) * Wc + mbx * 8 + xo;
                uint8_t* r = rp + (mby * 8 + yo) * Wc + mbx * 8 + xo;
                int use = (cbp_dc[c] || cbp_ac[c]) && (dcC[b] || (cbp_ac[c] & (1 << b)));
                if (use) {
                    int32_t l16[16], cc[16], rr[16];
                    l16[0] = dcC[b];
                    for (i = 1; i < 16; ++i) l16[i] = chroma_ac[(c * 4 + b) * 16 + i - 1];
                    hlo_inv_zigzag(l16, cc);
                    hlo_dequant_inv4x4(qpc, 1, cc, rr);
                    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) r[y * Wc + x] = (uint8_t)clip1((int)p[y * Wc + x] + rr[y * 4 + x]);
                } else {
                    for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) r[y * Wc + x] = p[y * Wc + x];
                }
            }
        }
    }
}

/* Whole ME candidate (one partition at one MV): hl_codec_264_me_ds_mb_compute_cost_mode, me_ds.c:527-688, without the
 * coeff_token bits (they depend on encoder history).  Outputs: dist, bits_rest, single_ctr, cbp4x4 and
 * per-block TotalCoeff / TrailingOnes indexed by luma4x4BlkIdx. */
HLO_API void hlo_me_cost(const uint8_t* src, const uint8_t* ref, int W, int H, int qp, int mbx, int mby, int px, int py, int pw, int ph, int mvx, int mvy,
                         int32_t* dist, int32_t* bits_rest, int32_t* single_ctr, int32_t* cbp, uint8_t* total_coeff /*16*/, uint8_t* trailing_ones /*16*/)
{
    uint8_t pred[256];
    int bx, by, i;
    *dist = *bits_rest = *single_ctr = *cbp = 0;
    for (i = 0; i < 16; ++i) total_coeff[i] = trailing_ones[i] = 0;
    hlo_interp_luma(ref, W, H, mbx * 16 + px, mby * 16 + py, pw, ph, mvx, mvy, pred);
    for (by = 0; by < ph; by += 4) for (bx = 0; bx < pw; bx += 4) {
        int32_t lv[16], nz, sc, tc, x = px + bx, y = py + by;
        int blk = ((y >> 3) << 3) | ((x >> 3) << 2) | (((y >> 2) & 1) << 1) | ((x >> 2) & 1);
        *dist += hlo_trial_luma4x4(src + (mby * 16 + y) * W + mbx * 16 + x, W, pred + by * 16 + bx, 16, qp, lv, &nz);
        if (nz) {
            int t1 = 0, k;
            int full = hlo_cavlc_bits(lv, 16, 0, &sc, &tc);
            /* trailing ones, residual.c:768-776 */
            for (k = 15; k >= 0; --k) { if (!lv[k]) continue; if ((lv[k] == 1 || lv[k] == -1) && t1 < 3) ++t1; else break; }
            *bits_rest += full - kCT[0][t1][tc];
            *single_ctr += sc;
            *cbp |= 1 << blk;
            total_coeff[blk] = (uint8_t)tc; trailing_ones[blk] = (uint8_t)t1;
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Batch drivers used by bench.py's cpu_baseline leg ("port") and by the parity tests at larger sizes: plain loops over
 * the functions above, no new arithmetic. */
/* cands: n x 8 int32 {mb_x, mb_y, part_x, part_y, part_w, part_h, mv_x, mv_y}; out: n x 4 int32 {dist, bits_rest, single_ctr, cbp} */
HLO_API void hlo_me_cost_batch(const uint8_t* src, const uint8_t* ref, int W, int H, int qp, const int32_t* cands, int n, int32_t* out)
{
    int i;
    uint8_t tc[16], t1[16];
    for (i = 0; i < n; ++i) {
        const int32_t* c = cands + 8 * i;
        hlo_me_cost(src, ref, W, H, qp, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], out + 4 * i, out + 4 * i + 1, out + 4 * i + 2, out + 4 * i + 3, tc, t1);
    }
}
/* Prediction (luma + chroma) and residual coding / reconstruction of macroblocks [mb_begin, mb_end) of a frame.
 * parts: per MB 16 entries x 7 int32 {valid, ox, oy, w, h, mvx, mvy} (list of the MB's partitions, valid=0 terminates).
 * Writes pred and rec planes (tight, Y|U|V) and returns the number of non-zero luma 4x4 blocks (a checksum for callers). */
HLO_API int hlo_predict_recon_mbs(const uint8_t* src_yuv, const uint8_t* ref_yuv, int W, int H, int qp, int qpc, const int32_t* parts, int mb_begin, int mb_end,
                                  uint8_t* pred_yuv, uint8_t* rec_yuv)
{
    const int Wc = W >> 1, Hc = H >> 1, mbw = W >> 4;
    const uint8_t *sy = src_yuv, *su = sy + W * H, *sv = su + Wc * Hc, *ry = ref_yuv, *ru = ry + W * H, *rv = ru + Wc * Hc;
    uint8_t *py = pred_yuv, *pu = py + W * H, *pv = pu + Wc * Hc, *oy = rec_yuv, *ou = oy + W * H, *ov = ou + Wc * Hc;
    int mb, k, x, y, nz = 0;
    for (mb = mb_begin; mb < mb_end; ++mb) {
        const int mbx = mb % mbw, mby = mb / mbw;
        int16_t ll[256], dc[8], ac[128];
        int32_t c4, cdc[2], cac[2];
        memset(ac, 0, sizeof(ac));
        for (k = 0; k < 16; ++k) {
            const int32_t* p = parts + ((size_t)mb * 16 + k) * 7;
            uint8_t tmp[256], tcu[64], tcv[64];
            int xl, yl;
            if (!p[0]) break;
            xl = mbx * 16 + p[1]; yl = mby * 16 + p[2];
            hlo_interp_luma(ry, W, H, xl, yl, p[3], p[4], p[5], p[6], tmp);
            for (y = 0; y < p[4]; ++y) for (x = 0; x < p[3]; ++x) py[(yl + y) * W + xl + x] = tmp[y * 16 + x];
            hlo_interp_chroma(ru, Wc, Hc, xl, yl, p[3] >> 1, p[4] >> 1, p[5], p[6], tcu);
            hlo_interp_chroma(rv, Wc, Hc, xl, yl, p[3] >> 1, p[4] >> 1, p[5], p[6], tcv);
            for (y = 0; y < (p[4] >> 1); ++y) for (x = 0; x < (p[3] >> 1); ++x) {
                pu[((yl >> 1) + y) * Wc + (xl >> 1) + x] = tcu[y * 8 + x];
                pv[((yl >> 1) + y) * Wc + (xl >> 1) + x] = tcv[y * 8 + x];
            }
        }
        hlo_recon_inter_mb(sy, su, sv, py, pu, pv, W, mbx, mby, qp, qpc, 0, ll, dc, ac, &c4, cdc, cac, oy, ou, ov);
        for (k = 0; k < 16; ++k) nz += (c4 >> k) & 1;
    }
    return nz;
}

/* ---- SVC: resampling of the reference layer's reconstruction into the Intra_Base (I_BL) prediction of an enhancement-layer I picture ----
 * _hl_codec_264_decode_svc_resample_intra_colour_comps decode_svc.c:2864 -> ..._ref_layer_array_construct_prior_to_intra_resampling :2952 (G.8.6.2.2; every
 * reference macroblock of an I picture is intra, so every sample is available and the array is a clamped gather, (G-280)/(G-281)) -> ..._interpol_intra_base :3071
 * (G.8.6.2.3: vertical pass (G-301) then horizontal pass + clip (G-305); chroma takes the two-tap table whatever filteringModeFlag says, :3116,:3150), with the
 * sample locations of hl_codec_264_utils_derivation_process_for_ref_layer_sample_locs_in_resampling_svc utils.c:1064-1157 (G.6.3) for frame macroblocks, no
 * cropping offsets, chroma phases as the encoder's SPS sets them (sps.c:810-813: phaseX = phaseY = refPhaseX = refPhaseY = 0).  The reference works per macroblock
 * on a local window; window offsets (G-270..G-273) are multiples of 16 samples, so phases and absolute positions do not depend on the macroblock and the process is a
 * function of the plane.  Tables: include/hartallo/h264/hl_codec_264_tables.h:624-664 (Table G-9 luma; bilinear chroma). */
static const int8_t kSvcLumaF[16][4] = { {0, 32, 0, 0}, {-1, 32, 2, -1}, {-2, 31, 4, -1}, {-3, 30, 6, -1}, {-3, 28, 8, -1}, {-4, 26, 11, -1}, {-4, 24, 14, -2},
    {-3, 22, 16, -3}, {-3, 19, 19, -3}, {-3, 16, 22, -3}, {-2, 14, 24, -4}, {-1, 11, 26, -4}, {-1, 8, 28, -3}, {-1, 6, 30, -3}, {-1, 4, 31, -2}, {-1, 2, 32, -1} };

static int svc_ceil_log2(int v) { int n = 0; while ((1 << n) < v) ++n; return n; }

/* (G-43)..(G-49),(G-59): 1/16-sample position in the reference plane of sample p of the current plane */
static int svc_ref16(int p, int refDim, int scaledDim, int level_idc)
{
    const int shift = level_idc <= 30 ? 16 : 31 - svc_ceil_log2(refDim);
    const int scale = ((refDim << shift) + (scaledDim >> 1)) / scaledDim;
    const int add = (((refDim * 2) << (shift - 2)) + (scaledDim >> 1)) / scaledDim + (1 << (shift - 5));
    return ((p * scale + add) >> (shift - 4)) - 8;
}

/* one plane: ref is refW x refH, out is W x H (both tight); chroma != 0 selects the two-tap filter */
HLO_API void hlo_svc_resample_intra_plane(const uint8_t* ref, int refW, int refH, int W, int H, int chroma, int level_idc, uint8_t* out)
{
    int x, y, k, j;
    for (y = 0; y < H; ++y) {
        const int y16 = svc_ref16(y, refH, H, level_idc), yr = y16 >> 4, yp = y16 & 15;
        for (x = 0; x < W; ++x) {
            const int x16 = svc_ref16(x, refW, W, level_idc), xr = x16 >> 4, xp = x16 & 15;
            int v = 0;
            if (chroma) {
                for (j = 0; j < 2; ++j) {
                    int t = 0;
                    for (k = 0; k < 2; ++k) t += (k ? 2 * yp : 32 - 2 * yp) * ref[clip3(0, refH - 1, yr + k) * refW + clip3(0, refW - 1, xr + j)];
                    v += (j ? 2 * xp : 32 - 2 * xp) * t;
                }
            }
            else {
                for (j = 0; j < 4; ++j) {
                    int t = 0;
                    for (k = 0; k < 4; ++k) t += kSvcLumaF[yp][k] * ref[clip3(0, refH - 1, yr - 1 + k) * refW + clip3(0, refW - 1, xr - 1 + j)];
                    v += kSvcLumaF[xp][j] * t;
                }
            }
            out[y * W + x] = (uint8_t)clip3(0, 255, (v + 512) >> 10);
        }
    }
}

/* whole picture, tight Y|U|V in and out (4:2:0) */
HLO_API void hlo_svc_resample_intra_yuv(const uint8_t* ref_yuv, int refW, int refH, int W, int H, int level_idc, uint8_t* out_yuv)
{
    const int rc = (refW >> 1) * (refH >> 1), oc = (W >> 1) * (H >> 1);
    hlo_svc_resample_intra_plane(ref_yuv, refW, refH, W, H, 0, level_idc, out_yuv);
    hlo_svc_resample_intra_plane(ref_yuv + refW * refH, refW >> 1, refH >> 1, W >> 1, H >> 1, 1, level_idc, out_yuv + W * H);
    hlo_svc_resample_intra_plane(ref_yuv + refW * refH + rc, refW >> 1, refH >> 1, W >> 1, H >> 1, 1, level_idc, out_yuv + W * H + oc);
}

/* --- SVC inter-layer motion derivation for an enhancement-layer P macroblock with base_mode_flag = 1 -------------------------------------------
 * What rdo.c:1318-1346 has computed before it predicts: hl_codec_264_utils_derivation_process_initialisation_svc (utils.c:1225) followed by
 * hl_codec_264_utils_derivation_process_for_mv_comps_and_ref_indices_svc (utils.c:1498).  Restated clause by clause with the reference's own intermediate arrays
 * (refLayerPartIdc, tempRefIdxPredL0, mvILPredL0, refIdxILPredL0), for frame macroblocks, EP slices and CroppingChangeFlag = 0: the restricted case (layers of equal
 * or doubled size) and the general one (run live with layers scaled 3:2, oracle/ref_driver.c --scale).  PINNED by tests/test_svc_derive.py against the
 * reference's trace (oracle/ref_driver.c tags 11 and 6: tests/golden/svc_derive.npz and, in the build container, live runs).
 * base records: 53 int32 per reference-layer macroblock in the order of tag 11 (intra by e_type, flags_type intra, e_type P_8X8 / P_8X8REF0, MbPartWidth, MbPartHeight,
 * SubMbPartWidth[4], SubMbPartHeight[4], predFlagL0[4], refIdxL0[4], mvL0[4][4][2]).
 * out16: [0] intraILPredFlag, [1] NumMbPart, [2] MbPartWidth, [3] MbPartHeight, [4..7] NumSubMbPart, [8..11] SubMbPartWidth, [12..15] SubMbPartHeight (entries of the
 * partitions that exist); ref_idx4 = refIdxL0[4]; mv32 = mvL0[4][4][2].
 * Returns 0, or -1 when a block maps outside the reference layer / onto an object the reference would divide by zero on / intra and inter blocks mix. */
static int svc_il_floor_div(int a, int b) { return a / b; }   /* operands are non-negative wherever the reference divides */
HLO_API int hlo_svc_derive_mb(const int32_t* base, int ref_w, int ref_h, int scaled_w, int scaled_h, int off_x, int off_y, int level_idc, int restricted, int mb_x, int mb_y,
                              int32_t* out16, int32_t* ref_idx4, int32_t* mv32)
{
    const int ref_mbw = ref_w >> 4, nref = ref_mbw * (ref_h >> 4);
    /* utils.c:988-993 (G-7..G-10) */
    const int shiftX = level_idc <= 30 ? 16 : 31 - svc_ceil_log2(ref_w), shiftY = level_idc <= 30 ? 16 : 31 - svc_ceil_log2(ref_h);
    const int scaleX = (int)((((long long)ref_w << shiftX) + (scaled_w >> 1)) / scaled_w), scaleY = (int)((((long long)ref_h << shiftY) + (scaled_h >> 1)) / scaled_h);
    int refLayerPartIdc[4][4], tempRefIdxPredL0[4][4], mvILPredL0[4][4][2], refIdxILPredL0[2][2];
    int x, y, intraILPredFlag = 1, any_intra = 0, i;
    memset(out16, 0, 16 * sizeof(int32_t)); memset(ref_idx4, 0, 4 * sizeof(int32_t)); memset(mv32, 0, 32 * sizeof(int32_t));
    /* G.8.6.1.1, utils.c:1690-1711 */
    for (y = 0; y < 4; ++y)
        for (x = 0; x < 4; ++x) {
            const int xP = (x << 2) + 1, yP = (y << 2) + 1;
            /* G.6.1, utils.c:995-1003 (frame macroblock in a frame picture: yC = yM + yP) */
            const int xC = mb_x * 16 + xP, yC = mb_y * 16 + yP;
            int xRef = (int)(((unsigned)(xC - off_x) * (unsigned)scaleX + (1u << (shiftX - 1)))) >> shiftX;
            int yRef = (int)(((unsigned)(yC - off_y) * (unsigned)scaleY + (1u << (shiftY - 1)))) >> shiftY;
            int addr, xB, yB, mbPartIdx, subMbPartIdx;
            const int32_t* b;
            if (xRef > ref_w - 1) xRef = ref_w - 1;
            if (yRef > ref_h - 1) yRef = ref_h - 1;
            if (xRef < 0 || yRef < 0) return -1;
            addr = (yRef >> 4) * ref_mbw + (xRef >> 4);   /* (G-15) */
            if (addr >= nref) return -1;
            xB = xRef & 15; yB = yRef & 15;
            b = base + 53 * addr;
            if (b[0]) { refLayerPartIdc[y][x] = -1; any_intra = 1; continue; }   /* utils.c:1701-1703 */
            /* G.6.4 -> 6.4.12.4 as mb.h:313-339 */
            if (b[1]) mbPartIdx = 0;
            else {
                if (b[3] <= 0 || b[4] <= 0) return -1;
                mbPartIdx = svc_il_floor_div(16, b[3]) * svc_il_floor_div(yB, b[4]) + svc_il_floor_div(xB, b[3]);
            }
            if (mbPartIdx > 3) return -1;
            if (!b[2]) subMbPartIdx = 0;
            else {
                const int sw = b[5 + mbPartIdx], sh = b[9 + mbPartIdx];
                if (sw <= 0 || sh <= 0) return -1;
                subMbPartIdx = svc_il_floor_div(8, sw) * svc_il_floor_div(yB % 8, sh) + svc_il_floor_div(xB % 8, sw);
            }
            if (subMbPartIdx > 3) return -1;
            refLayerPartIdc[y][x] = (addr << 4) + (mbPartIdx << 2) + subMbPartIdx;   /* (G-209) */
            intraILPredFlag = 0;
        }
    if (intraILPredFlag) {   /* utils.c:1272-1287: I_BL, G.8.4.1 then clears the motion (utils.c:1510-1528) */
        out16[0] = 1; out16[1] = 1; out16[2] = out16[3] = 16;
        for (i = 0; i < 4; ++i) ref_idx4[i] = -1;
        return 0;
    }
    if (any_intra && restricted) return -1;   /* cannot happen with aligned layers of equal or doubled size */
    if (!restricted) {
        /* utils.c:1713-1774, transcribed with its two departures from G.8.6.1.1: procI4x4Blk lives across the four 8x8 blocks, and the first 4x4 test reads column xO + 1 with `== -1` */
        int procI4x4Blk[2][2] = {{0, 0}, {0, 0}}, procI8x8Blk[2][2] = {{0, 0}, {0, 0}};
        int xP, yP, xS, yS;
        int (*idc)[4] = refLayerPartIdc;
        for (yP = 0; yP < 2; ++yP)
            for (xP = 0; xP < 2; ++xP) {
                const int xO = xP << 1, yO = yP << 1;
                for (yS = 0; yS < 2; ++yS)
                    for (xS = 0; xS < 2; ++xS)
                        if (idc[yO + yS][xO + xS] == -1) {
                            procI4x4Blk[yS][xS] = 1;
                            if (procI4x4Blk[yS][1 - xS] == 0 && idc[yO + yS][xO + 1] == -1) idc[yO + yS][xO + xS] = idc[yO + yS][xO + 1 - xS];
                            else if (procI4x4Blk[1 - yS][xS] == 0 && idc[yO + 1 - yS][xO + xS] != -1) idc[yO + yS][xO + xS] = idc[yO + 1 - yS][xO + xS];
                            else if (procI4x4Blk[1 - yS][1 - xS] == 0 && idc[yO + 1 - yS][xO + 1 - xS] != -1) idc[yO + yS][xO + xS] = idc[yO + 1 - yS][xO + 1 - xS];
                        }
            }
        for (yP = 0; yP < 2; ++yP)
            for (xP = 0; xP < 2; ++xP)
                if (idc[yP << 1][xP << 1] == -1) {
                    procI8x8Blk[yP][xP] = 1;
                    if (procI8x8Blk[yP][1 - xP] == 0 && idc[yP << 1][2 - xP] != -1) {
                        for (yS = 0; yS < 2; ++yS) for (xS = 0; xS < 2; ++xS) idc[(yP << 1) + yS][(xP << 1) + xS] = idc[(yP << 1) + yS][2 - xP];
                    }
                    else if (procI8x8Blk[1 - yP][xP] == 0 && idc[2 - yP][xP << 1] != -1) {
                        for (yS = 0; yS < 2; ++yS) for (xS = 0; xS < 2; ++xS) idc[(yP << 1) + yS][(xP << 1) + xS] = idc[2 - yP][(xP << 1) + xS];
                    }
                    else if (procI8x8Blk[1 - yP][1 - xP] == 0 && idc[2 - yP][2 - xP] != -1) {
                        for (yS = 0; yS < 2; ++yS) for (xS = 0; xS < 2; ++xS) idc[(yP << 1) + yS][(xP << 1) + xS] = idc[2 - yP][2 - xP];
                    }
                }
        for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) if (idc[y][x] == -1) return -1;   /* utils.c:1797 would index the macroblock list with -1 */
    }
    /* G.8.6.1.2, utils.c:1793-1880 */
    {
        const int mvScaleX = (int)((((long long)scaled_w << 16) + (ref_w >> 1)) / ref_w), mvScaleY = (int)((((long long)scaled_h << 16) + (ref_h >> 1)) / ref_h);   /* (G-232), (G-233) */
        for (y = 0; y < 4; ++y)
            for (x = 0; x < 4; ++x) {
                const int idc = refLayerPartIdc[y][x], refMbAddr = idc >> 4, refMbPartIdx = (idc & 15) >> 2, refSubMbPartIdx = idc & 3;   /* (G-219)..(G-221) */
                const int32_t* b = base + 53 * refMbAddr;
                if (b[13 + refMbPartIdx] == 0) { tempRefIdxPredL0[y][x] = -1; mvILPredL0[y][x][0] = mvILPredL0[y][x][1] = 0; }   /* (G-216)..(G-218) */
                else {
                    const int32_t* mv = b + 21 + (refMbPartIdx * 4 + refSubMbPartIdx) * 2;
                    tempRefIdxPredL0[y][x] = b[17 + refMbPartIdx];                 /* (G-222) */
                    mvILPredL0[y][x][0] = (mv[0] * mvScaleX + 32768) >> 16;        /* (G-234), (G-242) */
                    mvILPredL0[y][x][1] = (mv[1] * mvScaleY + 32768) >> 16;        /* (G-235), (G-243) */
                }
            }
        for (y = 0; y < 2; ++y) for (x = 0; x < 2; ++x) refIdxILPredL0[y][x] = tempRefIdxPredL0[y << 1][x << 1];   /* utils.c:1888; the rest is skipped when restricted */
        if (!restricted) {
            int yP, xP, yS, xS, k;
            /* utils.c:1889-1912: the minimum (G-244) is updated INSIDE the walk over the four blocks, each block is compared with the value reached so far */
            for (yP = 0; yP < 2; ++yP)
                for (xP = 0; xP < 2; ++xP)
                    for (yS = 0; yS < 2; ++yS)
                        for (xS = 0; xS < 2; ++xS) {
                            const int t = tempRefIdxPredL0[2 * yP + yS][2 * xP + xS];
                            int* r = &refIdxILPredL0[yP][xP];
                            *r = (*r >= 0 && t >= 0) ? (*r < t ? *r : t) : (*r > t ? *r : t);   /* HL_MATH_MIN_POSITIVE, hl_math.h:28 */
                            if (t != *r) {
                                int sy, sx;
                                if (tempRefIdxPredL0[2 * yP + yS][2 * xP + 1 - xS] == *r) { sy = 2 * yP + yS; sx = 2 * xP + 1 - xS; }            /* (G-246) */
                                else if (tempRefIdxPredL0[2 * yP + 1 - yS][2 * xP + xS] == *r) { sy = 2 * yP + 1 - yS; sx = 2 * xP + xS; }       /* (G-247) */
                                else { sy = 2 * yP + 1 - yS; sx = 2 * xP + 1 - xS; }                                                             /* (G-248) */
                                for (k = 0; k < 2; ++k) mvILPredL0[2 * yP + yS][2 * xP + xS][k] = mvILPredL0[sy][sx][k];
                            }
                        }
            /* utils.c:1916-1979 (maxX = 0 in EP slices): (G-251)..(G-261) */
            for (yP = 0; yP < 2; ++yP)
                for (xP = 0; xP < 2; ++xP) {
                    const int xO = xP << 1, yO = yP << 1;
                    int (*a)[2] = &mvILPredL0[yO][xO], (*b)[2] = &mvILPredL0[yO][xO + 1], (*c)[2] = &mvILPredL0[yO + 1][xO], (*d)[2] = &mvILPredL0[yO + 1][xO + 1];
#define MVD(p, q) (abs((*p)[0] - (*q)[0]) + abs((*p)[1] - (*q)[1]))
                    if (MVD(a, b) <= 1 && MVD(a, c) <= 1 && MVD(a, d) <= 1) {
                        for (k = 0; k < 2; ++k) { const int v = ((*a)[k] + (*b)[k] + (*c)[k] + (*d)[k] + 2) >> 2; (*a)[k] = (*b)[k] = (*c)[k] = (*d)[k] = v; }
                    }
                    else if (MVD(a, b) <= 1 && MVD(c, d) <= 1) {
                        for (k = 0; k < 2; ++k) { const int v = ((*a)[k] + (*b)[k] + 1) >> 1, w = ((*c)[k] + (*d)[k] + 1) >> 1; (*a)[k] = (*b)[k] = v; (*c)[k] = (*d)[k] = w; }
                    }
                    else if (MVD(a, c) <= 1 && MVD(b, d) <= 1) {
                        for (k = 0; k < 2; ++k) { const int v = ((*a)[k] + (*c)[k] + 1) >> 1, w = ((*b)[k] + (*d)[k] + 1) >> 1; (*a)[k] = (*c)[k] = v; (*b)[k] = (*d)[k] = w; }
                    }
#undef MVD
                }
        }
    }
    /* G.8.6.1.3, utils.c:2006-2122 (EP: one list) */
    {
        int partitionSize = 3, c, yy, xx;
        const int (*r)[2] = refIdxILPredL0;
#define MV_EQ(Y, X, Y0, X0) (mvILPredL0[Y][X][0] == mvILPredL0[Y0][X0][0] && mvILPredL0[Y][X][1] == mvILPredL0[Y0][X0][1])
        c = r[0][0] == r[0][1] && r[0][0] == r[1][0] && r[0][0] == r[1][1];
        for (yy = 0; yy < 4 && c; ++yy) for (xx = 0; xx < 4; ++xx) if (!MV_EQ(yy, xx, 0, 0)) { c = 0; break; }
        if (c) partitionSize = 0;
        else {
            c = r[0][0] == r[0][1] && r[1][0] == r[1][1];
            for (yy = 0; yy < 2 && c; ++yy) for (xx = 0; xx < 4; ++xx) if (!MV_EQ(yy, xx, 0, 0)) { c = 0; break; }
            for (yy = 2; yy < 4 && c; ++yy) for (xx = 0; xx < 4; ++xx) if (!MV_EQ(yy, xx, 2, 0)) { c = 0; break; }
            if (c) partitionSize = 1;
            else {
                c = r[0][0] == r[1][0] && r[0][1] == r[1][1];
                for (yy = 0; yy < 4 && c; ++yy) for (xx = 0; xx < 2; ++xx) if (!MV_EQ(yy, xx, 0, 0)) { c = 0; break; }
                for (yy = 0; yy < 4 && c; ++yy) for (xx = 2; xx < 4; ++xx) if (!MV_EQ(yy, xx, 0, 2)) { c = 0; break; }
                if (c) partitionSize = 2;
            }
        }
        /* Table G-7 (EP) -> hl_codec_264_mb_set_mb_type, mb.c:117-126 */
        {
            static const int N[4] = { 1, 2, 2, 4 }, W[4] = { 16, 16, 8, 8 }, H[4] = { 16, 8, 16, 8 };
            int subSize[4] = { 3, 0, 0, 0 };   /* utils.c:2145: `subPartitionSize[4] = { 4X4 }` leaves elements 1..3 at enum value 0 = 8X8 */
            int p, s;
            out16[1] = N[partitionSize]; out16[2] = W[partitionSize]; out16[3] = H[partitionSize];
            if (partitionSize == 3) {   /* utils.c:2143-2186, Table G-8 -> hl_codec_264_mb_set_sub_mb_type, mb.c:175-186 */
                static const int SN[4] = { 1, 2, 2, 4 }, SW[4] = { 8, 8, 4, 4 }, SH[4] = { 8, 4, 8, 4 };
                for (p = 0; p < 4; ++p) {
                    const int xO = (p & 1) << 1, yO = (p >> 1) << 1;
                    if (MV_EQ(yO, xO + 1, yO, xO) && MV_EQ(yO + 1, xO, yO, xO) && MV_EQ(yO + 1, xO + 1, yO, xO)) subSize[p] = 0;
                    else if (MV_EQ(yO, xO + 1, yO, xO) && MV_EQ(yO + 1, xO + 1, yO + 1, xO)) subSize[p] = 1;
                    else if (MV_EQ(yO + 1, xO, yO, xO) && MV_EQ(yO + 1, xO + 1, yO, xO + 1)) subSize[p] = 2;
                    out16[4 + p] = SN[subSize[p]]; out16[8 + p] = SW[subSize[p]]; out16[12 + p] = SH[subSize[p]];
                }
            }
            else for (p = 0; p < N[partitionSize]; ++p) { out16[4 + p] = 1; out16[8 + p] = W[partitionSize]; out16[12 + p] = H[partitionSize]; }   /* mb.c:234-243 */
            /* G.8.4.1, utils.c:1533-1545 and :1606-1633 */
            for (p = 0; p < out16[1]; ++p) {
                /* 6.4.2.1: upper-left sample of the macroblock partition, raster order inside the 16x16 */
                const int xP = (p % (16 / out16[2])) * out16[2], yP = (p / (16 / out16[2])) * out16[3];
                for (s = 0; s < out16[4 + p]; ++s) {
                    int xS = 0, yS = 0;
                    if (partitionSize == 3) { xS = (s % (8 / out16[8 + p])) * out16[8 + p]; yS = (s / (8 / out16[8 + p])) * out16[12 + p]; }   /* 6.4.2.2, mb.h:283-300 */
                    if (s == 0) ref_idx4[p] = refIdxILPredL0[(yP + yS) >> 3][(xP + xS) >> 3];   /* (G-94) */
                    mv32[(p * 4 + s) * 2] = mvILPredL0[(yP + yS) >> 2][(xP + xS) >> 2][0];       /* (G-93), (G-96) */
                    mv32[(p * 4 + s) * 2 + 1] = mvILPredL0[(yP + yS) >> 2][(xP + xS) >> 2][1];
                }
            }
        }
#undef MV_EQ
    }
    return 0;
}
/* whole picture: out = nmb x (16 + 4 + 32) int32 {out16, refIdxL0[4], mvL0[4][4][2]}; bad[] receives 1 where hlo_svc_derive_mb failed */
HLO_API void hlo_svc_derive_picture(const int32_t* base, int ref_w, int ref_h, int scaled_w, int scaled_h, int off_x, int off_y, int level_idc, int restricted, int width, int height,
                                    int32_t* out, uint8_t* bad)
{
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    int a;
    for (a = 0; a < nmb; ++a) bad[a] = hlo_svc_derive_mb(base, ref_w, ref_h, scaled_w, scaled_h, off_x, off_y, level_idc, restricted, a % mbw, a / mbw, out + 52 * a, out + 52 * a + 16, out + 52 * a + 20) != 0;
}
