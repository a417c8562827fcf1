// hlb_deblock.cuh -- in-loop deblocking of a finished picture (8.7), as the reference's baseline path runs it after the last macroblock of a picture
// (hl_codec_264_deblock_avc deblock.c:192, called slice.c:1897; per macroblock hl_codec_264_deblock_avc_mb deblock.c:573 ->
// ..._baseline_mb_luma_u8 :2701 / ..._baseline_mb_chroma_u8 :3229).  What is reproduced:
//   flags       disable_deblocking_filter_idc = 0, FilterOffsetA = FilterOffsetB = 0 (slice.c:296-297, :707): the left / top macroblock edge is filtered
//               unless it is the picture edge (deblock.c:645-646).  filterInternalEdgesFlag = 0 for P_L0_16x16 / P_Skip without luma residual
//               (deblock.c:637) only skips edges whose bS is 0 anyway (one vector, no coefficients).
//   bS          hl_codec_264_deblock_avc_baseline_get_bs_luma4lines deblock.c:1784-1834, one value per 4 lines: 4 / 3 intra (macroblock edge / inside),
//               2 a 4x4 block with coefficients (CodedBlockPatternLuma4x4), 1 different reference or a vector component 4 or more quarter samples
//               apart, else 0.  Chroma edges take the bS of luma edges 0 and 2, two lines per value (deblock.c:3296-3303, :3365-3372).
//   thresholds  indexA / alpha / beta deblock.c:1836-1844 (tables :51-58), filterSamplesFlag :1847-1881
//   filters     bS < 4 deblock.c:1883-1928 (tc0 table :63-72), bS = 4 :2338-2375; an 8-line group is filtered as "bS < 4" or "bS = 4" by the bS of its
//               first four lines (deblock.c:2772, :2787) -- the same thing, bS = 4 holds for all lines of a macroblock edge or none.
// Order (8.7): macroblocks in raster order; per macroblock vertical edges left to right, then horizontal edges top to bottom; luma and the two chroma
// planes are independent.  A vertical edge's 16 lines (chroma: 8) are independent of each other, so the lanes of a warp take one line each: lanes 0..15
// luma, 16..23 Cb, 24..31 Cr (chroma only has edges 0 and 2).
//
// Portable like hlb_mbcore.cuh: the emulation harness runs dbk_bs_mb / dbk_edge with lanes as loops.
#pragma once
#include "hlb_mbcore.cuh"
#include "../../include/hlb200.h"

namespace hlb {

struct DbkJob {
    uint8_t* plane[3];                 // reconstruction of the picture (pitch W / W/2), filtered in place
    const hlb200_mb_record_t* rec;     // the picture's decision records
    uint8_t* bs;                       // 32 bytes per macroblock: [dir 0 vertical / 1 horizontal][edge 0..3][group of 4 lines]
    int W, H, mbw, mbh;
    int enabled;
    int16_t alpha[2], beta[2];         // [0] luma (QPY), [1] chroma (QPC; Cb and Cr share chroma_qp_index_offset in the reference's PPS, pps.c:292)
    uint8_t tc0[2][4];                 // t'c0 for bS 1..3 at indexA of luma / chroma (index 0 unused)
};

// host: thresholds of a picture whose macroblocks all carry the same QP (no rate control on the device path): qPav = QP (8-461)
inline void dbk_job_thresholds(DbkJob& j, int qp, int qpc)
{
    static const uint8_t A[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13,
                                  15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
    static const uint8_t B[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                  6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
    static const uint8_t T[52][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
                                     {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1}, {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1},
                                     {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3}, {2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6},
                                     {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10}, {6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23},
                                     {13, 17, 25}};
    const int q[2] = {qp < 0 ? 0 : (qp > 51 ? 51 : qp), qpc < 0 ? 0 : (qpc > 51 ? 51 : qpc)};
    for (int k = 0; k < 2; ++k) {
        j.alpha[k] = A[q[k]]; j.beta[k] = B[q[k]];
        j.tc0[k][0] = 0;
        for (int b = 1; b <= 3; ++b) j.tc0[k][b] = T[q[k]][b - 1];
    }
}

// luma4x4BlkIdx of the 4x4 block at block coordinates (bx, by) (6.4.3 inverse; LumaBlockIndices4x4_YX of hl_codec_264_tables.h:215)
HLB_HD int dbk_blkidx(int bx, int by) { return ((by >> 1) << 3) | ((bx >> 1) << 2) | ((by & 1) << 1) | (bx & 1); }

// bS between the 4x4 block at block coordinates (pbx, pby) of macroblock P and (qbx, qby) of macroblock Q (deblock.c:1784-1834)
HLB_HD int dbk_bs(const hlb200_mb_record_t* P, const hlb200_mb_record_t* Q, int pbx, int pby, int qbx, int qby, int mb_edge)
{
    const int kp = P->mb_class, kq = Q->mb_class;
    if (kp >= HLB200_MB_I16x16 || kq >= HLB200_MB_I16x16) return mb_edge ? 4 : 3;
    if (((P->cbp_luma4x4 >> dbk_blkidx(pbx, pby)) & 1) || ((Q->cbp_luma4x4 >> dbk_blkidx(qbx, qby)) & 1)) return 2;
    int pp, ps, qp, qs;
    uint8_t subp[4], subq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { subp[i] = P->sub_mode[i]; subq[i] = Q->sub_mode[i]; }
    part_at(P->part_mode, subp, pbx * 4, pby * 4, pp, ps);
    part_at(Q->part_mode, subq, qbx * 4, qby * 4, qp, qs);
    if (P->ref_idx[pp] != Q->ref_idx[qp]) return 1;
    const int dx = (int)P->mv[pp][ps][0] - (int)Q->mv[qp][qs][0], dy = (int)P->mv[pp][ps][1] - (int)Q->mv[qp][qs][1];
    return (iabs(dx) >= 4 || iabs(dy) >= 4) ? 1 : 0;
}

// bS value `idx` (= dir * 16 + edge * 4 + group) of macroblock (mbx, mby)
HLB_HD int dbk_bs_mb(const DbkJob& j, int mbx, int mby, int idx)
{
    const int dir = idx >> 4, e = (idx >> 2) & 3, g = idx & 3;
    const hlb200_mb_record_t* Q = j.rec + mby * j.mbw + mbx;
    if (e == 0) {
        if (dir == 0) return mbx ? dbk_bs(Q - 1, Q, 3, g, 0, g, 1) : 0;
        return mby ? dbk_bs(Q - j.mbw, Q, g, 3, g, 0, 1) : 0;
    }
    return dir == 0 ? dbk_bs(Q, Q, e - 1, g, e, g, 0) : dbk_bs(Q, Q, g, e - 1, g, e, 0);
}

// One line of one edge: s points at q0, `step` is the distance between neighbouring samples across the edge (1: vertical edge, pitch: horizontal edge).
// 8.7.2.2 - 8.7.2.4 as deblock.c:1847-1928, :2338-2375.
HLB_HD void dbk_filter_line(uint8_t* s, int step, int bS, int alpha, int beta, int tc0, int chroma)
{
    const int p0 = s[-step], p1 = s[-2 * step], q0 = s[0], q1 = s[step];
    if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;   // filterSamplesFlag (8-468)
    if (chroma) {
        if (bS < 4) {
            const int tc = tc0 + 1, d0 = (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, d = d0 < -tc ? -tc : (d0 > tc ? tc : d0);
            s[-step] = (uint8_t)clip255(p0 + d); s[0] = (uint8_t)clip255(q0 - d);
        } else {
            s[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); s[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
        return;
    }
    const int p2 = s[-3 * step], q2 = s[2 * step], ap = iabs(p2 - p0), aq = iabs(q2 - q0);
    if (bS < 4) {
        const int tc = tc0 + (ap < beta) + (aq < beta), d0 = (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, d = d0 < -tc ? -tc : (d0 > tc ? tc : d0);
        s[-step] = (uint8_t)clip255(p0 + d); s[0] = (uint8_t)clip255(q0 - d);
        if (ap < beta) { const int v = (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1; s[-2 * step] = (uint8_t)(p1 + (v < -tc0 ? -tc0 : (v > tc0 ? tc0 : v))); }
        if (aq < beta) { const int v = (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1; s[step] = (uint8_t)(q1 + (v < -tc0 ? -tc0 : (v > tc0 ? tc0 : v))); }
    } else {
        const int small = iabs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && small) {
            const int p3 = s[-4 * step];
            s[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            s[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            s[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else s[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (aq < beta && small) {
            const int q3 = s[3 * step];
            s[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            s[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            s[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else s[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

// Work of lane `lane` (0..31) for edge `e` (0..3) in direction `dir` of macroblock (mbx, mby); bs32 = the macroblock's 32 bS bytes.
// Lanes of one call are independent; calls must follow each other in (dir, e) order with the earlier call's stores visible.
HLB_HD void dbk_edge(const DbkJob& j, int mbx, int mby, int dir, int e, int lane, const uint8_t* bs32)
{
    if (lane < 16) {
        const int bS = bs32[dir * 16 + e * 4 + (lane >> 2)];
        if (!bS) return;
        uint8_t* s = j.plane[0] + (size_t)(mby * 16 + (dir ? e * 4 : lane)) * j.W + mbx * 16 + (dir ? lane : e * 4);
        dbk_filter_line(s, dir ? j.W : 1, bS, j.alpha[0], j.beta[0], j.tc0[0][bS < 4 ? bS : 0], 0);
    } else {
        if (e & 1) return;
        const int l = lane & 7, bS = bs32[dir * 16 + e * 4 + (l >> 1)], Wc = j.W >> 1;
        if (!bS) return;
        uint8_t* s = j.plane[1 + ((lane >> 3) & 1)] + (size_t)(mby * 8 + (dir ? e * 2 : l)) * Wc + mbx * 8 + (dir ? l : e * 2);
        dbk_filter_line(s, dir ? Wc : 1, bS, j.alpha[1], j.beta[1], j.tc0[1][bS < 4 ? bS : 0], 1);
    }
}

#if !defined(__CUDACC__)
// the emulation harness: the whole picture, lanes as loops
inline void dbk_picture_serial(const DbkJob& j)
{
    for (int mby = 0; mby < j.mbh; ++mby)
        for (int mbx = 0; mbx < j.mbw; ++mbx) {
            uint8_t bs32[32];
            for (int i = 0; i < 32; ++i) bs32[i] = (uint8_t)dbk_bs_mb(j, mbx, mby, i);
            for (int dir = 0; dir < 2; ++dir)
                for (int e = 0; e < 4; ++e)
                    for (int lane = 0; lane < 32; ++lane) dbk_edge(j, mbx, mby, dir, e, lane, bs32);
        }
}
#endif

}  // namespace hlb
