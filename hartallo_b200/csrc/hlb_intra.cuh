// hlb_intra.cuh -- intra prediction sample generators (8.3.1.2, 8.3.3, 8.3.4), bit-exact with the reference:
//   source/h264/hl_codec_264_pred_intra.c:618-853 (4x4), :856-1041 (16x16), :1044-1230 (chroma)
// Neighbour arrays use the reference's index conventions:
//   p13: [0] = p[-1,-1], [1..4] = p[-1,0..3], [5..12] = p[0..7,-1]
//   p33: [0] = p[-1,-1], [1..16] = p[-1,0..15], [17..32] = p[0..15,-1]
//   p17 (chroma): [0] = p[-1,-1], [1..8] = p[-1,0..7], [9..16] = p[0..7,-1]
// HLB_NA marks an unavailable sample (HL_CODEC_264_SAMPLE_NOT_AVAIL).
#pragma once
#include "hlb_prims.cuh"

namespace hlb {

#define HLB_NA ((int)0xFFFF0000)

// p13 accessor: P(x,-1) for x=-1..7, P(-1,y) for y=-1..3
HLB_HD int p13_at(const int* p, int x, int y) { return y < 0 ? (x < 0 ? p[0] : p[5 + x]) : p[1 + y]; }

// which modes the reference tries (rdo.c:1907-1932)
HLB_HD bool i4_mode_allowed(int mode, const int* p)
{
    switch (mode) {
    case 0: case 3: case 7: return p[5] != HLB_NA;        // Vertical, Diagonal_Down_Left, Vertical_Left
    case 1: case 8: return p[1] != HLB_NA;                // Horizontal, Horizontal_Up
    case 4: case 5: case 6: return p[0] != HLB_NA;        // Diagonal_Down_Right, Vertical_Right, Horizontal_Down
    default: return true;                                 // DC
    }
}

HLB_HD void intra4x4_pred(int mode, const int* p, int out[16])
{
    switch (mode) {
    case 0:  // Vertical
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) out[y * 4 + x] = p[5 + x];
        break;
    case 1:  // Horizontal
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) out[y * 4 + x] = p[1 + y];
        break;
    case 2: {  // DC
        const bool xa = p[5] != HLB_NA && p[6] != HLB_NA && p[7] != HLB_NA && p[8] != HLB_NA;
        const bool ya = p[1] != HLB_NA && p[2] != HLB_NA && p[3] != HLB_NA && p[4] != HLB_NA;
        int r;
        if (xa && ya) r = (p[5] + p[6] + p[7] + p[8] + p[1] + p[2] + p[3] + p[4] + 4) >> 3;
        else if (ya) r = (p[1] + p[2] + p[3] + p[4] + 2) >> 2;
        else if (xa) r = (p[5] + p[6] + p[7] + p[8] + 2) >> 2;
        else r = 128;
        for (int i = 0; i < 16; ++i) out[i] = r;
        break;
    }
    case 3:  // Diagonal_Down_Left
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x)
            out[y * 4 + x] = (x == 3 && y == 3) ? (p[11] + 3 * p[12] + 2) >> 2 : (p[5 + x + y] + 2 * p[5 + x + y + 1] + p[5 + x + y + 2] + 2) >> 2;
        break;
    case 4:  // Diagonal_Down_Right
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
            int v;
            if (x > y) v = (p13_at(p, x - y - 2, -1) + 2 * p13_at(p, x - y - 1, -1) + p13_at(p, x - y, -1) + 2) >> 2;
            else if (x < y) v = (p13_at(p, -1, y - x - 2) + 2 * p13_at(p, -1, y - x - 1) + p13_at(p, -1, y - x) + 2) >> 2;
            else v = (p[5] + 2 * p[0] + p[1] + 2) >> 2;
            out[y * 4 + x] = v;
        }
        break;
    case 5:  // Vertical_Right
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
            const int z = 2 * x - y;
            int v;
            if (z >= 0 && (z & 1) == 0) v = (p13_at(p, x - (y >> 1) - 1, -1) + p13_at(p, x - (y >> 1), -1) + 1) >> 1;
            else if (z >= 0) v = (p13_at(p, x - (y >> 1) - 2, -1) + 2 * p13_at(p, x - (y >> 1) - 1, -1) + p13_at(p, x - (y >> 1), -1) + 2) >> 2;
            else if (z == -1) v = (p[1] + 2 * p[0] + p[5] + 2) >> 2;
            else v = (p13_at(p, -1, y - 1) + 2 * p13_at(p, -1, y - 2) + p13_at(p, -1, y - 3) + 2) >> 2;
            out[y * 4 + x] = v;
        }
        break;
    case 6:  // Horizontal_Down
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
            const int z = 2 * y - x;
            int v;
            if (z >= 0 && (z & 1) == 0) v = (p13_at(p, -1, y - (x >> 1) - 1) + p13_at(p, -1, y - (x >> 1)) + 1) >> 1;
            else if (z >= 0) v = (p13_at(p, -1, y - (x >> 1) - 2) + 2 * p13_at(p, -1, y - (x >> 1) - 1) + p13_at(p, -1, y - (x >> 1)) + 2) >> 2;
            else if (z == -1) v = (p[1] + 2 * p[0] + p[5] + 2) >> 2;
            else v = (p13_at(p, x - 1, -1) + 2 * p13_at(p, x - 2, -1) + p13_at(p, x - 3, -1) + 2) >> 2;
            out[y * 4 + x] = v;
        }
        break;
    case 7:  // Vertical_Left
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
            const int k = x + (y >> 1);
            out[y * 4 + x] = (y & 1) ? (p[5 + k] + 2 * p[5 + k + 1] + p[5 + k + 2] + 2) >> 2 : (p[5 + k] + p[5 + k + 1] + 1) >> 1;
        }
        break;
    default:  // Horizontal_Up
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
            const int z = x + 2 * y;
            int v;
            if (z > 5) v = p[4];
            else if (z == 5) v = (p[3] + 3 * p[4] + 2) >> 2;
            else if ((z & 1) == 0) v = (p[1 + y + (x >> 1)] + p[1 + y + (x >> 1) + 1] + 1) >> 1;
            else v = (p[1 + y + (x >> 1)] + 2 * p[1 + y + (x >> 1) + 1] + p[1 + y + (x >> 1) + 2] + 2) >> 2;
            out[y * 4 + x] = v;
        }
        break;
    }
}

// which Intra16x16 modes the reference tries (rdo.c:1611-1631): 0 V needs p[17], 1 H needs p[1], 3 Plane needs p[0]
HLB_HD bool i16_mode_allowed(int mode, const int* p33) { return mode == 0 ? p33[17] != HLB_NA : (mode == 1 ? p33[1] != HLB_NA : (mode == 3 ? p33[0] != HLB_NA : true)); }

// one sample of the Intra16x16 prediction (mode 0 V, 1 H, 2 DC, 3 Plane); dc/a/b/c precomputed by i16_params
struct I16Params { int dc, a, b, c; };
HLB_HD I16Params i16_params(const int* p)
{
    I16Params r;
    bool xa = true, ya = true;
    int xs = 0, ys = 0;
    for (int i = 0; i < 16; ++i) { if (p[17 + i] == HLB_NA) { xa = false; break; } xs += p[17 + i]; }
    for (int i = 0; i < 16; ++i) { if (p[1 + i] == HLB_NA) { ya = false; break; } ys += p[1 + i]; }
    if (xa && ya) r.dc = (xs + ys + 16) >> 5;
    else if (ya) r.dc = (ys + 8) >> 4;
    else if (xa) r.dc = (xs + 8) >> 4;
    else r.dc = 128;
    int H = 0, V = 0;
    for (int k = 1; k <= 7; ++k) { H += k * (p[24 + k] - p[24 - k]); V += k * (p[8 + k] - p[8 - k]); }
    H += 8 * (p[32] - p[0]);
    V += 8 * (p[16] - p[0]);
    r.a = (p[16] + p[32]) << 4;
    r.b = (5 * H + 32) >> 6;
    r.c = (5 * V + 32) >> 6;
    return r;
}
HLB_HD int i16_pred_px(int mode, const int* p, const I16Params& q, int x, int y)
{
    switch (mode) {
    case 0: return p[17 + x];
    case 1: return p[1 + y];
    case 2: return q.dc;
    default: return clip255((q.a + q.b * (x - 7) + q.c * (y - 7) + 16) >> 5);
    }
}

// chroma (8x8) prediction sample; mode 0 DC, 1 Horizontal, 2 Vertical, 3 Plane (Intra_Chroma_* enum order)
struct ICParams { int dc[4], a, b, c; };
HLB_HD ICParams ic_params(const int* p)
{
    ICParams r;
    bool xa = true, ya = true;  // NOTE: sticky across the four blocks, as in the reference (pred_intra.c:1046)
    for (int blk = 0; blk < 4; ++blk) {
        const int xO = (blk & 1) * 4, yO = (blk >> 1) * 4;
        int xs = 0, ys = 0;
        for (int i = 0; i < 4; ++i) { if (p[9 + xO + i] == HLB_NA) { xa = false; break; } xs += p[9 + xO + i]; }
        for (int i = 0; i < 4; ++i) { if (p[1 + yO + i] == HLB_NA) { ya = false; break; } ys += p[1 + yO + i]; }
        int t;
        if ((xO == 0 && yO == 0) || (xO > 0 && yO > 0)) {
            if (xa && ya) t = (xs + ys + 4) >> 3;
            else if (ya) t = (ys + 2) >> 2;
            else if (xa) t = (xs + 2) >> 2;
            else t = 128;
        } else if (xO > 0) {
            if (xa) t = (xs + 2) >> 2;
            else if (ya) t = (ys + 2) >> 2;
            else t = 128;
        } else {
            if (ya) t = (ys + 2) >> 2;
            else if (xa) t = (xs + 2) >> 2;
            else t = 128;
        }
        r.dc[blk] = t;
    }
    int H = 0, V = 0;
    for (int k = 1; k <= 3; ++k) { H += k * (p[12 + k] - p[12 - k]); V += k * (p[4 + k] - p[4 - k]); }
    H += 4 * (p[16] - p[0]);
    V += 4 * (p[8] - p[0]);
    r.a = (p[8] + p[16]) << 4;
    r.b = (34 * H + 32) >> 6;
    r.c = (34 * V + 32) >> 6;
    return r;
}
HLB_HD int ic_pred_px(int mode, const int* p, const ICParams& q, int x, int y)
{
    switch (mode) {
    case 0: return q.dc[((y >> 2) << 1) | (x >> 2)];
    case 1: return p[1 + y];
    case 2: return p[9 + x];
    default: {
        // the reference iterates with UNSIGNED x,y (pred_intra.c:1194,1214): the sum is formed modulo 2^32 and shifted LOGICALLY, so a
        // negative plane value becomes a large positive one and clips to 255 instead of 0
        const uint32_t v = ((uint32_t)q.a + (uint32_t)q.b * (uint32_t)(x - 3) + (uint32_t)q.c * (uint32_t)(y - 3) + 16u) >> 5;
        return clip255((int)v);
    }
    }
}

}  // namespace hlb
