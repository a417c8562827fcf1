// check_fast.cpp -- CPU check of the packed formulations (hlb_fast.cuh) against the plain ones of hlb_prims.cuh (which the oracle pins
// against the reference): luma prediction at all 16 fractional positions and every word alignment, and the trial encode at every QP.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../hartallo_b200/csrc/hlb_fast.cuh"
using namespace hlb;

static uint32_t rng_s = 12345;
static uint32_t rnd() { rng_s = rng_s * 1664525u + 1013904223u; return rng_s >> 8; }

// the formulation of me_phase_trial before hlb_fast.cuh existed
static uint32_t ref_trial(const uint8_t sv[16], const uint8_t pv[16], int qp)
{
    int m[16], lv[16];
    uint32_t any = 0, mask = 0;
    for (int i = 0; i < 16; ++i) { m[i] = (int)sv[i] - (int)pv[i]; any |= (uint32_t)m[i]; }
    if (any) { fwd_transform4x4(m); quant4x4_ac(m, qp, false); zigzag4x4(m, lv); mask = level_mask16(lv); }
    if (mask == 0) return (uint32_t)sad16(sv, pv);
    const CavlcInfo ci = cavlc_block_info16(lv, mask);
    int cc[16];
    inv_zigzag4x4(lv, cc); dequant4x4(cc, qp, false); inv_transform4x4(cc);
    uint8_t rec[16];
    for (int i = 0; i < 16; ++i) rec[i] = (uint8_t)((int)pv[i] + cc[i]);
    return (uint32_t)sad16(sv, rec) | ((uint32_t)ci.bits_rest << 12) | ((uint32_t)ci.total_coeff << 22) | ((uint32_t)ci.trailing_ones << 27) | ((uint32_t)(ci.single_ctr & 3) << 29);
}

int main()
{
    long bad = 0, n = 0;
    // ---- prediction ----
    alignas(16) static uint8_t tile[48 * 48 + 64];
    for (int it = 0; it < 4000; ++it) {
        const int kind = it % 4;
        for (int i = 0; i < 48 * 48 + 64; ++i) {
            const uint32_t r = rnd();
            tile[i] = kind == 0 ? (uint8_t)r : kind == 1 ? (uint8_t)((r & 1) ? r % 34 : r) : kind == 2 ? (uint8_t)((r & 256) ? 255 : 0) : (uint8_t)(((r >> 9) & 7) == 0 ? ((r & 1) ? 255 : 0) : (r & 255));
        }
        const int tx = 2 + (int)(rnd() % (48 - 8)), ty = 2 + (int)(rnd() % (48 - 8));
        for (int yf = 0; yf < 4; ++yf)
            for (int xf = 0; xf < 4; ++xf) {
                uint8_t a[16];
                interp_luma_4x4_unrolled(tile + ty * 48 + tx, 48, xf, yf, a);
                const Rows4 b = fast_pred_luma((const uint32_t*)tile, 12, tx, ty, xf, yf);
                ++n;
                for (int i = 0; i < 16; ++i)
                    if (a[i] != (uint8_t)(b.r[i >> 2] >> (8 * (i & 3)))) { if (bad < 10) printf("pred mismatch tx %d ty %d xf %d yf %d px %d: %d vs %d\n", tx, ty, xf, yf, i, a[i], (b.r[i >> 2] >> (8 * (i & 3))) & 255); ++bad; break; }
                const Rows4 c = fast_pred_luma_staged<LdPlain>((const uint32_t*)tile, 12, tx, ty, xf, yf);   // the whole-picture kernel's formulation
                ++n;
                for (int r = 0; r < 4; ++r)
                    if (c.r[r] != b.r[r]) { if (bad < 10) printf("staged pred mismatch tx %d ty %d xf %d yf %d row %d: %08x vs %08x\n", tx, ty, xf, yf, r, c.r[r], b.r[r]); ++bad; break; }
            }
    }
    // ---- trial encode ----
    long nz = 0, shortcut = 0;
    for (int qp = 12; qp <= 51; ++qp) {
        QuantK q;
        quantk_make(q, qp);
        for (int it = 0; it < 6000; ++it) {
            uint8_t sv[16], pv[16];
            const int kind = it % 6, amp = 1 + (int)(rnd() % 64);
            for (int i = 0; i < 16; ++i) {
                const uint32_t r = rnd(), r2 = rnd();
                pv[i] = (uint8_t)r;
                if (kind == 0) sv[i] = (uint8_t)r2;
                else if (kind == 1) sv[i] = (uint8_t)clip255((int)pv[i] + (int)(r2 % (2 * amp + 1)) - amp);
                else if (kind == 2) { pv[i] = (r & 256) ? 255 : 0; sv[i] = (r2 & 256) ? 255 : (uint8_t)(r2 & 3); }
                else if (kind == 3) sv[i] = (uint8_t)clip255((int)pv[i] + ((r2 & 7) == 0 ? 1 : 0));
                else if (kind == 4) sv[i] = (uint8_t)clip255((int)pv[i] + (int)(r2 % 5) - 2);
                else sv[i] = (uint8_t)clip255((int)pv[i] + (i == (int)(amp & 15) ? (int)(r2 % 200) - 100 : 0));
            }
            Rows4 s, p;
            for (int r = 0; r < 4; ++r) { s.r[r] = sv[4 * r] | (sv[4 * r + 1] << 8) | (sv[4 * r + 2] << 16) | ((uint32_t)sv[4 * r + 3] << 24); p.r[r] = pv[4 * r] | (pv[4 * r + 1] << 8) | (pv[4 * r + 2] << 16) | ((uint32_t)pv[4 * r + 3] << 24); }
            const uint32_t a = ref_trial(sv, pv, qp), b = fast_trial(s, p, q, false), c = fast_trial(s, p, q, true);
            ++n;
            if ((a >> 22) & 31) ++nz;
            if ((int)sad16(sv, pv) <= q.zero_sad) ++shortcut;
            if (a != b) { if (bad < 10) printf("trial mismatch qp %d kind %d: %08x vs %08x\n", qp, kind, a, b); ++bad; }
            // counts-only result: TotalCoeff, and for a lone +-1 TrailingOnes / Single_ctr
            const int tc = (a >> 22) & 31, t1 = (a >> 27) & 3;
            bool ok = (int)((c >> 22) & 31) == tc;
            if (tc == 1) ok = ok && (int)((c >> 27) & 3) == t1 && (t1 != 1 || ((c >> 29) & 3) == ((a >> 29) & 3));
            if (!ok) { if (bad < 10) printf("counts-only mismatch qp %d kind %d: %08x vs %08x\n", qp, kind, a, c); ++bad; }
        }
    }
    printf("check_fast: %ld cases (%ld non-zero trial blocks, %ld zero-SAD shortcuts), %ld mismatches\n", n, nz, shortcut, bad);
    return bad ? 1 : 0;
}
