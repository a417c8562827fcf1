// cross-check of the register-only CAVLC length function against the array formulation (hlb_prims.cuh) on random level blocks
#include <stdio.h>
#include <stdlib.h>
#include "../../hartallo_b200/csrc/hlb_prims.cuh"
int main()
{
    unsigned s = 7;
    long bad = 0, n = 0;
    for (int it = 0; it < 3000000; ++it) {
        int lv[16];
        s = s * 1664525u + 1013904223u;
        const int dens = (s >> 8) % 17, big = (s >> 16) % 4;
        for (int i = 0; i < 16; ++i) {
            s = s * 1664525u + 1013904223u;
            int v = 0;
            if ((int)((s >> 8) % 16) < dens) { const int m = big == 0 ? 2 : (big == 1 ? 5 : (big == 2 ? 40 : 3000)); v = (int)((s >> 12) % (2 * m + 1)) - m; }
            lv[i] = v;
        }
        const hlb::CavlcInfo a = hlb::cavlc_block_info(lv, 16, false), b = hlb::cavlc_block_info_ref(lv, 16, false);
        ++n;
        if (a.total_coeff != b.total_coeff || (a.total_coeff && (a.trailing_ones != b.trailing_ones || a.bits_rest != b.bits_rest || a.single_ctr != b.single_ctr))) {
            if (bad++ < 5) { printf("mismatch tc %d/%d t1 %d/%d bits %d/%d sc %d/%d :", a.total_coeff, b.total_coeff, a.trailing_ones, b.trailing_ones, a.bits_rest, b.bits_rest, a.single_ctr, b.single_ctr); for (int i = 0; i < 16; ++i) printf(" %d", lv[i]); printf("\n"); }
        }
    }
    // the slice kernel's trial memo packs bits_rest into 10 bits (hlb_mbcore.cuh me_phase_trial): bound it over the densest / largest blocks
    int maxbits = 0;
    for (int big = 1; big <= 4000; big = big * 3 + 1)
        for (int pat = 0; pat < 65536; pat += 257) {
            int lv[16];
            for (int i = 0; i < 16; ++i) lv[i] = ((pat >> i) & 1) ? ((i & 1) ? -big : big) : ((pat >> ((i + 5) & 15)) & 1 ? 1 : 0);
            const hlb::CavlcInfo a = hlb::cavlc_block_info(lv, 16, false);
            if (a.bits_rest > maxbits) maxbits = a.bits_rest;
        }
    {
        int lv[16];
        for (int i = 0; i < 16; ++i) lv[i] = (i & 1) ? -32000 : 32000;
        const hlb::CavlcInfo a = hlb::cavlc_block_info(lv, 16, false);
        if (a.bits_rest > maxbits) maxbits = a.bits_rest;
    }
    printf("%ld comparisons, %ld mismatches; largest bits_rest %d (memo field holds 1023)\n", n, bad, maxbits);
    return bad != 0 || maxbits > 1023;
}
