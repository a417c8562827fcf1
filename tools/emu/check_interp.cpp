// cross-check of the two formulations of the luma interpolation (hlb_prims.cuh) on random tiles, all 16 positions
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../hartallo_b200/csrc/hlb_prims.cuh"
int main()
{
    uint8_t tile[20 * 24];
    unsigned s = 1;
    long bad = 0, n = 0;
    for (int it = 0; it < 20000; ++it) {
        for (int i = 0; i < (int)sizeof(tile); ++i) { s = s * 1664525u + 1013904223u; int r = (s >> 8) & 1023; tile[i] = r < 200 ? 0 : (r < 400 ? 255 : (r < 700 ? (s >> 20) % 34 : (s >> 16) & 255)); }
        for (int yf = 0; yf < 4; ++yf) for (int xf = 0; xf < 4; ++xf) {
            uint8_t a[16], b[16];
            hlb::interp_luma_4x4(tile + 5 * 24 + 6, 24, xf, yf, a);
            hlb::interp_luma_4x4_unrolled(tile + 5 * 24 + 6, 24, xf, yf, b);
            ++n;
            if (memcmp(a, b, 16)) { if (bad++ < 5) printf("mismatch xf %d yf %d\n", xf, yf); }
        }
    }
    printf("%ld comparisons, %ld mismatches\n", n, bad);
    return bad != 0;
}
