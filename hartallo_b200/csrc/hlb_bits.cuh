// hlb_bits.cuh -- slice_data() of one picture from the decision records and the per-macroblock state the slice kernel left: the per-macroblock part shared by
// the device kernels (hlb_slice.cu: k_bits_len / k_bits_write) and the CPU check of the drop-in (tools/emu/svc_shim.cpp).  See hlb_cavlc.cuh for the syntax.
#pragma once
#include "hlb_mbcore.cuh"
#include "hlb_cavlc.cuh"

namespace hlb {

struct BitsJob {
    const hlb200_mb_record_t* rec;
    const MbState* st;
    uint32_t* len;       // [nmb] bit lengths, turned into offsets by the scan
    uint32_t* out;       // word buffer
    uint32_t* hdr;       // [0] total bits, [1] != 0: the picture does not fit cap_words
    int nmb, mbw, is_p, cap_words;
};
HLB_HD CavlcNb bits_nb(const BitsJob& j, int mb, bool avail)
{
    CavlcNb nb;
    nb.avail = avail;
    if (avail) { const MbState& s = j.st[mb]; nb.kind = s.kind; nb.cbp_luma = s.cbp_luma; nb.cbp_chroma = s.cbp_chroma; nb.tc_luma = s.tc_luma; nb.tc_cac = &s.tc_cac[0][0]; }
    else { nb.kind = 0; nb.cbp_luma = nb.cbp_chroma = 0; nb.tc_luma = nullptr; nb.tc_cac = nullptr; }
    return nb;
}
// number of P_Skip macroblocks immediately before `mb` (mb_skip_run of a coded macroblock, mb.c:604-616)
HLB_HD uint32_t bits_skip_run_before(const BitsJob& j, int mb)
{
    uint32_t run = 0;
    while (mb - 1 - (int)run >= 0 && j.rec[mb - 1 - (int)run].mb_class == HLB200_MB_P_SKIP) ++run;
    return run;
}
template <class S> HLB_HD void bits_put_mb(S& s, const BitsJob& j, int mb)
{
    const hlb200_mb_record_t& r = j.rec[mb];
    if (j.is_p) {
        if (r.mb_class == HLB200_MB_P_SKIP) {
            if (mb == j.nmb - 1) put_ue(s, bits_skip_run_before(j, mb) + 1);   // the run that ends the slice (mb.c:590-594)
            return;
        }
        put_ue(s, bits_skip_run_before(j, mb));
    }
    const int x = mb % j.mbw;
    cavlc_put_mb(s, r, bits_nb(j, mb - 1, x > 0), bits_nb(j, mb - j.mbw, mb >= j.mbw));
}

}  // namespace hlb
