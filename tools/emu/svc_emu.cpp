// CPU run of the SVC enhancement-layer macroblock phases of hartallo_b200/csrc/hlb_svc.cuh (the body of k_svc_inter_recon): the lanes of a warp become a
// loop, the __syncwarp between the phases becomes the end of the first loop.  Same argument list as hlb200_dev_svc_inter_recon_batch, host pointers.
// A debugging aid and the CPU-tier check of that source against the reference's trace (tests/test_svc_inter.py); NOT linked into libhl_b200.so and
// not an oracle (the oracle is oracle/hl_oracle.c: hlo_recon_svc_inter_mb).
#include <stddef.h>
#include <stdint.h>
#include "../../hartallo_b200/csrc/hlb_svc.cuh"

static int host_chroma_qp(int qp_y, int offset) { int q = qp_y + offset; q = q < 0 ? 0 : (q > 51 ? 51 : q); return hlb::kQpc[q]; }

// bl != 0: I_BL macroblocks, ref_* are the prediction planes and motion is not read (hlb200_dev_svc_bl_recon_batch)
extern "C" __attribute__((visibility("default"))) int svc_emu_recon_batch(int bl,
    const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v, int width, int height, int n_pics,
    size_t frame_stride, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* motion, hlb200_svc_mb_state_t* state, hlb200_mb_coeffs_t* coeffs, uint8_t* rec_y,
    uint8_t* rec_u, uint8_t* rec_v)
{
    if ((width & 15) || (height & 15) || qp < 0 || qp > 51 || n_pics < 1) return HLB200_ERR_INVALID_PARAMETER;
    const int mbw = width >> 4, nmb = mbw * (height >> 4), qpc = host_chroma_qp(qp, chroma_qp_index_offset);
    for (int pic = 0; pic < n_pics; ++pic) {
        const size_t o = (size_t)pic * frame_stride;
        hlb::SvcPlanes P;
        P.src_y = src_y + o; P.src_u = src_u + o; P.src_v = src_v + o; P.ref_y = ref_y + o; P.ref_u = ref_u + o; P.ref_v = ref_v + o;
        P.rec_y = rec_y + o; P.rec_u = rec_u + o; P.rec_v = rec_v + o; P.W = width; P.H = height;
        for (int mb = 0; mb < nmb; ++mb) {
            const size_t idx = (size_t)pic * nmb + mb;
            const int mbx = mb % mbw, mby = mb / mbw;
            hlb::SvcXchg X;
            hlb::SvcLane L[24];
            hlb::SvcPredSrc ps;
            if (bl) { ps.m = nullptr; ps.mbx = mbx; ps.mby = mby; ps.inherited = false; }
            else ps = hlb::svc_pred_src(motion + (size_t)pic * nmb, mb, mbw);
            for (int lane = 0; lane < 24; ++lane) {
                if (bl) hlb::svc_lane_a<true>(P, mbx, mby, lane, ps, qp, qpc, state[idx], L[lane], X);
                else hlb::svc_lane_a<false>(P, mbx, mby, lane, ps, qp, qpc, state[idx], L[lane], X);
            }
            // all phase-A reads of the state happen before any phase-B write, as on the device (phase B rewrites ChromaDCLevel with the values its
            // other lanes read, or with new ones that nobody reads)
            for (int lane = 0; lane < 24; ++lane) hlb::svc_lane_b(P, mbx, mby, lane, qp, qpc, bl || ps.inherited, state[idx], L[lane], X, coeffs[idx]);
            coeffs[idx].cbp_luma4x4 = (uint16_t)hlb::svc_luma_cbp(X);
        }
    }
    return HLB200_OK;
}

// CPU run of svc_resample_px (the body of k_svc_resample_intra): one plane, host pointers
extern "C" __attribute__((visibility("default"))) int svc_emu_resample_plane(const uint8_t* ref, int refW, int refH, uint8_t* out, int W, int H, int chroma, int level_idc)
{
    const hlb::SvcRsAxis ax = hlb::svc_rs_axis(refW, W, level_idc), ay = hlb::svc_rs_axis(refH, H, level_idc);
    // the form the kernel's threads run: four samples per call (svc_resample_row4), which itself falls back to the per-sample form at the picture edges
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; x += 4) {
            const uint32_t w = hlb::svc_resample_row4(ref, refW, refH, ax, ay, x, y, chroma != 0);
            for (int i = 0; i < 4; ++i) out[y * W + x + i] = (uint8_t)(w >> (8 * i));
        }
    return HLB200_OK;
}

// CPU run of the inter-layer motion derivation (hlb_svc_derive.cuh, the body of k_svc_derive / k_svc_derive_inherit): one picture, host pointers.  Same meaning of the
// arguments as hlb200_dev_svc_derive_motion_batch; returns the HLB200_SVC_DERIVE_* status bits in *status.
#include "../../hartallo_b200/csrc/hlb_svc_derive.cuh"
extern "C" __attribute__((visibility("default"))) int svc_emu_derive_motion(const hlb200_svc_base_mb_t* base, const hlb200_svc_layer_geom_t* geom, int width, int height,
                                                                            uint8_t* had_parts, hlb200_mb_motion_t* motion, int32_t* status)
{
    if (!base || !geom || !had_parts || !motion || !status || width < 16 || height < 16 || (width & 15) || (height & 15)) return HLB200_ERR_INVALID_PARAMETER;
    if (geom->cropping_change) return HLB200_ERR_NOT_IMPLEMENTED;
    hlb::SvcDeriveGeom g;
    if (!hlb::svc_derive_geom(geom->ref_width, geom->ref_height, geom->scaled_width, geom->scaled_height, geom->left_offset, geom->top_offset, geom->level_idc, geom->restricted, g)) return HLB200_ERR_NOT_IMPLEMENTED;
    const int mbw = width >> 4, nmb = mbw * (height >> 4);
    int st = 0;
    for (int mb = 0; mb < nmb; ++mb) st |= hlb::svc_derive_pass1(base, g, mb, mbw, had_parts, motion);
    for (int mb = 0; mb < nmb; ++mb) st |= hlb::svc_derive_pass2(mb, had_parts, motion);
    *status = st;
    return HLB200_OK;
}
