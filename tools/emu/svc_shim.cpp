// CPU stand-in for the few libhl_b200.so entry points the SVC part of host/hlb200_glue.c calls, backed by the CPU run of the device source
// (svc_emu.cpp).  TEST INFRASTRUCTURE: lets oracle/_ref/hl_svc_glue_check exercise the enhancement-layer hook (pre-pass derivation -> one picture call ->
// the reference's own writer) end to end without a GPU and compare the bitstream MD5 with the reference's.  Never part of the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/hlb200.h"

extern "C" int svc_emu_recon_batch(int bl, const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v,
                                   int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* motion,
                                   hlb200_svc_mb_state_t* state, hlb200_mb_coeffs_t* coeffs, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v);

struct hlb200_ctx {
    int w, h, nmb;
    uint8_t* src;
    uint8_t* slot[2];
    hlb200_svc_mb_state_t* state;
};
#define API extern "C" __attribute__((visibility("default")))
API int hlb200_init(int) { return HLB200_OK; }
API const char* hlb200_last_error(void) { return "svc_shim"; }
API int hlb200_stream_create(int w, int h, int, hlb200_ctx_t** out)
{
    hlb200_ctx* c = (hlb200_ctx*)calloc(1, sizeof(hlb200_ctx));
    const size_t fb = (size_t)w * h * 3 / 2;
    c->w = w; c->h = h; c->nmb = (w >> 4) * (h >> 4);
    c->src = (uint8_t*)calloc(fb, 1); c->slot[0] = (uint8_t*)calloc(fb, 1); c->slot[1] = (uint8_t*)calloc(fb, 1);
    c->state = (hlb200_svc_mb_state_t*)calloc((size_t)c->nmb, sizeof(hlb200_svc_mb_state_t));
    *out = c;
    return HLB200_OK;
}
API int hlb200_stream_destroy(hlb200_ctx_t* c) { free(c->src); free(c->slot[0]); free(c->slot[1]); free(c->state); free(c); return HLB200_OK; }
API int hlb200_frame_upload(hlb200_ctx_t* c, const uint8_t* y, const uint8_t* u, const uint8_t* v, int sy, int sc)
{
    const int W = c->w, H = c->h, Wc = W >> 1, Hc = H >> 1;
    for (int r = 0; r < H; ++r) memcpy(c->src + (size_t)r * W, y + (size_t)r * sy, W);
    for (int r = 0; r < Hc; ++r) { memcpy(c->src + (size_t)W * H + (size_t)r * Wc, u + (size_t)r * sc, Wc); memcpy(c->src + (size_t)W * H + (size_t)Wc * Hc + (size_t)r * Wc, v + (size_t)r * sc, Wc); }
    return HLB200_OK;
}
API int hlb200_slot_upload(hlb200_ctx_t* c, int s, const uint8_t* y, const uint8_t* u, const uint8_t* v)
{
    const size_t ysz = (size_t)c->w * c->h, csz = ysz >> 2;
    memcpy(c->slot[s], y, ysz); memcpy(c->slot[s] + ysz, u, csz); memcpy(c->slot[s] + ysz + csz, v, csz);
    return HLB200_OK;
}
API int hlb200_slot_download(hlb200_ctx_t* c, int s, uint8_t* y, uint8_t* u, uint8_t* v)
{
    const size_t ysz = (size_t)c->w * c->h, csz = ysz >> 2;
    memcpy(y, c->slot[s], ysz); memcpy(u, c->slot[s] + ysz, csz); memcpy(v, c->slot[s] + ysz + csz, csz);
    return HLB200_OK;
}
API int hlb200_slice_encode(hlb200_ctx_t*, const hlb200_slice_params_t*, hlb200_mb_record_t*) { return HLB200_ERR_SYSTEM; }   // base layer: not emulated here
API int hlb200_svc_layer_picture(hlb200_ctx_t* c, int ref_slot, int cur_slot, int qp, int off, const hlb200_mb_motion_t* motion, const uint8_t* pred_y, const uint8_t* pred_u,
                                 const uint8_t* pred_v, hlb200_mb_coeffs_t* out)
{
    const size_t ysz = (size_t)c->w * c->h, csz = ysz >> 2;
    const uint8_t *s = c->src, *ry, *ru, *rv;
    uint8_t* o = c->slot[cur_slot];
    if (ref_slot < 0) { ry = pred_y; ru = pred_u; rv = pred_v; }
    else { ry = c->slot[ref_slot]; ru = ry + ysz; rv = ru + csz; }
    return svc_emu_recon_batch(ref_slot < 0, s, s + ysz, s + ysz + csz, ry, ru, rv, c->w, c->h, 1, 0, qp, off, motion, c->state, out, o, o + ysz, o + ysz + csz);
}
