// CPU stand-in for the libhl_b200.so entry points host/hlb200_glue.c calls, backed by the CPU run of the device sources (svc_emu.cpp for the SVC kernel;
// with -DSVC_SHIM_WITH_SLICE also hlb_mbcore.cuh, the slice kernel's per-macroblock code, for the base layer).  TEST INFRASTRUCTURE: lets oracle/_ref/hl_svc_glue_check exercise the enhancement-layer hook (pre-pass derivation -> one picture call ->
// the reference's own writer) end to end without a GPU and compare the bitstream MD5 with the reference's.  Never part of the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/hlb200.h"
#ifdef SVC_SHIM_WITH_SLICE   /* also stand in for hlb200_slice_encode: the slice kernel's per-macroblock source run in raster order, as tools/emu/emu_main.cpp does */
#include <math.h>
#include <stdio.h>
int g_emu_dbg = 0;
#define HLB_EMU_DEBUG 1
#include "../../hartallo_b200/csrc/hlb_mbcore.cuh"
#include "../../hartallo_b200/csrc/hlb_bits.cuh"
#include "../../hartallo_b200/csrc/hlb_deblock.cuh"
#endif

extern "C" int svc_emu_recon_batch(int bl, const uint8_t* src_y, const uint8_t* src_u, const uint8_t* src_v, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v,
                                   int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* motion,
                                   hlb200_svc_mb_state_t* state, hlb200_mb_coeffs_t* coeffs, uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v);

struct hlb200_ctx {
    int w, h, nmb, nslots;
    uint8_t* src;
    uint8_t* slot[HLB200_MAX_REFS + 1];
    hlb200_svc_mb_state_t* state;
    uint8_t* had_parts;   /* hlb200_svc_layer_picture_derived: carried per macroblock from picture to picture */
#ifdef SVC_SHIM_WITH_SLICE
    hlb::MbState* mbstate;
    hlb::MbWork* work;
    int chain;
    hlb200_mb_record_t* rec;      /* records of the last picture (hlb200_slice_encode_async) */
    uint32_t* bits; uint32_t nbits;
#endif
};
#define API extern "C" __attribute__((visibility("default")))
API int hlb200_init(int) { return HLB200_OK; }
API int hlb200_host_register(void*, size_t) { return HLB200_OK; }
API int hlb200_host_unregister(void*) { return HLB200_OK; }
API const char* hlb200_last_error(void) { return "svc_shim"; }
API int hlb200_stream_create(int w, int h, int max_refs, hlb200_ctx_t** out)
{
    hlb200_ctx* c = (hlb200_ctx*)calloc(1, sizeof(hlb200_ctx));
    const size_t fb = (size_t)w * h * 3 / 2;
    c->w = w; c->h = h; c->nmb = (w >> 4) * (h >> 4); c->nslots = max_refs + 1;
    c->src = (uint8_t*)calloc(fb, 1);
    for (int s = 0; s < c->nslots; ++s) c->slot[s] = (uint8_t*)calloc(fb, 1);
    c->state = (hlb200_svc_mb_state_t*)calloc((size_t)c->nmb, sizeof(hlb200_svc_mb_state_t));
#ifdef SVC_SHIM_WITH_SLICE
    c->mbstate = (hlb::MbState*)calloc((size_t)c->nmb, sizeof(hlb::MbState));
    c->work = (hlb::MbWork*)calloc(1, sizeof(hlb::MbWork));
    c->rec = (hlb200_mb_record_t*)calloc((size_t)c->nmb, sizeof(hlb200_mb_record_t));
    c->bits = (uint32_t*)calloc((size_t)c->nmb * HLB200_BITS_WORDS_PER_MB + 64, sizeof(uint32_t));
#endif
    *out = c;
    return HLB200_OK;
}
API int hlb200_stream_destroy(hlb200_ctx_t* c)
{
    free(c->src); free(c->state); free(c->had_parts);
    for (int s = 0; s < c->nslots; ++s) free(c->slot[s]);
#ifdef SVC_SHIM_WITH_SLICE
    free(c->mbstate); free(c->work); free(c->rec); free(c->bits);
#endif
    free(c);
    return HLB200_OK;
}
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
#ifdef SVC_SHIM_WITH_SLICE
namespace {
struct CpuExec {   // the executor of tools/emu/emu_main.cpp: lanes become loops
    hlb::MbWork* w;
    const hlb::FrameCtx* f;
    int prev_frame_sctr;
    void run(int cmd, int nlanes)
    {
        w->cmd = cmd; w->arg0_lanes = nlanes;
        const int np = hlb::cmd_phases(cmd);
        for (int p = 0; p < np; ++p)
            for (int lane = 0; lane < nlanes; ++lane) hlb::cmd_phase(*w, *f, cmd, p, lane);
    }
    void trials(int nlanes) { for (int lane = 0; lane < nlanes; ++lane) hlb::me_phase_trial(*w, *f, lane); }
    int lane() const { return 0; }
    int nlanes() const { return 1; }
    void sync() const {}
    // cross-lane reductions of the master warp: one lane here, so each is the identity
    int reduce_min(int v) const { return v; }
    int reduce_max(int v) const { return v; }
    int reduce_add(int v) const { return v; }
    void reduce_argmin(double&, int&) const {}
    int prev_sctr(int mb)
    {
        for (int a = mb - 1; a >= 0; --a)
            if (f->st[a].last_sctr != 255) return f->st[a].last_sctr;
        return prev_frame_sctr;
    }
};
}
// build_job() of hartallo_b200/csrc/hlb_slice.cu, then the macroblocks in raster order
API int hlb200_slice_encode(hlb200_ctx_t* c, const hlb200_slice_params_t* p, hlb200_mb_record_t* out)
{
    const size_t ys = (size_t)c->w * c->h, cs = ys / 4;
    hlb::FrameCtx f;
    memset(&f, 0, sizeof(f));
    f.W = c->w; f.H = c->h; f.mbw = c->w / 16; f.mbh = c->h / 16;
    f.qp = p->qp;
    { int q = p->qp + p->chroma_qp_index_offset; q = q < 0 ? 0 : (q > 51 ? 51 : q); f.qpc = hlb::kQpc[q]; }
    f.is_p = p->slice_type == 1;
    f.me_range = p->me_range < 1 ? 1 : (p->me_range > 64 ? 64 : p->me_range);
    f.num_refs = f.is_p ? p->num_refs : 0;
    f.early_term = p->me_early_term_flag != 0;
    f.lambda = 0.852 * (double)(1 << ((p->qp - 12) / 3));
    hlb::frame_ctx_derive(f);
    f.src[0] = c->src; f.src[1] = c->src + ys; f.src[2] = c->src + ys + cs;
    f.cur[0] = c->slot[p->cur_slot]; f.cur[1] = f.cur[0] + ys; f.cur[2] = f.cur[1] + cs;
    for (int u = 0; u < f.num_refs; ++u) { f.ref[u][0] = c->slot[p->ref_slot[u]]; f.ref[u][1] = f.ref[u][0] + ys; f.ref[u][2] = f.ref[u][1] + cs; }
    f.st = c->mbstate; f.rec = out;
    CpuExec x{c->work, &f, c->chain};
    for (int mb = 0; mb < c->nmb; ++mb) hlb::mb_encode(x, *c->work, f, mb);
    c->chain = x.prev_sctr(c->nmb);
    if (p->deblock_flag) {   // k_dbk_bs + k_dbk of hlb_slice.cu
        hlb::DbkJob d;
        memset(&d, 0, sizeof(d));
        for (int k = 0; k < 3; ++k) d.plane[k] = f.cur[k];
        d.rec = out; d.W = f.W; d.H = f.H; d.mbw = f.mbw; d.mbh = f.mbh; d.enabled = 1;
        hlb::dbk_job_thresholds(d, f.qp, f.qpc);
        hlb::dbk_picture_serial(d);
    }
    return HLB200_OK;
}
API int hlb200_slice_encode_async(hlb200_ctx_t* c, const hlb200_slice_params_t* p) { return hlb200_slice_encode(c, p, c->rec); }
API int hlb200_slice_encode_batch_async(hlb200_ctx_t** ctxs, const hlb200_slice_params_t* p, int n)
{
    for (int i = 0; i < n; ++i) { const int rc = hlb200_slice_encode(ctxs[i], p + i, ctxs[i]->rec); if (rc) return rc; }
    return HLB200_OK;
}
// the three kernels of hlb_slice.cu (k_bits_len, k_bits_scan, k_bits_write) as loops over the same per-macroblock code (hlb_bits.cuh)
API int hlb200_slice_bits_batch_async(hlb200_ctx_t** ctxs, const int32_t* types, int n)
{
    for (int i = 0; i < n; ++i) {
        hlb200_ctx* c = ctxs[i];
        std::vector<uint32_t> len((size_t)c->nmb);
        hlb::BitsJob j;
        j.rec = c->rec; j.st = c->mbstate; j.len = len.data(); j.out = c->bits; j.hdr = nullptr; j.nmb = c->nmb; j.mbw = c->w >> 4; j.is_p = types[i]; j.cap_words = c->nmb * HLB200_BITS_WORDS_PER_MB + 64;
        uint32_t total = 0;
        for (int mb = 0; mb < c->nmb; ++mb) { hlb::BitCount k; k.n = 0; hlb::bits_put_mb(k, j, mb); len[mb] = total; total += k.n; }
        memset(c->bits, 0, sizeof(uint32_t) * ((total + 31) / 32 + 2));
        for (int mb = 0; mb < c->nmb; ++mb) {
            hlb::BitWriter w;
            w.buf = c->bits; w.pos = len[mb]; w.acc = 0; w.nacc = 0;
            hlb::bits_put_mb(w, j, mb);
            w.finish();
            if (w.pos != (mb + 1 < c->nmb ? len[mb + 1] : total)) return HLB200_ERR_INVALID_STATE;   // the counting and the writing sink must agree
        }
        c->nbits = total;
    }
    return HLB200_OK;
}
API int hlb200_slice_bits_download(hlb200_ctx_t* c, uint32_t* out, size_t cap, uint32_t* nbits)
{
    const size_t words = ((size_t)c->nbits + 31) / 32;
    if (words > cap) return HLB200_ERR_OUTOFMEMORY;
    memcpy(out, c->bits, sizeof(uint32_t) * words);
    *nbits = c->nbits;
    return HLB200_OK;
}
#else
API int hlb200_slice_encode(hlb200_ctx_t*, const hlb200_slice_params_t*, hlb200_mb_record_t*) { return HLB200_ERR_SYSTEM; }   // base layer: not emulated in this build
API int hlb200_slice_encode_async(hlb200_ctx_t*, const hlb200_slice_params_t*) { return HLB200_ERR_SYSTEM; }
API int hlb200_slice_encode_batch_async(hlb200_ctx_t**, const hlb200_slice_params_t*, int) { return HLB200_ERR_SYSTEM; }
API int hlb200_slice_bits_batch_async(hlb200_ctx_t**, const int32_t*, int) { return HLB200_ERR_SYSTEM; }
API int hlb200_slice_bits_download(hlb200_ctx_t*, uint32_t*, size_t, uint32_t*) { return HLB200_ERR_SYSTEM; }
#endif
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
API int hlb200_svc_layer_picture_resampled(hlb200_ctx_t* c, int cur_slot, int qp, int off, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v, int rw, int rh,
                                           int level_idc, hlb200_mb_coeffs_t* out);
API int hlb200_svc_layer_picture_resampled_from(hlb200_ctx_t* c, int cur_slot, int qp, int off, hlb200_ctx_t* r, int rslot, int level_idc, hlb200_mb_coeffs_t* out)
{
    const size_t rys = (size_t)r->w * r->h, rcs = rys >> 2;
    const uint8_t* p = r->slot[rslot];
    return hlb200_svc_layer_picture_resampled(c, cur_slot, qp, off, p, p + rys, p + rys + rcs, r->w, r->h, level_idc, out);
}
extern "C" int svc_emu_derive_motion(const hlb200_svc_base_mb_t* base, const hlb200_svc_layer_geom_t* geom, int width, int height, uint8_t* had_parts, hlb200_mb_motion_t* motion,
                                     int32_t* status);
API int hlb200_svc_layer_picture_derived(hlb200_ctx_t* c, int ref_slot, int cur_slot, int qp, int off, const hlb200_svc_base_mb_t* base, const hlb200_svc_layer_geom_t* geom,
                                         hlb200_mb_motion_t* out_motion, int32_t* out_status, hlb200_mb_coeffs_t* out)
{
    if (!c->had_parts) c->had_parts = (uint8_t*)calloc((size_t)c->nmb, 1);
    std::vector<hlb200_mb_motion_t> m((size_t)c->nmb);
    int rc = svc_emu_derive_motion(base, geom, c->w, c->h, c->had_parts, m.data(), out_status);
    if (rc) return rc;
    if (out_motion) memcpy(out_motion, m.data(), sizeof(hlb200_mb_motion_t) * (size_t)c->nmb);
    if (*out_status) return HLB200_ERR_NOT_IMPLEMENTED;
    return hlb200_svc_layer_picture(c, ref_slot, cur_slot, qp, off, m.data(), nullptr, nullptr, nullptr, out);
}
extern "C" int svc_emu_resample_plane(const uint8_t* ref, int refW, int refH, uint8_t* out, int W, int H, int chroma, int level_idc);
API int hlb200_svc_layer_picture_resampled(hlb200_ctx_t* c, int cur_slot, int qp, int off, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v, int rw, int rh,
                                           int level_idc, hlb200_mb_coeffs_t* out)
{
    const int dims[4] = {rw, rh, rw >> 1, rh >> 1};
    for (int i = 0; i < 4; ++i) if (level_idc > 30 && (dims[i] & (dims[i] - 1)) == 0) return HLB200_ERR_INVALID_PARAMETER;   // as the library: (G-43) overflows the reference's int32
    if (rw > c->w || rh > c->h) return HLB200_ERR_INVALID_PARAMETER;
    const size_t ysz = (size_t)c->w * c->h, csz = ysz >> 2;
    uint8_t* pred = (uint8_t*)malloc(ysz + 2 * csz);
    svc_emu_resample_plane(ref_y, rw, rh, pred, c->w, c->h, 0, level_idc);
    svc_emu_resample_plane(ref_u, rw >> 1, rh >> 1, pred + ysz, c->w >> 1, c->h >> 1, 1, level_idc);
    svc_emu_resample_plane(ref_v, rw >> 1, rh >> 1, pred + ysz + csz, c->w >> 1, c->h >> 1, 1, level_idc);
    const int rc = hlb200_svc_layer_picture(c, -1, cur_slot, qp, off, nullptr, pred, pred + ysz, pred + ysz + csz, out);
    free(pred);
    return rc;
}
