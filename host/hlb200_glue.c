/*
 * hlb200_glue.c -- the reference-side binding of libhl_b200.so: drops the B200 pixel hot path in under source/h264
 * WITHOUT editing any reference file.  Link this object with the reference library and
 *     -Wl,--wrap=hl_codec_264_nal_slice_data_encode
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_avc
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_avc
 * (or, inside the reference tree, call hlb200_glue_slice_begin() at the top of hl_codec_264_nal_slice_data_encode and replace
 * the two guess functions' bodies by hlb200_glue_apply()).
 *
 * Flow per picture (hl_codec_264_nal_slice_data_encode, source/h264/hl_codec_264_slice.c:1701):
 *   1. upload the source picture, run hlb200_slice_encode (ME + mode decision + reconstruction on the device; the device
 *      keeps the reconstructed pictures, so nothing but decision records comes back);
 *   2. call the ORIGINAL slice function: its MB loop still does init_mb_current / default quant values / the real CAVLC writer
 *      (_hl_codec_264_mb_write_no_pcm, source/h264/hl_codec_264_mb.c:543); the two decision functions it calls per MB are
 *      replaced by copies of the device's records into hl_codec_264_mb_t.
 * Host code stays C; this file is host-side product code (not test infrastructure) but can only be compiled where the
 * reference headers are available.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hartallo/hl_api.h"
#include "hartallo/hl_codec.h"
#include "hartallo/hl_frame.h"
#include "hartallo/hl_debug.h"
#include "hartallo/h264/hl_codec_264.h"
#include "hartallo/h264/hl_codec_264_mb.h"
#include "hartallo/h264/hl_codec_264_layer.h"
#include "hartallo/h264/hl_codec_264_encode.h"
#include "hartallo/h264/hl_codec_264_slice.h"
#include "hartallo/h264/hl_codec_264_pict.h"
#include "hartallo/h264/hl_codec_264_dpb.h"
#include "hartallo/h264/hl_codec_264_pps.h"
#include "hartallo/h264/hl_codec_264_macros.h"

#include "hlb200.h"

static struct {
    hlb200_ctx_t* ctx;
    int w, h, nmb, nslots;
    hlb200_mb_record_t* rec;
    const void* fs_of_slot[HLB200_MAX_REFS + 1];   /* host frame store whose picture lives in each device slot */
    int is_p;
} g;

static HL_ERROR_T glue_fail(const char* what, int rc)
{
    HL_DEBUG_ERROR("hlb200: %s failed (%d): %s", what, rc, hlb200_last_error());
    return rc == HLB200_ERR_OUTOFMEMORY ? HL_ERROR_OUTOFMEMMORY : HL_ERROR_SYSTEM;   /* no fallback: the error is propagated */
}

/* copies one decision record into the macroblock object: exactly the fields the writer reads (mb.c:584-860, residual.c:903-1094) */
static void glue_apply(hl_codec_264_mb_t* p_mb, const hlb200_mb_record_t* r, int32_t* pi_mad)
{
    int p, s, b, i, c;
    p_mb->mb_type = r->mb_type;
    p_mb->mb_qp_delta = r->mb_qp_delta;
    p_mb->coded_block_pattern = r->coded_block_pattern;
    p_mb->CodedBlockPatternLuma = r->cbp_luma;
    p_mb->CodedBlockPatternChroma = r->cbp_chroma;
    p_mb->CodedBlockPatternLuma4x4 = r->cbp_luma4x4;
    for (c = 0; c < 2; ++c) {
        p_mb->CodedBlockPatternChromaDC4x4[c] = r->cbp_chroma_dc4x4[c];
        p_mb->CodedBlockPatternChromaAC4x4[c] = r->cbp_chroma_ac4x4[c];
        for (b = 0; b < 4; ++b) {
            p_mb->ChromaDCLevel[c][b] = r->chroma_dc_level[c][b];
            for (i = 0; i < 16; ++i) p_mb->ChromaACLevel[c][b][i] = r->chroma_ac_level[c][b][i];
        }
    }
    if (r->mb_class == HLB200_MB_I16x16 || r->mb_class == HLB200_MB_I4x4) {
        const int i16 = r->mb_class == HLB200_MB_I16x16;
        p_mb->e_type = i16 ? HL_CODEC_264_MB_TYPE_I_16X16_0_0_0 : HL_CODEC_264_MB_TYPE_I_NXN;
        p_mb->flags_type = i16 ? HL_CODEC_264_MB_TYPE_FLAGS_INTRA_16x16 : HL_CODEC_264_MB_TYPE_FLAGS_INTRA_4x4;
        p_mb->MbPartPredMode[0] = i16 ? HL_CODEC_264_MB_MODE_INTRA_16X16 : HL_CODEC_264_MB_MODE_INTRA_4X4;
        p_mb->NumMbPart = 1;
        p_mb->Intra16x16PredMode = (HL_CODEC_264_I16x16_MODE_T)r->i16_pred_mode;
        p_mb->intra_chroma_pred_mode = r->intra_chroma_pred_mode;
        for (b = 0; b < 16; ++b) {
            p_mb->Intra4x4PredMode[b] = (HL_CODEC_264_I4x4_MODE_T)r->i4_pred_mode[b];
            p_mb->prev_intra4x4_pred_mode_flag[b] = r->prev_intra4x4_pred_mode_flag[b];
            p_mb->rem_intra4x4_pred_mode[b] = r->rem_intra4x4_pred_mode[b];
            p_mb->Intra16x16DCLevel[b] = r->i16_dc_level[b];
            for (i = 0; i < 16; ++i) {
                p_mb->Intra16x16ACLevel[b][i] = r->i16_ac_level[b][i];
                p_mb->LumaLevel[b][i] = r->luma_level[b][i];
            }
        }
    }
    else {
        static const int W[4] = { 16, 16, 8, 8 }, H[4] = { 16, 8, 16, 8 }, N[4] = { 1, 2, 2, 4 };
        static const HL_CODEC_264_MB_TYPE_T T[4] = { HL_CODEC_264_MB_TYPE_P_L0_16X16, HL_CODEC_264_MB_TYPE_P_L0_L0_16X8, HL_CODEC_264_MB_TYPE_P_L0_L0_8X16, HL_CODEC_264_MB_TYPE_P_8X8REF0 };
        static const int SW[4] = { 8, 8, 4, 4 }, SH[4] = { 8, 4, 8, 4 }, SN[4] = { 1, 2, 2, 4 };
        const int skip = r->mb_class == HLB200_MB_P_SKIP, pm = r->part_mode;
        p_mb->e_type = skip ? HL_CODEC_264_MB_TYPE_P_SKIP : T[pm];
        p_mb->flags_type = skip ? (HL_CODEC_264_MB_TYPE_FLAGS_INTER | HL_CODEC_264_MB_TYPE_FLAGS_SKIP) : HL_CODEC_264_MB_TYPE_FLAGS_INTER;
        p_mb->mb_type = (p_mb->e_type - HL_CODEC_264_MB_TYPE_START_SLICE_P_AND_SP - 1);
        p_mb->NumMbPart = N[pm]; p_mb->MbPartWidth = W[pm]; p_mb->MbPartHeight = H[pm];
        for (p = 0; p < 4; ++p) {
            const int sm = pm == 3 ? r->sub_mode[p] : 0;
            p_mb->MbPartPredMode[p] = HL_CODEC_264_MB_MODE_PRED_L0;
            p_mb->predFlagL0[p] = p_mb->PredFlagL0[p] = p < N[pm];
            p_mb->refIdxL0[p] = p_mb->RefIdxL0[p] = p < N[pm] ? r->ref_idx[p] : 0;
            if (pm == 3) {
                p_mb->SubMbPredType[p] = (HL_CODEC_264_SUBMB_TYPE_T)(HL_CODEC_264_SUBMB_TYPE_P_L0_8X8 + sm);
                p_mb->SubMbPredMode[p] = HL_CODEC_264_SUBMB_MODE_PRED_L0;
                p_mb->sub_mb_type[p] = sm;
                p_mb->NumSubMbPart[p] = SN[sm]; p_mb->SubMbPartWidth[p] = SW[sm]; p_mb->SubMbPartHeight[p] = SH[sm];
            }
            else {
                p_mb->NumSubMbPart[p] = 1; p_mb->SubMbPartWidth[p] = W[pm]; p_mb->SubMbPartHeight[p] = H[pm];
            }
            for (s = 0; s < 4; ++s) {
                p_mb->mvL0[p][s].x = p_mb->MvL0[p][s].x = r->mv[p][s][0];
                p_mb->mvL0[p][s].y = p_mb->MvL0[p][s].y = r->mv[p][s][1];
                p_mb->mvd_l0[p][s].x = r->mvd[p][s][0];
                p_mb->mvd_l0[p][s].y = r->mvd[p][s][1];
            }
        }
        for (b = 0; b < 16; ++b) for (i = 0; i < 16; ++i) p_mb->LumaLevel[b][i] = r->luma_level[b][i];
    }
    if (pi_mad) *pi_mad = r->mad;
}

extern HL_ERROR_T __real_hl_codec_264_nal_slice_data_encode(hl_codec_264_t*, hl_codec_264_encode_slice_data_t*);
HL_ERROR_T __wrap_hl_codec_264_nal_slice_data_encode(hl_codec_264_t* p_codec, hl_codec_264_encode_slice_data_t* p_esd)
{
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    const hl_codec_264_nal_slice_header_t* hdr = p_esd->pc_slice->p_header;
    const hl_frame_video_t* frame = p_codec->encoder.pc_frame;
    const int W = (int)hdr->PicWidthInSamplesL, H = (int)hdr->PicHeightInSamplesL;
    hlb200_slice_params_t prm;
    int rc, u, s, cur = -1;

    if (p_codec->layers.currDQId > 0 || p_esd->i_mb_start != 0 || p_esd->i_mb_end != (int32_t)hdr->PicSizeInMbs) {
        HL_DEBUG_ERROR("hlb200: only single-slice AVC pictures are supported by the device path");
        return HL_ERROR_NOT_IMPLEMENTED;
    }
    if (!g.ctx || g.w != W || g.h != H) {
        const char* dev = getenv("HLB200_DEVICE");
        int refs = (int)p_codec->pc_base->max_ref_frame;
        if (refs < 1) refs = 1;
        if (refs > HLB200_MAX_REFS) refs = HLB200_MAX_REFS;
        if (g.ctx) { hlb200_stream_destroy(g.ctx); g.ctx = NULL; free(g.rec); }
        if ((rc = hlb200_init(dev ? atoi(dev) : 0))) return glue_fail("hlb200_init", rc);
        if ((rc = hlb200_stream_create(W, H, refs, &g.ctx))) return glue_fail("hlb200_stream_create", rc);
        g.w = W; g.h = H; g.nmb = (W >> 4) * (H >> 4); g.nslots = refs + 1;
        g.rec = (hlb200_mb_record_t*)malloc(sizeof(hlb200_mb_record_t) * (size_t)g.nmb);
        memset(g.fs_of_slot, 0, sizeof(g.fs_of_slot));
    }
    memset(&prm, 0, sizeof(prm));
    g.is_p = IsSliceHeaderP(hdr) ? 1 : 0;
    prm.slice_type = g.is_p;
    prm.qp = p_codec->encoder.rc.b_enabled ? p_codec->encoder.rc.qp : p_codec->encoder.i_qp;
    prm.me_range = (int32_t)p_codec->pc_base->me_range;
    prm.chroma_qp_index_offset = hdr->pc_pps->chroma_qp_index_offset;
    prm.num_refs = g.is_p ? (int32_t)hdr->num_ref_idx_l0_active_minus1 + 1 : 0;
    /* device slots of the reference pictures (RefPicList0 order), then a free slot for the current picture */
    for (u = 0; u < prm.num_refs; ++u) {
        const void* fs = pc_layer->pobj_poc->RefPicList0[u];
        prm.ref_slot[u] = -1;
        for (s = 0; s < g.nslots; ++s) if (fs && g.fs_of_slot[s] == fs) prm.ref_slot[u] = s;
        if (prm.ref_slot[u] < 0) { HL_DEBUG_ERROR("hlb200: reference picture %d is not resident on the device", u); return HL_ERROR_INVALID_STATE; }
    }
    for (s = 0; s < g.nslots && cur < 0; ++s) {
        int used = 0;
        for (u = 0; u < prm.num_refs; ++u) used |= (prm.ref_slot[u] == s);
        if (!used) cur = s;
    }
    prm.cur_slot = cur;
    for (s = 0; s < g.nslots; ++s) if (g.fs_of_slot[s] == (const void*)pc_layer->pc_fs_curr) g.fs_of_slot[s] = NULL;
    g.fs_of_slot[cur] = pc_layer->pc_fs_curr;

    if ((rc = hlb200_frame_upload(g.ctx, frame->data_ptr[0], frame->data_ptr[1], frame->data_ptr[2], W, W >> 1))) return glue_fail("hlb200_frame_upload", rc);
    if ((rc = hlb200_slice_encode(g.ctx, &prm, g.rec))) return glue_fail("hlb200_slice_encode", rc);
    if (getenv("HLB200_SYNC_RECON")) {   /* only when something on the host reads the reconstruction (MD5 hooks, decoder round trip) */
        const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
        if ((rc = hlb200_slot_download(g.ctx, cur, (uint8_t*)pict->pc_data_y, (uint8_t*)pict->pc_data_u, (uint8_t*)pict->pc_data_v))) return glue_fail("hlb200_slot_download", rc);
    }
    return __real_hl_codec_264_nal_slice_data_encode(p_codec, p_esd);
}

HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_inter_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    (void)p_codec;
    if (!g.rec || p_mb->u_addr >= (uint32_t)g.nmb) return HL_ERROR_INVALID_STATE;
    glue_apply(p_mb, &g.rec[p_mb->u_addr], pi_mad);
    return HL_ERROR_SUCCESS;
}

HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_intra_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    (void)p_codec;
    if (!g.rec || p_mb->u_addr >= (uint32_t)g.nmb) return HL_ERROR_INVALID_STATE;
    glue_apply(p_mb, &g.rec[p_mb->u_addr], pi_mad);
    return HL_ERROR_SUCCESS;
}
