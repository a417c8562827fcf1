/*
 * hlb200_glue.c -- the reference-side binding of libhl_b200.so: drops the B200 pixel hot path in under source/h264
 * WITHOUT editing any reference file.  Link this object with the reference library and
 *     -Wl,--wrap=hl_codec_264_nal_slice_data_encode
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_avc
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_avc
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_svc      (SVC enhancement layers)
 *     -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_svc
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
#include "hartallo/h264/hl_codec_264_bits.h"
#include "hartallo/h264/hl_codec_264_rbsp.h"

#include "hlb200.h"

#include "hlb200_glue.h"

/* one entry per codec instance (= per stream): many streams may live in one process, each with its own device context */
typedef struct glue_stream_s {
    const void* codec;                             /* key: the hl_codec_264_t of the stream */
    hlb200_ctx_t* ctx;
    int w, h, nmb, nslots;
    hlb200_mb_record_t* rec;                       /* records mode only */
    uint32_t* bits; size_t bits_cap;               /* bits mode: landing buffer of the device-written slice data */
    const void* fs_of_slot[HLB200_MAX_REFS + 1];   /* host frame store whose picture lives in each device slot */
    int is_p;
    int svc;                                       /* the stream has enhancement layers: they derive their motion from this layer's macroblock objects */
} glue_stream_t;
#define GLUE_MAX_STREAMS 1024
static glue_stream_t g_streams[GLUE_MAX_STREAMS];
static int g_nstreams = 0;
static glue_stream_t* g_cur = NULL;                /* stream whose reference macroblock loop is running (records mode) */
static glue_stream_t* glue_stream_of(const void* codec, int create)
{
    int i;
    for (i = 0; i < g_nstreams; ++i) if (g_streams[i].codec == codec) return &g_streams[i];
    if (!create || g_nstreams >= GLUE_MAX_STREAMS) return NULL;
    memset(&g_streams[g_nstreams], 0, sizeof(glue_stream_t));
    g_streams[g_nstreams].codec = codec;
    return &g_streams[g_nstreams++];
}

/* ---- batch mode (hlb200_glue.h): the slice hooks of many codec instances submit their pictures, ONE launch encodes and serialises them all ---- */
static struct {
    int active;
    hlb200_glue_yield_fn yield; void* yield_arg;
    int n, rc;
    hlb200_ctx_t* ctxs[GLUE_MAX_STREAMS];
    hlb200_slice_params_t prm[GLUE_MAX_STREAMS];
    int32_t types[GLUE_MAX_STREAMS];
} g_batch;
void hlb200_glue_batch_begin(hlb200_glue_yield_fn yield, void* arg) { g_batch.active = 1; g_batch.yield = yield; g_batch.yield_arg = arg; g_batch.n = 0; g_batch.rc = 0; }
void hlb200_glue_batch_end(void) { g_batch.active = 0; g_batch.n = 0; }
int hlb200_glue_batch_pending(void) { return g_batch.n; }
int hlb200_glue_batch_flush(void)
{
    int rc = 0;
    if (g_batch.n > 0) {
        rc = hlb200_slice_encode_batch_async(g_batch.ctxs, g_batch.prm, g_batch.n);
        if (!rc) rc = hlb200_slice_bits_batch_async(g_batch.ctxs, g_batch.types, g_batch.n);
    }
    g_batch.rc = rc; g_batch.n = 0;
    return rc;
}

/* Which GPU a layer's context lives on: HLB200_SVC_DEVICES = comma-separated device ordinals, layer 0 first (the layers of an SVC stream sharded over GPUs,
 * "0,1,2" = one layer per GPU); layers beyond the list, and streams without the variable, use HLB200_DEVICE (default 0).  The host code is single-threaded: the device
 * of the context about to be used is made current before every group of library calls. */
static int g_dev_cur = -1;
static int glue_layer_device(int li)
{
    const char* list = getenv("HLB200_SVC_DEVICES");
    const char* one = getenv("HLB200_DEVICE");
    int i;
    for (i = 0; list && *list; ++i) {
        if (i == li) return atoi(list);
        list = strchr(list, ',');
        if (list) ++list;
    }
    return one ? atoi(one) : 0;
}
static int glue_use_device(int dev)
{
    int rc = 0;
    if (dev != g_dev_cur && !(rc = hlb200_init(dev))) g_dev_cur = dev;
    return rc;
}

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
        if (g_cur->svc && g_cur->is_p) {
            /* The enhancement layers' inter-layer derivation reads this object (utils.c:1701,1807-1834).  An Intra16x16 macroblock is recognised as intra by its
             * e_type whatever else it holds; an Intra4x4 macroblock of a P picture is NOT (HL_CODEC_264_MB_TYPE_IS_I_4X4 tests the type that only I pictures get
             * patched in, mb.c:326-345), so the reference derives enhancement-layer motion from it.  What that derivation can see of it: flags_type is intra
             * (rdo.c:194), so hl_codec_264_mb_get_sub_partition_indices (mb.h:313-339) answers partition 0 / sub-partition 0 for every 4x4 block, and
             * utils.c:1807-1834 then reads predFlagL0[0], refIdxL0[0], mvL0[0][0] -- the LOWER-case fields.  The base layer's search (SVCExtFlag = 0) works on the
             * upper-case ones (rdo.c:849-852) and only marks predFlagL0[0] = 1 (me_ds.c:183, the 16x16 mode is always tried); when intra wins nothing commits
             * (rdo.c:1161-1167), so refIdxL0[0] / mvL0[0][0] are those of this macroblock's LAST INTER COMMIT in an earlier picture (0 if none: the object is
             * never reset, utils.c:73-89).  The object here is the same persistent one and the inter branch below writes the same values at [0][0] as
             * rdo.c:1180,1194, so leaving them alone reproduces the reference. */
            p_mb->MbPartWidth = p_mb->MbPartHeight = 16;
            p_mb->predFlagL0[0] = 1;
            for (p = 0; p < 4; ++p) p_mb->NumSubMbPart[p] = 1, p_mb->SubMbPartWidth[p] = p_mb->SubMbPartHeight[p] = 16;
        }
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

/* Settings the device path does not reproduce are REFUSED, never silently diverged from and never handed to the reference's CPU path:
 *   deblock_flag != 0 with SVC layers   the inter-layer deblocking of deblock.c:175-186 and the host-side layer derivation that reads the base picture
 *                                       (single-layer streams: the device filters the picture, hlb_deblock.cuh; library default 1, hl_types.h:69)
 *   rate control                        QP changes per picture / macroblock (rc.c)
 * me_early_term_flag (library default 1, hl_types.h:67) is reproduced on the device (homogeneous-block mode mask, rdo.c:889-935).
 * test_encoder.c:139-146 itself runs with rc_bitrate = -1, deblock_flag = 0 and me_early_term_flag = 0. */
static HL_ERROR_T glue_check_settings(const hl_codec_264_t* p_codec)
{
    if (p_codec->pc_base->deblock_flag && (p_codec->encoder.b_svc_enabled || getenv("HLB200_GLUE_RECORDS"))) {
        HL_DEBUG_ERROR("hlb200: deblock_flag = 1 is not implemented by the device path for streams with SVC layers (set deblock_flag = 0)");
        return HL_ERROR_NOT_IMPLEMENTED;
    }
    if (p_codec->encoder.rc.b_enabled) { HL_DEBUG_ERROR("hlb200: rate control is not implemented by the device path (set rc_bitrate <= 0)"); return HL_ERROR_NOT_IMPLEMENTED; }
    return HL_ERROR_SUCCESS;
}

/* ================================================================================================================================
 * SVC enhancement layers (currDQId > 0).  The reference does no search there (base_mode_flag = 1 for every macroblock): what it computes per
 * macroblock is (a) the inter-layer derivation of partitions / vectors (P; utils.c:1225 + :1498) or the resampling of the base reconstruction (I;
 * decode_svc.c:2864) -- both read the reference layer only -- and (b) prediction + residual coding + reconstruction,
 * hl_codec_264_rdo_mb_guess_best_inter_pred_svc rdo.c:1273 / ..._intra_pred_svc rdo.c:301.  All of it is independent per macroblock, so the hook hands the
 * WHOLE picture to the device in ONE call -- hlb200_svc_layer_picture_derived (P: derivation kernel + fused prediction / residual kernel, from the
 * reference layer's macroblock fields) or hlb200_svc_layer_picture_resampled (I: resampling kernel + residual kernel, from the reference layer's
 * reconstruction) -- and then lets the reference's own loop serialise the result.  Inside that loop the wrapped guess functions still call the reference's
 * derivation for each macroblock: the reference's writer and the NEXT layer's derivation read the macroblock object's syntax fields (mb_type, sub_mb_type,
 * refIdx, mv, flags), which is bookkeeping the north star leaves on the host; no sample and no level depends on it, and the device's field is cross-checked
 * against it macroblock by macroblock (glue_svc_apply).
 * ================================================================================================================================ */
#include "hartallo/h264/hl_codec_264_utils.h"
#include "hartallo/h264/hl_codec_264_sps.h"

#define GLUE_SVC_MAX_LAYERS 8
typedef struct glue_svc_layer_s {
    hlb200_ctx_t* ctx;
    int w, h, nmb;
    hlb200_mb_motion_t* motion;   /* the motion field the device derived for the picture (read back for the cross-check in glue_svc_apply) */
    hlb200_mb_coeffs_t* coeffs;
    uint8_t* rec;            /* tight Y|U|V */
    hlb200_svc_base_mb_t* base; int nbase;   /* the reference layer's macroblock fields, as uploaded for the derivation */
    int dev;                                 /* GPU of this layer's context (glue_layer_device) */
    const void* codec;                       /* the hl_codec_264_t this layer context belongs to: another instance gets a fresh context (carried state, resident pictures) */
    const void* fs_of_slot[2];               /* host frame store whose picture lives in each of the layer context's two device slots (reconstructions stay resident) */
} glue_svc_layer_t;
static glue_svc_layer_t g_svc[GLUE_SVC_MAX_LAYERS];
static glue_svc_layer_t* g_svc_active = NULL;
static int g_svc_intra = 0;

/* the part of the two guess functions that precedes the prediction (rdo.c:1318-1346 / rdo.c:353-368): defaults + G.8.1.5.1 (+ G.8.4.1 for P) */
static HL_ERROR_T glue_svc_derive(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int intra)
{
    HL_ERROR_T err;
    p_mb->ext.svc.base_mode_flag = 1;
    p_mb->mb_type = HL_CODEC_264_SVC_MB_TYPE_INFERRED;
    p_mb->e_type = HL_CODEC_264_MB_TYPE_SVC_I_BL;
    p_mb->flags_type = HL_CODEC_264_MB_TYPE_FLAGS_INFERRED;
    p_mb->MbPartPredMode[0] = intra ? HL_CODEC_264_MB_MODE_INTRA_BL : HL_CODEC_264_MB_MODE_INTER_BL;
    p_mb->NumMbPart = 1;
    p_mb->MbPartWidth = p_mb->MbPartHeight = 16;
    p_mb->CodedBlockPatternLuma4x4 = 0;
    if ((err = hl_codec_264_utils_derivation_process_initialisation_svc(p_codec, p_mb))) return err;
    if (!intra && (err = hl_codec_264_utils_derivation_process_for_mv_comps_and_ref_indices_svc(p_codec, p_mb))) return err;
    return HL_ERROR_SUCCESS;
}

/* what the reference's macroblock loop does before it calls the guess function (slice.c:1787-1865) */
static HL_ERROR_T glue_svc_loop_prologue(hl_codec_264_t* p_codec, hl_codec_264_encode_slice_data_t* p_esd, uint32_t addr, hl_codec_264_mb_t** pp_mb)
{
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    const hl_codec_264_nal_slice_header_t* hdr = p_esd->pc_slice->p_header;
    const hl_codec_264_nal_sps_t* pc_sps = hdr->pc_pps->pc_sps;
    hl_codec_264_mb_t* p_mb;
    HL_ERROR_T err;
    if (!(p_mb = pc_layer->pp_list_macroblocks[addr])) {
        if ((err = hl_codec_264_mb_create(addr, &p_mb))) return err;
        pc_layer->pp_list_macroblocks[addr] = p_mb;
    }
    p_mb->u_slice_idx = p_esd->pc_slice->u_idx;
    if ((err = hl_codec_264_utils_init_mb_current_avc(p_codec, addr, hdr->l_id, HL_FALSE))) return err;
    if (hdr->SVCExtFlag) {
        p_mb->mb_type = HL_CODEC_264_SVC_MB_TYPE_INFERRED;
        if (pc_sps->profile_idc == HL_CODEC_264_PROFILE_BASELINE_SVC) p_mb->ext.svc.InCropWindow = 1;
        else {
            const int32_t mbX = p_mb->u_x, mbY0 = p_mb->u_y, mbY1 = hdr->MbaffFrameFlag ? (mbY0 + 1) : mbY0, scalMbH = ((1 + hdr->field_pic_flag) << 4);
            p_mb->ext.svc.InCropWindow = (!hdr->ext.svc.NoInterLayerPredFlag && (mbX >= ((hdr->ext.svc.ScaledRefLayerLeftOffset + 15) >> 4)) &&
                                          (mbX < ((hdr->ext.svc.ScaledRefLayerLeftOffset + hdr->ext.svc.ScaledRefLayerPicWidthInSamplesL) >> 4)) &&
                                          (mbY0 >= ((hdr->ext.svc.ScaledRefLayerTopOffset + scalMbH - 1) / scalMbH)) &&
                                          (mbY1 < ((hdr->ext.svc.ScaledRefLayerTopOffset + hdr->ext.svc.ScaledRefLayerPicHeightInSamplesL) / scalMbH)));
        }
        p_mb->ext.svc.base_mode_flag = 1;
    }
    if ((err = hl_codec_264_mb_set_default_quant_values(p_mb, p_codec))) return err;
    *pp_mb = p_mb;
    return HL_ERROR_SUCCESS;
}

static HL_ERROR_T glue_svc_slice(hl_codec_264_t* p_codec, hl_codec_264_encode_slice_data_t* p_esd)
{
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    const hl_codec_264_nal_slice_header_t* hdr = p_esd->pc_slice->p_header;
    const hl_frame_video_t* frame = p_codec->encoder.pc_frame;
    const int W = (int)hdr->PicWidthInSamplesL, H = (int)hdr->PicHeightInSamplesL, Wc = W >> 1, Hc = H >> 1, mbw = W >> 4;
    const int li = (int)(p_codec->layers.currDQId >> 4), intra = (p_codec->encoder.encoding_curr == HL_VIDEO_ENCODING_TYPE_INTRA);
    const size_t ysz = (size_t)W * H, csz = (size_t)Wc * Hc;
    glue_svc_layer_t* L;
    HL_ERROR_T err;
    int rc, qp = -1, dev_rs = 0, level_idc = 0, ref_slot = 0, cur_slot = 1;
    int32_t status = 0;
    hlb200_svc_layer_geom_t geom;
    uint32_t addr;

    if (li <= 0 || li >= GLUE_SVC_MAX_LAYERS || (p_codec->layers.currDQId & 15) || p_esd->i_mb_start != 0 || p_esd->i_mb_end != (int32_t)hdr->PicSizeInMbs) {
        HL_DEBUG_ERROR("hlb200: only single-slice spatial enhancement layers are supported by the device path");
        return HL_ERROR_NOT_IMPLEMENTED;
    }
    /* I pictures: the Intra_Base resampling (rdo.c:363-377 -> decode_svc.c:2864) runs on the device (hlb200_svc_layer_picture_resampled) when the picture is inside what
     * hlb200_dev_svc_resample_intra_batch is pinned for (frame macroblocks, no cropping offsets, the chroma phases sps.c:810-813 writes, unconstrained resampling;
     * level_idc as utils.c:1075,1123 reads it selects the precision there).  Anything else is REFUSED (HL_ERROR_NOT_IMPLEMENTED): the product path never runs the reference's CPU resampling. */
    if (intra) {
        const hl_codec_264_nal_sps_t* sps = hdr->pc_pps->pc_sps;
        const hl_codec_264_layer_t* top = p_codec->layers.p_list[(p_codec->layers.currDQId >> 4) << 4];
        /* Small pictures are refused as well: the reference sizes its window array as PicSizeInMbs << 8 BYTES (layer.c:202) but fills it with refArrayW x refArrayH int32
         * (48x48 at the dyadic ratio, 64x64 at 1:1), so below 36 / 64 macroblocks it writes past the allocation and its own prediction depends on the heap (traced:
         * saturated rows in 96x32 pictures) -- there is no reference behaviour to reproduce. */
        const int dyadic = pc_layer->RefLayerPicWidthInSamplesL * 2 == (uint32_t)W && pc_layer->RefLayerPicHeightInSamplesL * 2 == (uint32_t)H;
        dev_rs = mbw * (H >> 4) >= (dyadic ? 36 : 64) && pc_layer->pc_ref && pc_layer->pc_ref->pc_fs_curr && pc_layer->pc_ref->pc_fs_curr->p_pict && pc_layer->RefLayerFrameMbsOnlyFlag && sps->frame_mbs_only_flag &&
                 !hdr->field_pic_flag && sps->ChromaArrayType == 1 && sps->p_svc && sps->p_svc->chroma_phase_x_plus1_flag == 1 && sps->p_svc->chroma_phase_y_plus1 == 1 &&
                 hdr->ext.svc.ref_layer_chroma_phase_x_plus1_flag == 1 && hdr->ext.svc.ref_layer_chroma_phase_y_plus1 == 1 && !hdr->ext.svc.constrained_intra_resampling_flag &&
                 hdr->ext.svc.ScaledRefLayerLeftOffset == 0 && hdr->ext.svc.ScaledRefLayerTopOffset == 0 && hdr->ext.svc.ScaledRefLayerPicWidthInSamplesL == W &&
                 hdr->ext.svc.ScaledRefLayerPicHeightInSamplesL == H && top && top->pc_slice_hdr &&
                 !(pc_layer->RefLayerPicWidthInSamplesL & 15) && !(pc_layer->RefLayerPicHeightInSamplesL & 15) && pc_layer->RefLayerPicWidthInSamplesL <= W &&
                 pc_layer->RefLayerPicHeightInSamplesL <= H;
        if (!dev_rs) {
            HL_DEBUG_ERROR("hlb200: Intra_Base resampling of this enhancement-layer I picture is not implemented by the device path (fewer than %d macroblocks, cropping, chroma phases or field coding)", dyadic ? 36 : 64);
            return HL_ERROR_NOT_IMPLEMENTED;
        }
        level_idc = (int)top->pc_slice_hdr->pc_pps->pc_sps->level_idc;
    }
    L = &g_svc[li];
    if ((rc = glue_use_device(glue_layer_device(li)))) return glue_fail("hlb200_init", rc);
    if (!L->ctx || L->w != W || L->h != H || L->dev != g_dev_cur || L->codec != (const void*)p_codec) {
        if (L->ctx) { const int now = g_dev_cur; if (glue_use_device(L->dev) == 0) hlb200_stream_destroy(L->ctx); glue_use_device(now); L->ctx = NULL; free(L->motion); free(L->coeffs); free(L->rec); free(L->base); L->motion = NULL; L->coeffs = NULL; L->rec = NULL; L->base = NULL; L->nbase = 0; }
        if ((rc = hlb200_stream_create(W, H, 1, &L->ctx))) return glue_fail("hlb200_stream_create", rc);
        L->dev = g_dev_cur; L->codec = (const void*)p_codec;
        L->w = W; L->h = H; L->nmb = mbw * (H >> 4); L->fs_of_slot[0] = L->fs_of_slot[1] = NULL;
        L->motion = (hlb200_mb_motion_t*)calloc((size_t)L->nmb, sizeof(hlb200_mb_motion_t));
        L->coeffs = (hlb200_mb_coeffs_t*)calloc((size_t)L->nmb, sizeof(hlb200_mb_coeffs_t));
        L->rec = (uint8_t*)malloc(ysz + 2 * csz);
        if (!L->motion || !L->coeffs || !L->rec) {
            hlb200_stream_destroy(L->ctx); L->ctx = NULL; free(L->motion); free(L->coeffs); free(L->rec); L->motion = NULL; L->coeffs = NULL; L->rec = NULL;
            return HL_ERROR_OUTOFMEMMORY;
        }
    }
    /* (a) what the device needs besides the pictures.  I pictures: nothing (the resampling runs there).  P pictures: the reference layer's macroblock fields the
     * inter-layer derivation reads (utils.c:1701, mb.h:313-339, utils.c:1807-1834), copied verbatim from the reference layer's macroblock objects -- 84 bytes per
     * reference-layer macroblock; the derivation itself (utils.c:1225 + :1498 per macroblock) runs on the device (hlb_svc_derive.cuh).  The QP is the slice's: the
     * prologue of macroblock 0 computes it exactly as the reference's loop will (per-macroblock QP needs rate control, which glue_check_settings refuses). */
    {
        hl_codec_264_mb_t* p_mb;
        if ((err = glue_svc_loop_prologue(p_codec, p_esd, 0, &p_mb))) return err;
        qp = p_mb->QPy;
    }
    if (!intra) {
        const hl_codec_264_layer_t* rl = pc_layer->pc_ref;
        const hl_codec_264_layer_t* top = p_codec->layers.p_list[(p_codec->layers.currDQId >> 4) << 4];
        const uint32_t nref = (pc_layer->RefLayerPicWidthInSamplesL >> 4) * (pc_layer->RefLayerPicHeightInSamplesL >> 4);
        if (!rl || rl == pc_layer || !top || !top->pc_slice_hdr || rl->u_list_macroblocks_count < nref || hdr->field_pic_flag || hdr->MbaffFrameFlag || pc_layer->RefLayerMbaffFrameFlag ||
            pc_layer->RefLayerFieldPicFlag || !IsSliceHeaderEP(hdr)) {
            HL_DEBUG_ERROR("hlb200: the inter-layer motion derivation of this picture (field / MBAFF coding, B slices or a missing reference layer) is not implemented by the device path");
            return HL_ERROR_NOT_IMPLEMENTED;
        }
        if (L->nbase < (int)nref) {
            free(L->base);
            if (!(L->base = (hlb200_svc_base_mb_t*)calloc(nref, sizeof(hlb200_svc_base_mb_t)))) { L->nbase = 0; return HL_ERROR_OUTOFMEMMORY; }
            L->nbase = (int)nref;
        }
        for (addr = 0; addr < nref; ++addr) {
            const hl_codec_264_mb_t* b = rl->pp_list_macroblocks[addr];
            hlb200_svc_base_mb_t* o = &L->base[addr];
            int p, q;
            memset(o, 0, sizeof(*o));
            if (!b) { HL_DEBUG_ERROR("hlb200: reference-layer macroblock %u does not exist", addr); return HL_ERROR_INVALID_STATE; }
            o->flags = (uint8_t)(((HL_CODEC_264_MB_TYPE_IS_I_PCM(b) || HL_CODEC_264_MB_TYPE_IS_I_16X16(b) || HL_CODEC_264_MB_TYPE_IS_I_8X8(b) || HL_CODEC_264_MB_TYPE_IS_I_4X4(b) || HL_CODEC_264_MB_TYPE_IS_I_BL(b)) ? 1 : 0) |
                                 (HL_CODEC_264_MB_TYPE_IS_INTRA(b) ? 2 : 0) | ((b->e_type == HL_CODEC_264_MB_TYPE_P_8X8 || b->e_type == HL_CODEC_264_MB_TYPE_P_8X8REF0) ? 4 : 0));
            o->part_w = (uint8_t)b->MbPartWidth; o->part_h = (uint8_t)b->MbPartHeight;
            for (p = 0; p < 4; ++p) {
                o->sub_w[p] = (uint8_t)b->SubMbPartWidth[p]; o->sub_h[p] = (uint8_t)b->SubMbPartHeight[p];
                o->pred_flag[p] = (int8_t)b->predFlagL0[p]; o->ref_idx[p] = (int8_t)b->refIdxL0[p];
                for (q = 0; q < 4; ++q) { o->mv[p][q][0] = (int16_t)b->mvL0[p][q].x; o->mv[p][q][1] = (int16_t)b->mvL0[p][q].y; }
            }
        }
        geom.ref_width = (int32_t)pc_layer->RefLayerPicWidthInSamplesL; geom.ref_height = (int32_t)pc_layer->RefLayerPicHeightInSamplesL;
        geom.scaled_width = (int32_t)hdr->ext.svc.ScaledRefLayerPicWidthInSamplesL; geom.scaled_height = (int32_t)hdr->ext.svc.ScaledRefLayerPicHeightInSamplesL;
        geom.left_offset = (int32_t)hdr->ext.svc.ScaledRefLayerLeftOffset; geom.top_offset = (int32_t)hdr->ext.svc.ScaledRefLayerTopOffset;
        geom.level_idc = (int32_t)top->pc_slice_hdr->pc_pps->pc_sps->level_idc;
        geom.restricted = (int32_t)pc_layer->RestrictedSpatialResolutionChangeFlag; geom.cropping_change = (int32_t)pc_layer->CroppingChangeFlag;
    }
    /* (b) device: one call for the picture */
    if ((rc = hlb200_frame_upload(L->ctx, frame->data_ptr[0], frame->data_ptr[1], frame->data_ptr[2], W, Wc))) return glue_fail("hlb200_frame_upload", rc);
    /* the layer's two device slots hold its last reconstructions; the picture being coded goes to the one that does not hold the reference */
    {
        const void* ref_fs = intra ? NULL : (const void*)pc_layer->pobj_poc->RefPicList0[0];
        ref_slot = (ref_fs && L->fs_of_slot[1] == ref_fs) ? 1 : 0;
        cur_slot = ref_slot ^ 1;
        if (!intra) {
            const hl_codec_264_pict_t* ref = ref_fs ? pc_layer->pobj_poc->RefPicList0[0]->p_pict : NULL;
            if (!ref) return HL_ERROR_INVALID_STATE;
            /* normally the reference is the picture this context reconstructed last time and is still resident; otherwise (a reordered list, a picture this context did not
             * produce) it travels from the host DPB */
            if (L->fs_of_slot[ref_slot] != ref_fs) {
                if ((rc = hlb200_slot_upload(L->ctx, ref_slot, ref->pc_data_y, ref->pc_data_u, ref->pc_data_v))) return glue_fail("hlb200_slot_upload", rc);
                L->fs_of_slot[ref_slot] = ref_fs;
            }
        }
        if (L->fs_of_slot[ref_slot] == (const void*)pc_layer->pc_fs_curr) L->fs_of_slot[ref_slot] = NULL;   /* the host recycles the reference's frame store for this picture */
        L->fs_of_slot[cur_slot] = NULL;
    }
    if (intra) {
        const hl_codec_264_pict_t* rp = pc_layer->pc_ref->pc_fs_curr->p_pict;   /* what decode_svc.c:2970 reads: the reference layer's reconstruction of this access unit */
        /* that picture was reconstructed by the reference layer's own device context a moment ago and is still resident there (same samples as the host copy: the glue
         * itself wrote the host copy from it): hand it over device to device -- GPU to GPU when the layers live on different ones -- instead of uploading the host copy */
        const int li_ref = (int)(pc_layer->pc_ref->DQId >> 4);
        const void* rfs = (const void*)pc_layer->pc_ref->pc_fs_curr;
        hlb200_ctx_t* rctx = NULL;
        int rslot = -1, k;
        if (li_ref == 0) {
            const glue_stream_t* gb = glue_stream_of(p_codec, 0);
            for (k = 0; gb && gb->ctx && k < gb->nslots; ++k) if (gb->fs_of_slot[k] == rfs && gb->w == (int)pc_layer->RefLayerPicWidthInSamplesL && gb->h == (int)pc_layer->RefLayerPicHeightInSamplesL) { rctx = gb->ctx; rslot = k; }
        }
        else if (li_ref > 0 && li_ref < li && g_svc[li_ref].ctx && g_svc[li_ref].w == (int)pc_layer->RefLayerPicWidthInSamplesL && g_svc[li_ref].h == (int)pc_layer->RefLayerPicHeightInSamplesL) {
            for (k = 0; k < 2; ++k) if (g_svc[li_ref].fs_of_slot[k] == rfs) { rctx = g_svc[li_ref].ctx; rslot = k; }
        }
        if (rctx && !getenv("HLB200_SVC_HOST_HANDOFF"))
            rc = hlb200_svc_layer_picture_resampled_from(L->ctx, cur_slot, qp, hdr->pc_pps->chroma_qp_index_offset, rctx, rslot, level_idc, L->coeffs);
        else
        rc = hlb200_svc_layer_picture_resampled(L->ctx, cur_slot, qp, hdr->pc_pps->chroma_qp_index_offset, rp->pc_data_y, rp->pc_data_u, rp->pc_data_v,
                                                (int)pc_layer->RefLayerPicWidthInSamplesL, (int)pc_layer->RefLayerPicHeightInSamplesL, level_idc, L->coeffs);
    }
    else {
        rc = hlb200_svc_layer_picture_derived(L->ctx, ref_slot, cur_slot, qp, hdr->pc_pps->chroma_qp_index_offset, L->base, &geom, L->motion, &status, L->coeffs);
        if (rc == HLB200_ERR_NOT_IMPLEMENTED) {
            /* status bits (hlb200.h HLB200_SVC_DERIVE_*): partitions the fused kernel is not pinned for, a macroblock whose base macroblock is intra while its object still holds
             * partitions of an earlier picture, or one with no earlier macroblock of the picture to inherit a prediction from -- after a layer's I picture the reference codes
             * those against scratch memory of the I_BL function (rdo.c:1304-1312 / rdo.c:344-349 / hl_memory.h:226-241).  Not reproduced, never handed to the reference's CPU
             * function: the picture is refused. */
            HL_DEBUG_ERROR("hlb200: layer %d: the derived motion of this picture is outside what the device path reproduces (status bits %d): not implemented", li, (int)status);
            return HL_ERROR_NOT_IMPLEMENTED;
        }
    }
    if (rc) return glue_fail("hlb200_svc_layer_picture", rc);
    if ((rc = hlb200_slot_download(L->ctx, cur_slot, L->rec, L->rec + ysz, L->rec + ysz + csz))) return glue_fail("hlb200_slot_download", rc);
    L->fs_of_slot[cur_slot] = (const void*)pc_layer->pc_fs_curr;   /* the host's loop copies L->rec into this frame store (glue_svc_apply): same samples on both sides */
    /* the reference's own loop: same prologue again, the wrapped guess functions copy the device's results, the real writer serialises them */
    g_svc_active = L; g_svc_intra = intra;
    err = __real_hl_codec_264_nal_slice_data_encode(p_codec, p_esd);
    g_svc_active = NULL;
    return err;
}

static HL_ERROR_T glue_svc_apply(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int intra)
{
    glue_svc_layer_t* L = g_svc_active;
    const hl_codec_264_pict_t* pict = p_codec->layers.pc_active->pc_fs_curr->p_pict;
    const hlb200_mb_coeffs_t* c;
    const int W = L->w, Wc = W >> 1;
    const size_t ysz = (size_t)W * L->h, csz = ysz >> 2;
    HL_ERROR_T err;
    int b, i, k, y, i8;
    if (p_mb->u_addr >= (uint32_t)L->nmb) return HL_ERROR_INVALID_STATE;
    if ((err = glue_svc_derive(p_mb, p_codec, intra))) return err;   /* the syntax fields the writer serialises and later layers / pictures derive from */
    if (!intra) {
        /* the device coded the macroblock with ITS derivation of the same reference-layer fields: both must agree, or the stream would silently diverge */
        const hlb200_mb_motion_t* m = &L->motion[p_mb->u_addr];
        static const int N[4] = { 1, 2, 2, 4 }, PW[4] = { 16, 16, 8, 8 }, PH[4] = { 16, 8, 16, 8 }, SN[4] = { 1, 2, 2, 4 }, SW[4] = { 8, 8, 4, 4 }, SH[4] = { 8, 4, 8, 4 };
        int p, q, same;
        if (m->pad[0] & 1) same = p_mb->predFlagL0[0] == 0 && HL_CODEC_264_MB_TYPE_IS_I_BL(p_mb);
        else {
            same = m->part_mode < 4 && p_mb->NumMbPart == N[m->part_mode] && p_mb->MbPartWidth == PW[m->part_mode] && p_mb->MbPartHeight == PH[m->part_mode];
            for (p = 0; same && p < N[m->part_mode]; ++p) {
                const int sm = m->part_mode == 3 ? (m->sub_mode[p] & 3) : 0, ns = m->part_mode == 3 ? SN[sm] : 1;
                const int w = m->part_mode == 3 ? SW[sm] : PW[m->part_mode], h = m->part_mode == 3 ? SH[sm] : PH[m->part_mode];
                same = p_mb->predFlagL0[p] == 1 && p_mb->refIdxL0[p] == m->ref_idx[p] && p_mb->NumSubMbPart[p] == ns && p_mb->partWidth[p][0] == w && p_mb->partHeight[p][0] == h;
                for (q = 0; same && q < ns; ++q) same = p_mb->mvL0[p][q].x == m->mv[p][q][0] && p_mb->mvL0[p][q].y == m->mv[p][q][1];
            }
        }
        if (!same) {
            HL_DEBUG_ERROR("hlb200: macroblock %u: the motion derived on the device differs from the reference's derivation", p_mb->u_addr);
            return HL_ERROR_INVALID_STATE;
        }
    }
    c = &L->coeffs[p_mb->u_addr];
    p_mb->CodedBlockPatternLuma4x4 = c->cbp_luma4x4;
    for (b = 0; b < 16; ++b) for (i = 0; i < 16; ++i) p_mb->LumaLevel[b][i] = c->luma_level[b][i];
    for (k = 0; k < 2; ++k) {
        p_mb->CodedBlockPatternChromaDC4x4[k] = c->cbp_chroma_dc4x4[k];
        p_mb->CodedBlockPatternChromaAC4x4[k] = c->cbp_chroma_ac4x4[k];
        for (b = 0; b < 4; ++b) {
            p_mb->ChromaDCLevel[k][b] = c->chroma_dc_level[k][b];
            for (i = 0; i < 16; ++i) p_mb->ChromaACLevel[k][b][i] = c->chroma_ac_level[k][b][i];
        }
    }
    for (y = 0; y < 16; ++y) memcpy(pict->pc_data_y + (size_t)(p_mb->yL + y) * W + p_mb->xL, L->rec + (size_t)(p_mb->yL + y) * W + p_mb->xL, 16);
    for (y = 0; y < 8; ++y) {
        memcpy(pict->pc_data_u + (size_t)(p_mb->yL / 2 + y) * Wc + p_mb->xL / 2, L->rec + ysz + (size_t)(p_mb->yL / 2 + y) * Wc + p_mb->xL / 2, 8);
        memcpy(pict->pc_data_v + (size_t)(p_mb->yL / 2 + y) * Wc + p_mb->xL / 2, L->rec + ysz + csz + (size_t)(p_mb->yL / 2 + y) * Wc + p_mb->xL / 2, 8);
    }
    /* _hl_codec_264_rdo_mb_guess_cbp, rdo.c:2703-2781 (static there), for a macroblock that is not Intra16x16 */
    p_mb->CodedBlockPatternLuma = 0;
    for (i8 = 0; i8 < 4; ++i8) if ((p_mb->CodedBlockPatternLuma4x4 >> (i8 << 2)) & 15) p_mb->CodedBlockPatternLuma |= (1 << i8);
    if ((p_mb->CodedBlockPatternChromaDC4x4[0] || p_mb->CodedBlockPatternChromaDC4x4[1]) && !p_mb->CodedBlockPatternChromaAC4x4[0] && !p_mb->CodedBlockPatternChromaAC4x4[1])
        p_mb->CodedBlockPatternChroma = 1;
    else if (p_mb->CodedBlockPatternChromaAC4x4[0] || p_mb->CodedBlockPatternChromaAC4x4[1]) p_mb->CodedBlockPatternChroma = 2;
    else p_mb->CodedBlockPatternChroma = 0;
    p_mb->coded_block_pattern = (p_mb->CodedBlockPatternChroma << 4 | p_mb->CodedBlockPatternLuma);
    if (p_mb->coded_block_pattern > 47) { p_mb->coded_block_pattern -= 16; p_mb->CodedBlockPatternChroma = p_mb->coded_block_pattern >> 4; }
    return HL_ERROR_SUCCESS;
}

HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_inter_pred_svc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec)
{
    if (!g_svc_active) return HL_ERROR_INVALID_STATE;
    return glue_svc_apply(p_mb, p_codec, 0);
}
HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_intra_pred_svc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec)
{
    if (!g_svc_active) return HL_ERROR_INVALID_STATE;
    return glue_svc_apply(p_mb, p_codec, 1);
}

HL_ERROR_T __wrap_hl_codec_264_nal_slice_data_encode(hl_codec_264_t* p_codec, hl_codec_264_encode_slice_data_t* p_esd)
{
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    const hl_codec_264_nal_slice_header_t* hdr = p_esd->pc_slice->p_header;
    const hl_frame_video_t* frame = p_codec->encoder.pc_frame;
    const int W = (int)hdr->PicWidthInSamplesL, H = (int)hdr->PicHeightInSamplesL;
    hlb200_slice_params_t prm;
    glue_stream_t* g;
    int rc, u, s, cur = -1, bits_mode;
    HL_ERROR_T err;

    if ((err = glue_check_settings(p_codec))) return err;
    if (p_codec->layers.currDQId > 0) return glue_svc_slice(p_codec, p_esd);
#ifdef HLB200_GLUE_HOST_BASE_LAYER   /* test builds without a GPU: the base layer stays on the reference's CPU path, only the enhancement-layer hook is exercised */
    return __real_hl_codec_264_nal_slice_data_encode(p_codec, p_esd);
#endif
    if (p_esd->i_mb_start != 0 || p_esd->i_mb_end != (int32_t)hdr->PicSizeInMbs) {
        HL_DEBUG_ERROR("hlb200: only single-slice AVC pictures are supported by the device path");
        return HL_ERROR_NOT_IMPLEMENTED;
    }
    if (!(g = glue_stream_of(p_codec, 1))) { HL_DEBUG_ERROR("hlb200: too many codec instances"); return HL_ERROR_OUTOFCAPACITY; }
    if ((rc = glue_use_device(glue_layer_device(0)))) return glue_fail("hlb200_init", rc);
    if (!g->ctx || g->w != W || g->h != H) {
        int refs = (int)p_codec->pc_base->max_ref_frame;
        if (refs < 1) refs = 1;
        if (refs > HLB200_MAX_REFS) refs = HLB200_MAX_REFS;
        if (g->ctx) { hlb200_stream_destroy(g->ctx); g->ctx = NULL; free(g->rec); free(g->bits); g->rec = NULL; g->bits = NULL; }
        if ((rc = hlb200_stream_create(W, H, refs, &g->ctx))) return glue_fail("hlb200_stream_create", rc);
        g->w = W; g->h = H; g->nmb = (W >> 4) * (H >> 4); g->nslots = refs + 1;
        memset(g->fs_of_slot, 0, sizeof(g->fs_of_slot));
    }
    memset(&prm, 0, sizeof(prm));
    g->svc = p_codec->encoder.b_svc_enabled ? 1 : 0;
    g->is_p = IsSliceHeaderP(hdr) ? 1 : 0;
    /* Single-layer streams: the device also serialises slice_data() (hlb200_slice_bits_*), the reference's macroblock loop and CAVLC writer do not run.  Streams with
     * SVC enhancement layers keep the decision records: the layers' inter-layer derivation reads this layer's macroblock objects on the host (utils.c:966-2439). */
    bits_mode = !g->svc && !getenv("HLB200_GLUE_RECORDS");
    prm.slice_type = g->is_p;
    prm.qp = p_codec->encoder.i_qp;
    prm.me_range = (int32_t)p_codec->pc_base->me_range;
    prm.deblock_flag = p_codec->pc_base->deblock_flag ? 1 : 0;                /* loop filter over the finished picture, slice.c:1897 */
    prm.me_early_term_flag = p_codec->pc_base->me_early_term_flag ? 1 : 0;   /* homogeneous-block mode mask, rdo.c:889-935 (the one live part of the flag) */
    prm.chroma_qp_index_offset = hdr->pc_pps->chroma_qp_index_offset;
    prm.num_refs = g->is_p ? (int32_t)hdr->num_ref_idx_l0_active_minus1 + 1 : 0;
    /* device slots of the reference pictures (RefPicList0 order), then a free slot for the current picture */
    for (u = 0; u < prm.num_refs; ++u) {
        const void* fs = pc_layer->pobj_poc->RefPicList0[u];
        prm.ref_slot[u] = -1;
        for (s = 0; s < g->nslots; ++s) if (fs && g->fs_of_slot[s] == fs) prm.ref_slot[u] = s;
        if (prm.ref_slot[u] < 0) { HL_DEBUG_ERROR("hlb200: reference picture %d is not resident on the device", u); return HL_ERROR_INVALID_STATE; }
    }
    for (s = 0; s < g->nslots && cur < 0; ++s) {
        int used = 0;
        for (u = 0; u < prm.num_refs; ++u) used |= (prm.ref_slot[u] == s);
        if (!used) cur = s;
    }
    prm.cur_slot = cur;
    for (s = 0; s < g->nslots; ++s) if (g->fs_of_slot[s] == (const void*)pc_layer->pc_fs_curr) g->fs_of_slot[s] = NULL;
    g->fs_of_slot[cur] = pc_layer->pc_fs_curr;

    if ((rc = hlb200_frame_upload(g->ctx, frame->data_ptr[0], frame->data_ptr[1], frame->data_ptr[2], W, W >> 1))) return glue_fail("hlb200_frame_upload", rc);
    if (bits_mode) {
        uint32_t nbits = 0, k, nw;
        const size_t need = (size_t)g->nmb * HLB200_BITS_WORDS_PER_MB + 64;   /* capacity of the device buffer (hlb200.h) */
        if (!g->bits) { g->bits = (uint32_t*)malloc(sizeof(uint32_t) * need); g->bits_cap = need; if (!g->bits) return HL_ERROR_OUTOFMEMMORY; }
        if (g_batch.active) {
            /* submit, hand control back to the driver until every stream of the batch has submitted and ONE launch has encoded them all */
            if (g_batch.n >= GLUE_MAX_STREAMS) return HL_ERROR_OUTOFCAPACITY;
            g_batch.ctxs[g_batch.n] = g->ctx; g_batch.prm[g_batch.n] = prm; g_batch.types[g_batch.n] = prm.slice_type; ++g_batch.n;
            g_batch.yield(g_batch.yield_arg);
            if (g_batch.rc) return glue_fail("batch launch", g_batch.rc);
        }
        else {
            int32_t type = prm.slice_type;
            if ((rc = hlb200_slice_encode_async(g->ctx, &prm))) return glue_fail("hlb200_slice_encode_async", rc);
            if ((rc = hlb200_slice_bits_batch_async(&g->ctx, &type, 1))) return glue_fail("hlb200_slice_bits_batch_async", rc);
        }
        if ((rc = hlb200_slice_bits_download(g->ctx, g->bits, g->bits_cap, &nbits))) return glue_fail("hlb200_slice_bits_download", rc);
        /* slice_data() continues the slice header bit for bit: appended with the reference's own bit writer, then rbsp_trailing_bits() as slice.c:1931 */
        nw = nbits >> 5;
        for (k = 0; k < nw; ++k) hl_codec_264_bits_write_u(p_esd->pobj_bits, g->bits[k], 32);
        if (nbits & 31) hl_codec_264_bits_write_u(p_esd->pobj_bits, g->bits[nw] >> (32 - (nbits & 31)), (nbits & 31));
        pc_layer->i_mb_encode_count += g->nmb;
        if (getenv("HLB200_SYNC_RECON")) {
            const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
            if ((rc = hlb200_slot_download(g->ctx, cur, (uint8_t*)pict->pc_data_y, (uint8_t*)pict->pc_data_u, (uint8_t*)pict->pc_data_v))) return glue_fail("hlb200_slot_download", rc);
        }
        return hl_codec_264_rbsp_avc_trailing_bits_write(p_esd->pobj_bits);
    }
    if (!g->rec && !(g->rec = (hlb200_mb_record_t*)malloc(sizeof(hlb200_mb_record_t) * (size_t)g->nmb))) return HL_ERROR_OUTOFMEMMORY;
    if ((rc = hlb200_slice_encode(g->ctx, &prm, g->rec))) return glue_fail("hlb200_slice_encode", rc);
    if (getenv("HLB200_SYNC_RECON") || p_codec->encoder.b_svc_enabled) {   /* SVC: the enhancement layers resample / derive from the base picture on the host */
        const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
        if ((rc = hlb200_slot_download(g->ctx, cur, (uint8_t*)pict->pc_data_y, (uint8_t*)pict->pc_data_u, (uint8_t*)pict->pc_data_v))) return glue_fail("hlb200_slot_download", rc);
    }
    g_cur = g;
    err = __real_hl_codec_264_nal_slice_data_encode(p_codec, p_esd);
    g_cur = NULL;
    return err;
}

HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_inter_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    (void)p_codec;
    if (!g_cur || !g_cur->rec || p_mb->u_addr >= (uint32_t)g_cur->nmb) return HL_ERROR_INVALID_STATE;
    glue_apply(p_mb, &g_cur->rec[p_mb->u_addr], pi_mad);
    return HL_ERROR_SUCCESS;
}

HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_intra_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    (void)p_codec;
    if (!g_cur || !g_cur->rec || p_mb->u_addr >= (uint32_t)g_cur->nmb) return HL_ERROR_INVALID_STATE;
    glue_apply(p_mb, &g_cur->rec[p_mb->u_addr], pi_mad);
    return HL_ERROR_SUCCESS;
}
