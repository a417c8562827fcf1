/*
 * ref_driver.c -- drives the UNMODIFIED reference encoder (hartallo C path) on synthetic YUV 4:2:0 and records what
 * a bit-exact re-implementation must reproduce.  TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product.
 *
 * Shape mirrors the reference's own integration harness, source/test_encoder.c:78-235 (engine init, plugin find,
 * codec create, public fields, one hl_codec_encode per frame, HDR/DATA result handling).  Decision records are taken
 * with link-time wrappers (-Wl,--wrap=...), not by editing reference files:
 *   hl_codec_264_nal_slice_data_encode              source/h264/hl_codec_264_slice.c:1701  (per frame: recon planes)
 *   hl_codec_264_rdo_mb_guess_best_inter_pred_avc   source/h264/hl_codec_264_rdo.c:678     (per MB decision record)
 *   hl_codec_264_rdo_mb_guess_best_intra_pred_avc   source/h264/hl_codec_264_rdo.c:99      (per MB, I slices)
 *   hl_codec_264_me_ds_mb_find_best_cost            source/h264/hl_codec_264_me_ds.c:104   (per mode search result)
 *   hl_codec_264_interpol_luma                      source/h264/hl_codec_264_pred_inter.c:339 (per candidate, opt.)
 *   hl_codec_264_residual_write_block_cavlc         source/h264/hl_codec_264_residual.c:587   (per block bits, opt.)
 *
 * Trace file = stream of int32 records: [tag, nwords, payload...].  Layout documented at each emit_* below and
 * mirrored by tests/reftrace.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

#include "hartallo/hl_api.h"
#include "hartallo/hl_codec.h"
#include "hartallo/hl_frame.h"
#include "hartallo/hl_debug.h"
#include "hartallo/hl_math.h"
#include "hartallo/hl_md5.h"
#include "hartallo/hl_memory.h"
#include "hartallo/h264/hl_codec_264.h"
#include "hartallo/h264/hl_codec_264_mb.h"
#include "hartallo/h264/hl_codec_264_layer.h"
#include "hartallo/h264/hl_codec_264_encode.h"
#include "hartallo/h264/hl_codec_264_slice.h"
#include "hartallo/h264/hl_codec_264_pict.h"
#include "hartallo/h264/hl_codec_264_dpb.h"
#include "hartallo/h264/hl_codec_264_bits.h"
#include "hartallo/h264/hl_codec_264_residual.h"
#include "hartallo/h264/hl_codec_264_me.h"
#include "hartallo/h264/hl_codec_264_me_ds.h"
#include "hartallo/h264/hl_codec_264_macros.h"
#include "hartallo/h264/hl_codec_264_pps.h"
#include "hartallo/h264/hl_codec_264_sps.h"

static FILE* g_trace = NULL;       /* int32 record stream */
static FILE* g_recon = NULL;       /* raw recon planes, frame after frame */
static int g_trace_cand = 0;       /* per-candidate records */
static int g_trace_levels = 1;
static int g_trace_state = 0;      /* per-MB post-writer state records (tag 5) */     /* include level arrays in MB records */
static int g_frame_idx = -1;       /* index of the frame being encoded */
static int g_width = 0, g_height = 0;

/* ------------------------------------------------------------------------------------------------------------ */
/* candidate accumulator                                                                                         */
static struct {
    int active, addr, mode, part, sub, mvx, mvy, dist, bits, nz_blocks, single_ctr;
} g_cand;
static int g_cur_mode = -1;
static int32_t (*g_real_sad)(const uint8_t*, int32_t, const uint8_t*, int32_t) = NULL;

static void put32(const int32_t* v, size_t n) { if (g_trace) fwrite(v, sizeof(int32_t), n, g_trace); }

static void flush_cand(void)
{
    if (g_cand.active && g_trace && g_trace_cand) {
        /* tag 4: [4, 13, frame, addr, mode, part, sub, mvx, mvy, dist, bits, nz_blocks, single_ctr] */
        int32_t r[13] = { 4, 13, g_frame_idx, g_cand.addr, g_cand.mode, g_cand.part, g_cand.sub, g_cand.mvx, g_cand.mvy,
                          g_cand.dist, g_cand.bits, g_cand.nz_blocks, g_cand.single_ctr };
        put32(r, 13);
    }
    g_cand.active = 0;
}

static int32_t traced_sad(const uint8_t* b1, int32_t s1, const uint8_t* b2, int32_t s2)
{
    int32_t v = g_real_sad(b1, s1, b2, s2);
    if (g_cand.active) g_cand.dist += v;
    return v;
}

#ifndef HL_DRIVER_NO_WRAPS   /* tracing wrappers (oracle); the glue build (host/hlb200_glue.c) provides its own hooks instead */
extern HL_ERROR_T __real_hl_codec_264_interpol_luma(hl_codec_264_t*, hl_codec_264_mb_t*, int32_t, int32_t,
        const hl_codec_264_mv_xt*, const hl_pixel_t*, void*, int32_t);
HL_ERROR_T __wrap_hl_codec_264_interpol_luma(hl_codec_264_t* p_codec, hl_codec_264_mb_t* p_mb, int32_t mbPartIdx,
        int32_t subMbPartIdx, const hl_codec_264_mv_xt* mvLX, const hl_pixel_t* cSL, void* pred, int32_t sampleSize)
{
    if (g_trace_cand && g_cur_mode >= 0 && sampleSize == 1) {
        flush_cand();
        g_cand.active = 1; g_cand.addr = (int)p_mb->u_addr; g_cand.mode = g_cur_mode;
        g_cand.part = mbPartIdx; g_cand.sub = subMbPartIdx; g_cand.mvx = mvLX->x; g_cand.mvy = mvLX->y;
        g_cand.dist = g_cand.bits = g_cand.nz_blocks = g_cand.single_ctr = 0;
    }
    return __real_hl_codec_264_interpol_luma(p_codec, p_mb, mbPartIdx, subMbPartIdx, mvLX, cSL, pred, sampleSize);
}

extern HL_ERROR_T __real_hl_codec_264_residual_write_block_cavlc(struct hl_codec_264_residual_inv_xs*,
        const struct hl_codec_264_s*, struct hl_codec_264_mb_s*, struct hl_codec_264_bits_s*, int32_t*, int32_t, int32_t, int32_t);
HL_ERROR_T __wrap_hl_codec_264_residual_write_block_cavlc(struct hl_codec_264_residual_inv_xs* p_inv,
        const struct hl_codec_264_s* pc_codec, struct hl_codec_264_mb_s* p_mb, struct hl_codec_264_bits_s* p_bits,
        int32_t coeffLevel[16], int32_t startIdx, int32_t endIdx, int32_t maxNumCoef)
{
    int32_t before = (int32_t)hl_codec_264_bits_get_stream_index(p_bits);
    HL_ERROR_T err = __real_hl_codec_264_residual_write_block_cavlc(p_inv, pc_codec, p_mb, p_bits, coeffLevel, startIdx, endIdx, maxNumCoef);
    if (g_cand.active) {
        hl_codec_264_encode_slice_data_t* pc_esd = pc_codec->layers.pc_active->encoder.p_list_esd[p_mb->u_slice_idx];
        g_cand.bits += (int32_t)hl_codec_264_bits_get_stream_index(p_bits) - before;
        g_cand.nz_blocks += 1;
        g_cand.single_ctr += pc_esd->rdo.Single_ctr;
    }
    return err;
}

/* ------------------------------------------------------------------------------------------------------------ */
extern HL_ERROR_T __real_hl_codec_264_me_ds_mb_find_best_cost(struct hl_codec_264_mb_s*, struct hl_codec_264_s*, const hl_codec_264_me_part_xt*);
HL_ERROR_T __wrap_hl_codec_264_me_ds_mb_find_best_cost(struct hl_codec_264_mb_s* p_mb, struct hl_codec_264_s* p_codec, const hl_codec_264_me_part_xt* pc_part)
{
    HL_ERROR_T err;
    g_cur_mode = (int)pc_part->Mode;
    err = __real_hl_codec_264_me_ds_mb_find_best_cost(p_mb, p_codec, pc_part);
    flush_cand();
    g_cur_mode = -1;
    if (g_trace) {
        /* tag 2: [2, n, frame, addr, Mode, refIdxLX, b_probably_pskip, NumMbPart, NumSubMbPart[4],
         *         mvBest[16][2], mvpLX[16][2], i_best_dist[16], i_Single_ctr[16], cbp4x4[16], d_best_cost[16] (double = 2 words),
         *         TotalCoeffsLuma[16] (after the call)] */
        hl_codec_264_encode_slice_data_t* pc_esd = p_codec->layers.pc_active->encoder.p_list_esd[p_mb->u_slice_idx];
        int32_t r[12 + 32 + 32 + 16 + 16 + 16 + 32 + 16];
        int i, j, k = 0;
        r[k++] = 2; r[k++] = (int32_t)(sizeof(r) / sizeof(r[0])); r[k++] = g_frame_idx; r[k++] = (int32_t)p_mb->u_addr;
        r[k++] = (int32_t)pc_part->Mode; r[k++] = pc_esd->rdo.me.refIdxLX; r[k++] = pc_esd->rdo.me.b_probably_pskip ? 1 : 0;
        r[k++] = p_mb->NumMbPart;
        for (i = 0; i < 4; ++i) r[k++] = p_mb->NumSubMbPart[i];
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = pc_esd->rdo.me.mvBest[i][j].x; r[k++] = pc_esd->rdo.me.mvBest[i][j].y; }
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = pc_esd->rdo.me.mvpLX[i][j].x; r[k++] = pc_esd->rdo.me.mvpLX[i][j].y; }
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) r[k++] = pc_esd->rdo.me.i_best_dist[i][j];
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) r[k++] = pc_esd->rdo.me.i_Single_ctr[i][j];
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) r[k++] = pc_esd->rdo.me.i_best_CodedBlockPatternLuma4x4[i][j];
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { memcpy(&r[k], &pc_esd->rdo.me.d_best_cost[i][j], 8); k += 2; }
        for (i = 0; i < 16; ++i) r[k++] = p_mb->TotalCoeffsLuma[i];
        put32(r, (size_t)k);
    }
    return err;
}

#define MB_REC_HDR 188
static void emit_mb(struct hl_codec_264_mb_s* p_mb, struct hl_codec_264_s* p_codec, int which, int32_t mad, int err)
{
    /* tag 1: see field order below (MB_REC_HDR header words, then optional levels: LumaLevel[256], ChromaDCLevel[2][4],
     * ChromaACLevel[2][4][16], Intra16x16DCLevel[16], Intra16x16ACLevel[256]) */
    static int32_t r[MB_REC_HDR + 256 + 8 + 128 + 16 + 256];
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    hl_codec_264_encode_slice_data_t* pc_esd = pc_layer->encoder.p_list_esd[p_mb->u_slice_idx];
    int i, j, k = 0, n;
    if (!g_trace) return;
    r[k++] = 1; r[k++] = 0; r[k++] = g_frame_idx; r[k++] = (int32_t)p_mb->u_addr;
    r[k++] = IsSliceHeaderP(pc_esd->pc_slice->p_header) ? 1 : 0; r[k++] = which;
    r[k++] = (int32_t)p_mb->e_type; r[k++] = (int32_t)p_mb->mb_type; r[k++] = (int32_t)p_mb->flags_type;
    r[k++] = p_mb->NumMbPart; r[k++] = p_mb->MbPartWidth; r[k++] = p_mb->MbPartHeight;             /* 9..11 */
    for (i = 0; i < 4; ++i) r[k++] = p_mb->NumSubMbPart[i];                                         /* 12 */
    for (i = 0; i < 4; ++i) r[k++] = (int32_t)p_mb->sub_mb_type[i];                                 /* 16 */
    for (i = 0; i < 4; ++i) r[k++] = p_mb->SubMbPartWidth[i];                                       /* 20 */
    for (i = 0; i < 4; ++i) r[k++] = p_mb->SubMbPartHeight[i];                                      /* 24 */
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = p_mb->mvL0[i][j].x; r[k++] = p_mb->mvL0[i][j].y; }      /* 28 */
    for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = p_mb->mvd_l0[i][j].x; r[k++] = p_mb->mvd_l0[i][j].y; }  /* 60 */
    for (i = 0; i < 4; ++i) r[k++] = p_mb->RefIdxL0[i];                                             /* 92 */
    for (i = 0; i < 4; ++i) r[k++] = p_mb->PredFlagL0[i];                                           /* 96 */
    r[k++] = (int32_t)p_mb->coded_block_pattern; r[k++] = (int32_t)p_mb->CodedBlockPatternLuma;     /* 100,101 */
    r[k++] = (int32_t)p_mb->CodedBlockPatternChroma; r[k++] = (int32_t)p_mb->CodedBlockPatternLuma4x4; /* 102,103 */
    r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[1]; /* 104 */
    r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[1]; /* 106 */
    r[k++] = (int32_t)p_mb->Intra16x16PredMode;                                                      /* 108 */
    for (i = 0; i < 16; ++i) r[k++] = (int32_t)p_mb->Intra4x4PredMode[i];                           /* 109 */
    r[k++] = (int32_t)p_mb->intra_chroma_pred_mode;                                                  /* 125 */
    for (i = 0; i < 16; ++i) r[k++] = p_mb->prev_intra4x4_pred_mode_flag[i];                        /* 126 */
    for (i = 0; i < 16; ++i) r[k++] = p_mb->rem_intra4x4_pred_mode[i];                              /* 142 */
    r[k++] = p_mb->mb_qp_delta; r[k++] = p_mb->QPy; r[k++] = p_mb->QPc[0]; r[k++] = p_mb->QPc[1];   /* 158..161 */
    r[k++] = mad; r[k++] = err;                                                                      /* 162,163 */
    for (i = 0; i < 16; ++i) r[k++] = p_mb->TotalCoeffsLuma[i];                                     /* 164 */
    for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = p_mb->TotalCoeffsChromaACCbCr[i][j];   /* 180 */
    if (k != MB_REC_HDR) { fprintf(stderr, "emit_mb layout error %d\n", k); exit(3); }
    if (g_trace_levels) {
        for (i = 0; i < 16; ++i) for (j = 0; j < 16; ++j) r[k++] = p_mb->LumaLevel[i][j];
        for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = p_mb->ChromaDCLevel[i][j];
        for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = p_mb->ChromaACLevel[i][n][j];
        for (i = 0; i < 16; ++i) r[k++] = p_mb->Intra16x16DCLevel[i];
        for (i = 0; i < 16; ++i) for (j = 0; j < 16; ++j) r[k++] = p_mb->Intra16x16ACLevel[i][j];
    }
    r[1] = k;
    put32(r, (size_t)k);
}

extern HL_ERROR_T __real_hl_codec_264_rdo_mb_guess_best_inter_pred_avc(hl_codec_264_mb_t*, hl_codec_264_t*, int32_t*);
HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_inter_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    int32_t mad = 0;
    HL_ERROR_T err = __real_hl_codec_264_rdo_mb_guess_best_inter_pred_avc(p_mb, p_codec, &mad);
    if (pi_mad) *pi_mad = mad;
    emit_mb(p_mb, p_codec, 0, mad, (int)err);
    return err;
}

extern HL_ERROR_T __real_hl_codec_264_rdo_mb_guess_best_intra_pred_avc(hl_codec_264_mb_t*, hl_codec_264_t*, int32_t*);
HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_intra_pred_avc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec, int32_t* pi_mad)
{
    int32_t mad = 0;
    HL_ERROR_T err = __real_hl_codec_264_rdo_mb_guess_best_intra_pred_avc(p_mb, p_codec, &mad);
    if (pi_mad) *pi_mad = mad;
    emit_mb(p_mb, p_codec, 1, mad, (int)err);
    return err;
}

extern HL_ERROR_T __real_hl_codec_264_nal_slice_data_encode(hl_codec_264_t*, hl_codec_264_encode_slice_data_t*);
HL_ERROR_T __wrap_hl_codec_264_nal_slice_data_encode(hl_codec_264_t* p_codec, hl_codec_264_encode_slice_data_t* p_esd)
{
    HL_ERROR_T err = __real_hl_codec_264_nal_slice_data_encode(p_codec, p_esd);
    hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
    const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
    if (g_trace) {
        /* tag 3: [3, 10, frame, W, H, is_p, mb_start, mb_end, qp, lambda_mode (double, 2 words)] */
        int32_t r[10] = { 3, 10, g_frame_idx, (int32_t)pict->uWidthL, (int32_t)pict->uHeightL,
                          IsSliceHeaderP(p_esd->pc_slice->p_header) ? 1 : 0, p_esd->i_mb_start, p_esd->i_mb_end, p_esd->i_qp, 0 };
        double l = p_codec->encoder.rdo.d_lambda_mode; float lf = (float)l; memcpy(&r[9], &lf, 4);
        put32(r, 10);
    }
    if (g_trace && g_trace_state) {
        /* tag 5: per-MB state AFTER the writer (what the next MBs / the next frame see):
         * [5, n, frame, addr, e_type, flags_type, NumMbPart, MbPartWidth, MbPartHeight, NumSubMbPart[4], SubMbPartWidth[4], SubMbPartHeight[4],
         *  CodedBlockPatternLuma, CodedBlockPatternChroma, RefIdxL0[4], predFlagL0[4], MvL0[16][2], TotalCoeffsLuma[16], TotalCoeffsChromaACCbCr[2][4],
         *  Intra4x4PredMode[16], ChromaACLevel[2][4][16], last rdo.Single_ctr] */
        int a;
        for (a = p_esd->i_mb_start; a < p_esd->i_mb_end; ++a) {
            const hl_codec_264_mb_t* m = pc_layer->pp_list_macroblocks[a];
            int32_t r[512]; int k = 0, i, j, n;
            if (!m) continue;
            r[k++] = 5; r[k++] = 0; r[k++] = g_frame_idx; r[k++] = a; r[k++] = (int32_t)m->e_type; r[k++] = (int32_t)m->flags_type;
            r[k++] = m->NumMbPart; r[k++] = m->MbPartWidth; r[k++] = m->MbPartHeight;
            for (i = 0; i < 4; ++i) r[k++] = m->NumSubMbPart[i];
            for (i = 0; i < 4; ++i) r[k++] = m->SubMbPartWidth[i];
            for (i = 0; i < 4; ++i) r[k++] = m->SubMbPartHeight[i];
            r[k++] = (int32_t)m->CodedBlockPatternLuma; r[k++] = (int32_t)m->CodedBlockPatternChroma;
            for (i = 0; i < 4; ++i) r[k++] = m->RefIdxL0[i];
            for (i = 0; i < 4; ++i) r[k++] = m->predFlagL0[i];
            for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = m->MvL0[i][j].x; r[k++] = m->MvL0[i][j].y; }
            for (i = 0; i < 16; ++i) r[k++] = m->TotalCoeffsLuma[i];
            for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = m->TotalCoeffsChromaACCbCr[i][j];
            for (i = 0; i < 16; ++i) r[k++] = (int32_t)m->Intra4x4PredMode[i];
            for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = m->ChromaACLevel[i][n][j];
            r[k++] = p_esd->rdo.Single_ctr;
            r[1] = k;
            put32(r, (size_t)k);
        }
    }
    if (g_recon) {
        fwrite(pict->pc_data_y, 1, (size_t)pict->uWidthL * pict->uHeightL, g_recon);
        fwrite(pict->pc_data_u, 1, (size_t)pict->uWidthC * pict->uHeightC, g_recon);
        fwrite(pict->pc_data_v, 1, (size_t)pict->uWidthC * pict->uHeightC, g_recon);
    }
    return err;
}


/* SVC enhancement-layer inter macroblocks: hl_codec_264_rdo_mb_guess_best_inter_pred_svc, rdo.c:1273-1521.
 * tag 7 (once per layer picture, before its first tag 6): [7, n, frame, DQId, W, H, then W*H*3/2 source bytes and W*H*3/2 bytes of RefPicList0[0], 4 per word]
 * tag 6 (per macroblock, after the call): [6, n, frame, DQId, addr, QPy, QPc[2], NumMbPart, MbPartWidth, MbPartHeight, NumSubMbPart[4], SubMbPartWidth[4],
 *         SubMbPartHeight[4], predFlagL0[4], refIdxL0[4], mvL0[4][4][2], CodedBlockPatternLuma4x4, CodedBlockPatternChromaDC4x4[2], CodedBlockPatternChromaAC4x4[2],
 *         CodedBlockPatternLuma, CodedBlockPatternChroma, LumaLevel[16][16], ChromaDCLevel[2][4], ChromaACLevel[2][4][16], reconstructed Y 16x16, Cb 8x8, Cr 8x8 (one sample per word),
 *         ChromaACLevel[2][4][16] as it was BEFORE the call, e_type, partWidth/partHeight[4][0],
 *         ChromaDCLevel[2][4] as it was BEFORE the call] */
static int g_svc_last_frame = -1, g_svc_last_dqid = -1;
static void put_planes(const uint8_t* y, const uint8_t* u, const uint8_t* v, int W, int H)
{
    fwrite(y, 1, (size_t)W * H, g_trace); fwrite(u, 1, (size_t)W * H / 4, g_trace); fwrite(v, 1, (size_t)W * H / 4, g_trace);
}
extern HL_ERROR_T __real_hl_codec_264_rdo_mb_guess_best_inter_pred_svc(hl_codec_264_mb_t*, hl_codec_264_t*);
HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_inter_pred_svc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec)
{
    int32_t dc_in[2][4];
    int32_t ac_in[2][4][16];   /* ChromaACLevel persists in the macroblock object from picture to picture and is read again when a block's residual is zero (transf.c:236-245) */
    HL_ERROR_T err;
    { int c_, b_, i_; for (c_ = 0; c_ < 2; ++c_) for (b_ = 0; b_ < 4; ++b_) for (i_ = 0; i_ < 16; ++i_) ac_in[c_][b_][i_] = p_mb->ChromaACLevel[c_][b_][i_];
      for (c_ = 0; c_ < 2; ++c_) for (b_ = 0; b_ < 4; ++b_) dc_in[c_][b_] = p_mb->ChromaDCLevel[c_][b_]; }
    err = __real_hl_codec_264_rdo_mb_guess_best_inter_pred_svc(p_mb, p_codec);
    if (g_trace && !err) {
        hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
        const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
        const int W = (int)pict->uWidthL, H = (int)pict->uHeightL, dq = (int)p_codec->layers.currDQId;
        int32_t r[1200]; int k = 0, i, j, n, x, y;
        if (g_svc_last_frame != g_frame_idx || g_svc_last_dqid != dq) {
            const hl_codec_264_dpb_fs_t* fs = pc_layer->pobj_poc->RefPicList0[0];
            const hl_frame_video_t* in = p_codec->encoder.pc_frame;
            int32_t h7[6] = { 7, 6 + 2 * (W * H * 3 / 2) / 4, g_frame_idx, dq, W, H };
            put32(h7, 6);
            put_planes((const uint8_t*)in->data_ptr[0], (const uint8_t*)in->data_ptr[1], (const uint8_t*)in->data_ptr[2], W, H);
            put_planes(fs->p_pict->pc_data_y, fs->p_pict->pc_data_u, fs->p_pict->pc_data_v, W, H);
            g_svc_last_frame = g_frame_idx; g_svc_last_dqid = dq;
            /* tag 11 (once per enhancement-layer P picture, after its tag 7): what the inter-layer motion derivation (utils.c:966-2439) reads --
             * [11, n, frame, DQId, RefLayerPicWidthInSamplesL, RefLayerPicHeightInSamplesL, ScaledRefLayerPicWidthInSamplesL, ScaledRefLayerPicHeightInSamplesL,
             *  ScaledRefLayerLeftOffset, ScaledRefLayerTopOffset, level_idc (utils.c:989), RestrictedSpatialResolutionChangeFlag, CroppingChangeFlag,
             *  SpatialResolutionChangeFlag, number of reference-layer macroblocks, then per reference-layer macroblock 53 words: intra by e_type (utils.c:1701),
             *  HL_CODEC_264_MB_TYPE_IS_INTRA (flags_type, mb.h:322), e_type is P_8X8 / P_8X8REF0 (mb.h:329), MbPartWidth, MbPartHeight, SubMbPartWidth[4],
             *  SubMbPartHeight[4], predFlagL0[4], refIdxL0[4], mvL0[4][4][2]]; the reference layer's macroblock objects do not change while this picture is coded */
            if (pc_layer->pc_ref && pc_layer->pc_slice_hdr) {
                const hl_codec_264_layer_t* rl = pc_layer->pc_ref;
                const hl_codec_264_nal_slice_header_t* sh = pc_layer->pc_slice_hdr;
                const int nref = (int)rl->u_list_macroblocks_count;
                int32_t h11[15]; int a;
                h11[0] = 11; h11[1] = 15 + 53 * nref; h11[2] = g_frame_idx; h11[3] = dq;
                h11[4] = (int32_t)pc_layer->RefLayerPicWidthInSamplesL; h11[5] = (int32_t)pc_layer->RefLayerPicHeightInSamplesL;
                h11[6] = (int32_t)sh->ext.svc.ScaledRefLayerPicWidthInSamplesL; h11[7] = (int32_t)sh->ext.svc.ScaledRefLayerPicHeightInSamplesL;
                h11[8] = (int32_t)sh->ext.svc.ScaledRefLayerLeftOffset; h11[9] = (int32_t)sh->ext.svc.ScaledRefLayerTopOffset;
                h11[10] = (int32_t)p_codec->layers.p_list[(p_codec->layers.currDQId >> 4) << 4]->pc_slice_hdr->pc_pps->pc_sps->level_idc;
                h11[11] = (int32_t)pc_layer->RestrictedSpatialResolutionChangeFlag; h11[12] = (int32_t)pc_layer->CroppingChangeFlag;
                h11[13] = (int32_t)pc_layer->SpatialResolutionChangeFlag; h11[14] = nref;
                put32(h11, 15);
                for (a = 0; a < nref; ++a) {
                    const hl_codec_264_mb_t* b = rl->pp_list_macroblocks[a];
                    int32_t w[53]; int q = 0;
                    memset(w, 0, sizeof(w));
                    if (b) {
                        w[q++] = (HL_CODEC_264_MB_TYPE_IS_I_PCM(b) || HL_CODEC_264_MB_TYPE_IS_I_16X16(b) || HL_CODEC_264_MB_TYPE_IS_I_8X8(b) || HL_CODEC_264_MB_TYPE_IS_I_4X4(b) || HL_CODEC_264_MB_TYPE_IS_I_BL(b)) ? 1 : 0;
                        w[q++] = HL_CODEC_264_MB_TYPE_IS_INTRA(b) ? 1 : 0;
                        w[q++] = (b->e_type == HL_CODEC_264_MB_TYPE_P_8X8 || b->e_type == HL_CODEC_264_MB_TYPE_P_8X8REF0) ? 1 : 0;
                        w[q++] = b->MbPartWidth; w[q++] = b->MbPartHeight;
                        for (i = 0; i < 4; ++i) w[q++] = b->SubMbPartWidth[i];
                        for (i = 0; i < 4; ++i) w[q++] = b->SubMbPartHeight[i];
                        for (i = 0; i < 4; ++i) w[q++] = b->predFlagL0[i];
                        for (i = 0; i < 4; ++i) w[q++] = b->refIdxL0[i];
                        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { w[q++] = b->mvL0[i][j].x; w[q++] = b->mvL0[i][j].y; }
                    }
                    put32(w, 53);
                }
            }
        }
        r[k++] = 6; r[k++] = 0; r[k++] = g_frame_idx; r[k++] = dq; r[k++] = (int32_t)p_mb->u_addr; r[k++] = p_mb->QPy; r[k++] = p_mb->QPc[0]; r[k++] = p_mb->QPc[1];
        r[k++] = p_mb->NumMbPart; r[k++] = p_mb->MbPartWidth; r[k++] = p_mb->MbPartHeight;
        for (i = 0; i < 4; ++i) r[k++] = p_mb->NumSubMbPart[i];
        for (i = 0; i < 4; ++i) r[k++] = p_mb->SubMbPartWidth[i];
        for (i = 0; i < 4; ++i) r[k++] = p_mb->SubMbPartHeight[i];
        for (i = 0; i < 4; ++i) r[k++] = p_mb->predFlagL0[i];
        for (i = 0; i < 4; ++i) r[k++] = p_mb->refIdxL0[i];
        for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) { r[k++] = p_mb->mvL0[i][j].x; r[k++] = p_mb->mvL0[i][j].y; }
        r[k++] = (int32_t)p_mb->CodedBlockPatternLuma4x4; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[1];
        r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[1];
        r[k++] = (int32_t)p_mb->CodedBlockPatternLuma; r[k++] = (int32_t)p_mb->CodedBlockPatternChroma;
        for (i = 0; i < 16; ++i) for (j = 0; j < 16; ++j) r[k++] = p_mb->LumaLevel[i][j];
        for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = p_mb->ChromaDCLevel[i][j];
        for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = p_mb->ChromaACLevel[i][n][j];
        for (y = 0; y < 16; ++y) for (x = 0; x < 16; ++x) r[k++] = pict->pc_data_y[(p_mb->yL + y) * W + p_mb->xL + x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pict->pc_data_u[(p_mb->yL / 2 + y) * (W / 2) + p_mb->xL / 2 + x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pict->pc_data_v[(p_mb->yL / 2 + y) * (W / 2) + p_mb->xL / 2 + x];
        for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = ac_in[i][n][j];
        r[k++] = (int32_t)p_mb->e_type;
        for (i = 0; i < 4; ++i) { r[k++] = p_mb->partWidth[i][0]; r[k++] = p_mb->partHeight[i][0]; }
        for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = dc_in[i][j];
        r[1] = k;
        put32(r, (size_t)k);
        /* tag 10: a macroblock without any partition (base macroblock intra inside a P picture) is coded against whatever the scratch blocks of the
         * function's predMbL / predMbCb / predMbCr hold (4th, 5th, 6th block mapped: rdo.c:1309-1311; the allocator hands out blocks 31, 30, ... in order,
         * hl_memory.h:226-241) -- the function does not write them for such a macroblock, so they still hold what it used:
         * [10, n, frame, DQId, addr, Y 16x16, Cb 8x8, Cr 8x8] */
        if (p_mb->NumSubMbPart[0] == 0) {
            const int32_t* mem = pc_layer->encoder.p_list_esd[p_mb->u_slice_idx]->pc_mem_blocks->p_memory;
            const int32_t *bl = mem + (28 << 8), *bcb = mem + (27 << 8), *bcr = mem + (26 << 8);
            k = 0;
            r[k++] = 10; r[k++] = 0; r[k++] = g_frame_idx; r[k++] = dq; r[k++] = (int32_t)p_mb->u_addr;
            for (y = 0; y < 16; ++y) for (x = 0; x < 16; ++x) r[k++] = bl[y * 16 + x];
            for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = bcb[y * 16 + x];
            for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = bcr[y * 16 + x];
            r[1] = k;
            put32(r, (size_t)k);
        }
    }
    return err;
}

/* SVC enhancement-layer I_BL macroblocks (enhancement I pictures): hl_codec_264_rdo_mb_guess_best_intra_pred_svc, rdo.c:301-461.  The prediction is the
 * resampled base-layer reconstruction (G.8.6.2.1, _hl_codec_264_decode_svc_resample_intra_colour_comps in decode_svc.c -- host side); it is not
 * stored anywhere, so the hook calls the same resampling function once more after the macroblock is done (it reads the reference layer only).
 * tag 9 (once per layer picture): [9, n, frame, DQId, W, H, then W*H*3/2 source bytes, 4 per word]
 * tag 8 (per macroblock): [8, n, frame, DQId, addr, QPy, QPc[2], CodedBlockPatternLuma4x4, CodedBlockPatternChromaDC4x4[2], CodedBlockPatternChromaAC4x4[2],
 *         CodedBlockPatternLuma, CodedBlockPatternChroma, LumaLevel[16][16], ChromaDCLevel[2][4], ChromaACLevel[2][4][16], reconstructed Y 16x16, Cb 8x8, Cr 8x8,
 *         ChromaACLevel / ChromaDCLevel as they were BEFORE the call, prediction Y 16x16, Cb 8x8, Cr 8x8] */
extern HL_ERROR_T _hl_codec_264_decode_svc_resample_intra_colour_comps(hl_codec_264_t* p_codec, hl_codec_264_mb_t* p_mb, int32_t chromaFlag, int32_t iCbCr, int32_t mbW,
                                                                        int32_t mbH, int32_t mbPred[16][16]);
extern HL_ERROR_T __real_hl_codec_264_rdo_mb_guess_best_intra_pred_svc(hl_codec_264_mb_t*, hl_codec_264_t*);
HL_ERROR_T __wrap_hl_codec_264_rdo_mb_guess_best_intra_pred_svc(hl_codec_264_mb_t* p_mb, hl_codec_264_t* p_codec)
{
    int32_t dc_in[2][4], ac_in[2][4][16];
    HL_ERROR_T err;
    { int c_, b_, i_; for (c_ = 0; c_ < 2; ++c_) for (b_ = 0; b_ < 4; ++b_) { dc_in[c_][b_] = p_mb->ChromaDCLevel[c_][b_]; for (i_ = 0; i_ < 16; ++i_) ac_in[c_][b_][i_] = p_mb->ChromaACLevel[c_][b_][i_]; } }
    err = __real_hl_codec_264_rdo_mb_guess_best_intra_pred_svc(p_mb, p_codec);
    if (g_trace && !err) {
        static HL_ALIGNED(16) int32_t pl[16][16], pcb[16][16], pcr[16][16];
        hl_codec_264_layer_t* pc_layer = p_codec->layers.pc_active;
        const hl_codec_264_pict_t* pict = pc_layer->pc_fs_curr->p_pict;
        const int W = (int)pict->uWidthL, H = (int)pict->uHeightL, dq = (int)p_codec->layers.currDQId;
        int32_t r[1400]; int k = 0, i, j, n, x, y;
        if (g_svc_last_frame != g_frame_idx || g_svc_last_dqid != dq) {
            const hl_frame_video_t* in = p_codec->encoder.pc_frame;
            int32_t h9[6] = { 9, 6 + (W * H * 3 / 2) / 4, g_frame_idx, dq, W, H };
            put32(h9, 6);
            put_planes((const uint8_t*)in->data_ptr[0], (const uint8_t*)in->data_ptr[1], (const uint8_t*)in->data_ptr[2], W, H);
            g_svc_last_frame = g_frame_idx; g_svc_last_dqid = dq;
        }
        if (_hl_codec_264_decode_svc_resample_intra_colour_comps(p_codec, p_mb, 0, -1, 16, 16, pl) || _hl_codec_264_decode_svc_resample_intra_colour_comps(p_codec, p_mb, 1, 0, 8, 8, pcb) ||
            _hl_codec_264_decode_svc_resample_intra_colour_comps(p_codec, p_mb, 1, 1, 8, 8, pcr)) return err;
        r[k++] = 8; r[k++] = 0; r[k++] = g_frame_idx; r[k++] = dq; r[k++] = (int32_t)p_mb->u_addr; r[k++] = p_mb->QPy; r[k++] = p_mb->QPc[0]; r[k++] = p_mb->QPc[1];
        r[k++] = (int32_t)p_mb->CodedBlockPatternLuma4x4; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaDC4x4[1];
        r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[0]; r[k++] = (int32_t)p_mb->CodedBlockPatternChromaAC4x4[1];
        r[k++] = (int32_t)p_mb->CodedBlockPatternLuma; r[k++] = (int32_t)p_mb->CodedBlockPatternChroma;
        for (i = 0; i < 16; ++i) for (j = 0; j < 16; ++j) r[k++] = p_mb->LumaLevel[i][j];
        for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = p_mb->ChromaDCLevel[i][j];
        for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = p_mb->ChromaACLevel[i][n][j];
        for (y = 0; y < 16; ++y) for (x = 0; x < 16; ++x) r[k++] = pict->pc_data_y[(p_mb->yL + y) * W + p_mb->xL + x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pict->pc_data_u[(p_mb->yL / 2 + y) * (W / 2) + p_mb->xL / 2 + x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pict->pc_data_v[(p_mb->yL / 2 + y) * (W / 2) + p_mb->xL / 2 + x];
        for (i = 0; i < 2; ++i) for (n = 0; n < 4; ++n) for (j = 0; j < 16; ++j) r[k++] = ac_in[i][n][j];
        for (i = 0; i < 2; ++i) for (j = 0; j < 4; ++j) r[k++] = dc_in[i][j];
        for (y = 0; y < 16; ++y) for (x = 0; x < 16; ++x) r[k++] = pl[y][x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pcb[y][x];
        for (y = 0; y < 8; ++y) for (x = 0; x < 8; ++x) r[k++] = pcr[y][x];
        r[1] = k;
        put32(r, (size_t)k);
    }
    return err;
}

#endif /* HL_DRIVER_NO_WRAPS */

/* ------------------------------------------------------------------------------------------------------------ */
/* Synthetic inputs, SURVEY.md 8(d).  Mirrored bit-for-bit by hartallo_b200/synth.py.                            */
static uint32_t g_lcg = 12345u;
static uint32_t rnd(void) { g_lcg = g_lcg * 1664525u + 1013904223u; return g_lcg >> 8; }

/* G1 "pan": translating checkerboard + ramp + 2-bit noise; LCG continues across frames */
static void gen_g1(uint8_t* yuv, int w, int h, int n)
{
    int x, y, i;
    uint8_t* Y = yuv; uint8_t* UV = yuv + (size_t)w * h;
    for (y = 0; y < h; ++y) for (x = 0; x < w; ++x) {
        int v = 128 + 60 * ((((x + 2 * n) / 8) + ((y + n) / 8)) & 1) + ((((x + 2 * n) * 7) + ((y + n) * 13)) & 31) - 16 + (int)(rnd() & 3);
        Y[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    for (i = 0; i < (w * h) / 2; ++i) UV[i] = (uint8_t)(128 + ((i + n) & 15));
}

/* G2 "stress": a fixed random base picture (half the pixels < 34, 8x8 patches of 0 / 255), translated by (3n, -2n)
 * with wrap-around, plus chroma from the same recipe.  Exercises both clip branches, the F7 wrap and F13. */
/* one base picture per picture size (SVC: every spatial layer has its own, same recipe and seed at the layer's size) */
static struct { uint8_t* base; int w, h; } g2_bases[4];
static void gen_g2(uint8_t* yuv, int w, int h, int n, uint32_t seed)
{
    int x, y, c, k;
    uint8_t* g2_base = NULL;
    for (k = 0; k < 4 && g2_bases[k].base; ++k) if (g2_bases[k].w == w && g2_bases[k].h == h) g2_base = g2_bases[k].base;
    if (!g2_base) {
        uint32_t s = seed * 2654435761u + 97u;
        size_t i, tot = (size_t)w * h * 3 / 2;
        if (k >= 4) { fprintf(stderr, "gen_g2: too many picture sizes\n"); exit(2); }
        g2_base = (uint8_t*)malloc(tot); g2_bases[k].base = g2_base; g2_bases[k].w = w; g2_bases[k].h = h;
        for (i = 0; i < tot; ++i) {
            uint32_t a, b;
            s = s * 1664525u + 1013904223u; a = s >> 8;
            s = s * 1664525u + 1013904223u; b = s >> 8;
            g2_base[i] = (uint8_t)(b % ((a & 1) ? 34u : 256u));
        }
        for (y = 0; y + 8 <= h; y += 8) for (x = 0; x + 8 <= w; x += 8) {
            uint32_t a; int yy, xx;
            s = s * 1664525u + 1013904223u; a = (s >> 8) & 15;
            if (a < 2) for (yy = 0; yy < 8; ++yy) for (xx = 0; xx < 8; ++xx) g2_base[(size_t)(y + yy) * w + x + xx] = (a == 0) ? 0 : 255;
        }
    }
    for (y = 0; y < h; ++y) for (x = 0; x < w; ++x) {
        int sx = ((x + 3 * n) % w + w) % w, sy = ((y - 2 * n) % h + h) % h;
        yuv[(size_t)y * w + x] = g2_base[(size_t)sy * w + sx];
    }
    for (c = 0; c < 2; ++c) {
        int cw = w / 2, ch = h / 2;
        const uint8_t* b = g2_base + (size_t)w * h + (size_t)c * cw * ch;
        uint8_t* o = yuv + (size_t)w * h + (size_t)c * cw * ch;
        for (y = 0; y < ch; ++y) for (x = 0; x < cw; ++x) {
            int sx = ((x + (3 * n) / 2) % cw + cw) % cw, sy = ((y - n) % ch + ch) % ch;
            o[(size_t)y * cw + x] = b[(size_t)sy * cw + sx];
        }
    }
}

/* G3 "texture": a translating picture whose 8x8 blocks carry hashed texture of seven amplitudes (0..90) over a slow ramp, plus 2-bit noise from the LCG.
 * The block amplitudes straddle the homogeneity thresholds of the early-termination mode mask (rdo.c:889-935), so all four of its outcomes occur. */
static void gen_g3(uint8_t* yuv, int w, int h, int n, uint32_t seed)
{
    static const int amp[7] = {0, 4, 10, 20, 36, 56, 90};
    int x, y, i;
    uint8_t* Y = yuv; uint8_t* UV = yuv + (size_t)w * h;
    for (y = 0; y < h; ++y) for (x = 0; x < w; ++x) {
        const uint32_t X = (uint32_t)(x + 2 * n), Yc = (uint32_t)(y + n);
        uint32_t v = X * 374761393u + Yc * 668265263u + seed * 2246822519u;
        int a, s;
        v = (v ^ (v >> 13)) * 1274126177u;
        a = amp[(((X >> 3) * 73u + (Yc >> 3) * 151u + seed) % 7u)];
        s = 96 + (int)(((X + Yc) >> 2) & 31u) + (int)((((v >> 16) & 127u) * (uint32_t)a) >> 7) + (int)(rnd() & 3);
        Y[(size_t)y * w + x] = (uint8_t)(s > 255 ? 255 : s);
    }
    for (i = 0; i < (w * h) / 2; ++i) UV[i] = (uint8_t)(128 + ((i + n) & 15));
}

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

static void md5_hex(const uint8_t* p, size_t n, char out[33])
{
    hl_md5context_t ctx; hl_md5digest_t d; int i;
    hl_md5init(&ctx); hl_md5update(&ctx, p, n); hl_md5final(d, &ctx);
    for (i = 0; i < 16; ++i) sprintf(out + 2 * i, "%02x", d[i]);
    out[32] = 0;
}

/* size of spatial layer l: doubled per layer as source/test_encoder.c:150-202 does, or scaled by --scale n d per layer */
static int g_scale_n = 2, g_scale_d = 1;
static int layer_dim(int base, int l) { int v = base, k; for (k = 0; k < l; ++k) v = v * g_scale_n / g_scale_d; return v; }

int main(int argc, char** argv)
{
    int w = 352, h = 288, frames = 3, qp = 31, me_range = 16, refs = 1, gen = 1, gop = 400, early = 0, deblock = 0, defaults = 0, layers = 1, l, i;
    uint32_t seed = 1;
    const char *in_path = NULL, *out_path = NULL, *trace_path = NULL, *recon_path = NULL, *dump_in = NULL;
    const struct hl_codec_plugin_def_s* plugin = NULL;
    struct hl_codec_s* codec = NULL;
    struct hl_codec_result_s* result = NULL;
    struct hl_frame_video_s* frame = NULL;
    FILE *fin = NULL, *fout = NULL, *fdump = NULL;
    uint8_t *yuv, *stream; size_t frame_bytes, stream_n = 0, stream_cap;
    double t_total = 0, t_p = 0; int n_p = 0; char md5[33];
    HL_ERROR_T err;

    for (i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--size") && i + 2 < argc) { w = atoi(argv[++i]); h = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--qp") && i + 1 < argc) qp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--me-range") && i + 1 < argc) me_range = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--refs") && i + 1 < argc) refs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gop") && i + 1 < argc) gop = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--layers") && i + 1 < argc) layers = atoi(argv[++i]);   /* SVC spatial layers: layer l is (w << l) x (h << l), source/test_encoder.c:150-202 */
        else if (!strcmp(argv[i], "--scale") && i + 2 < argc) { g_scale_n = atoi(argv[++i]); g_scale_d = atoi(argv[++i]); }   /* layer l is (w, h) * (n / d)^l instead of doubled: 3 2 = extended spatial scalability */
        else if (!strcmp(argv[i], "--early-term") && i + 1 < argc) early = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--deblock") && i + 1 < argc) deblock = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--defaults")) defaults = 1;   /* keep what hl_codec_create sets for deblock_flag / me_early_term_flag (both 1, hl_types.h:67,69) */
        else if (!strcmp(argv[i], "--gen") && i + 1 < argc) { ++i; gen = !strcmp(argv[i], "g2") ? 2 : (!strcmp(argv[i], "g3") ? 3 : 1); }
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--in") && i + 1 < argc) in_path = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
        else if (!strcmp(argv[i], "--trace") && i + 1 < argc) trace_path = argv[++i];
        else if (!strcmp(argv[i], "--recon") && i + 1 < argc) recon_path = argv[++i];
        else if (!strcmp(argv[i], "--dump-input") && i + 1 < argc) dump_in = argv[++i];
        else if (!strcmp(argv[i], "--trace-cand")) g_trace_cand = 1;
        else if (!strcmp(argv[i], "--no-levels")) g_trace_levels = 0;
        else if (!strcmp(argv[i], "--trace-state")) g_trace_state = 1;
        else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
    }
    if ((w & 15) || (h & 15)) { fprintf(stderr, "W and H must be multiples of 16 (hl_codec_264.c:428-439)\n"); return 2; }
    g_width = w; g_height = h;
    if (layers < 1 || layers > 3) layers = 1;
    if (g_scale_n < g_scale_d || g_scale_d < 1) { fprintf(stderr, "--scale n d needs n >= d >= 1\n"); return 2; }
    for (l = 0; l < layers; ++l) if ((layer_dim(w, l) & 15) || (layer_dim(h, l) & 15)) { fprintf(stderr, "layer %d is %d x %d: not whole macroblocks\n", l, layer_dim(w, l), layer_dim(h, l)); return 2; }
    frame_bytes = (size_t)layer_dim(w, layers - 1) * layer_dim(h, layers - 1) * 3 / 2;   /* largest layer */
    yuv = (uint8_t*)malloc(frame_bytes);
    stream_cap = frame_bytes * (size_t)(frames + 1) + 65536; stream = (uint8_t*)malloc(stream_cap);
    if (in_path && !(fin = fopen(in_path, "rb"))) { perror(in_path); return 2; }
    if (out_path && !(fout = fopen(out_path, "wb"))) { perror(out_path); return 2; }
    if (trace_path && !(g_trace = fopen(trace_path, "wb"))) { perror(trace_path); return 2; }
    if (recon_path && !(g_recon = fopen(recon_path, "wb"))) { perror(recon_path); return 2; }
    if (dump_in && !(fdump = fopen(dump_in, "wb"))) { perror(dump_in); return 2; }

    hl_debug_set_level(HL_DEBUG_LEVEL_ERROR);
    hl_engine_set_cpu_flags(0);                       /* pure C path: the oracle */
    if ((err = hl_engine_init())) { fprintf(stderr, "engine init %d\n", err); return 1; }
#ifndef HL_DRIVER_NO_WRAPS
    if (g_trace_cand) { g_real_sad = hl_math_sad4x4_u8; hl_math_sad4x4_u8 = traced_sad; }
#endif
    if ((err = hl_codec_plugin_find(HL_CODEC_TYPE_H264_SVC, &plugin))) { fprintf(stderr, "plugin find %d\n", err); return 1; }
    if ((err = hl_codec_create(plugin, &codec))) { fprintf(stderr, "codec create %d\n", err); return 1; }
    if ((err = hl_codec_result_create(&result))) return 1;
    if ((err = hl_frame_video_create(&frame))) return 1;

    /* same knobs as source/test_encoder.c:135-146 */
    codec->gop_size = gop;
    codec->me_range = me_range;
    codec->qp = qp;
    codec->fps.num = 1; codec->fps.den = 30;
    codec->rc_bitrate = -1;
    if (!defaults) codec->deblock_flag = deblock;
    codec->threads_count = 1;
    codec->max_ref_frame = refs;
    codec->distortion_mesure_type = HL_VIDEO_DISTORTION_MESURE_TYPE_SAD;
    codec->me_type = (HL_VIDEO_ME_TYPE_INTEGER | HL_VIDEO_ME_TYPE_HALF | HL_VIDEO_ME_TYPE_QUATER);
    codec->me_part_types = HL_VIDEO_ME_PART_TYPE_ALL;
    codec->me_subpart_types = HL_VIDEO_ME_SUBPART_TYPE_ALL;
    if (!defaults) codec->me_early_term_flag = early;

    if (layers > 1)
        for (l = 0; l < layers; ++l)
            if ((err = hl_codec_add_layer(codec, (uint32_t)layer_dim(w, l), (uint32_t)layer_dim(h, l), 0, 0))) { fprintf(stderr, "add_layer %d failed: %d\n", l, err); return 1; }
    for (i = 0; i < frames; ++i) {
        for (l = 0; l < layers; ++l) {   /* one hl_codec_encode per layer per access unit (test_encoder.c:174-202) */
            double t0, t1;
            const int lw = layer_dim(w, l), lh = layer_dim(h, l);
            const size_t lbytes = (size_t)lw * lh * 3 / 2;
            if (fin) { if (fread(yuv, 1, lbytes, fin) != lbytes) goto done; }
            else if (gen == 1) gen_g1(yuv, lw, lh, i);
            else if (gen == 3) gen_g3(yuv, lw, lh, i, seed);
            else gen_g2(yuv, lw, lh, i, seed);
            if (fdump) fwrite(yuv, 1, lbytes, fdump);
            if ((err = hl_frame_video_fill(frame, HL_VIDEO_CHROMA_YUV420, lw, lh, yuv, lbytes))) { fprintf(stderr, "fill %d\n", err); return 1; }
            frame->encoding = HL_VIDEO_ENCODING_TYPE_AUTO;
            g_frame_idx = i;
            t0 = now_ms();
            err = hl_codec_encode(codec, (hl_frame_t*)frame, result);
            t1 = now_ms();
            if (err) { fprintf(stderr, "encode frame %d layer %d failed: %d\n", i, l, err); return 1; }
            t_total += t1 - t0; if (i > 0) { t_p += t1 - t0; if (l == layers - 1) ++n_p; }
            if (result->type & HL_CODEC_RESULT_TYPE_HDR) {
                memcpy(stream + stream_n, codec->hdr_bytes, codec->hdr_bytes_count); stream_n += codec->hdr_bytes_count;
            }
            if ((result->type & HL_CODEC_RESULT_TYPE_DATA) && l == layers - 1) {   /* the access unit is returned with its last layer */
                static const uint8_t scp[3] = { 0, 0, 1 };
                memcpy(stream + stream_n, scp, 3); stream_n += 3;
                memcpy(stream + stream_n, result->data_ptr, result->data_size); stream_n += result->data_size;
            }
            fprintf(stderr, "frame %d layer %d: %.1f ms, %zu bytes so far\n", i, l, t1 - t0, stream_n);
        }
    }
done:
    if (fout) fwrite(stream, 1, stream_n, fout);
    md5_hex(stream, stream_n, md5);
    {
        int mbs = (w / 16) * (h / 16);
        printf("{\"frames\": %d, \"width\": %d, \"height\": %d, \"bytes\": %zu, \"md5\": \"%s\", \"ms_total\": %.3f, "
               "\"ms_p_frames\": %.3f, \"p_frames\": %d, \"mb_per_s_p\": %.1f, \"mb_per_s_all\": %.1f}\n",
               i, w, h, stream_n, md5, t_total, t_p, n_p,
               n_p ? (double)mbs * n_p / (t_p * 1e-3) : 0.0, t_total > 0 ? (double)mbs * i / (t_total * 1e-3) : 0.0);
    }
    if (g_trace) fclose(g_trace);
    if (g_recon) fclose(g_recon);
    if (fdump) fclose(fdump);
    if (fout) fclose(fout);
    if (fin) fclose(fin);
    return 0;
}
