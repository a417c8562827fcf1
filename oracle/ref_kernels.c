/*
 * ref_kernels.c -- thin C entry points around the UNMODIFIED reference kernels so that tests can call them with plain
 * pointers (ctypes).  Built by oracle/build_ref.sh into oracle/_ref/libref_kernels.so together with the reference
 * objects.  TEST INFRASTRUCTURE ONLY.  Nothing here re-implements arithmetic: each function fills the minimum of the
 * reference's own structs (calloc'ed `hl_codec_264_t`, `hl_codec_264_mb_t`, ...) and calls the reference function
 * named in its comment.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "hartallo/hl_api.h"
#include "hartallo/hl_math.h"
#include "hartallo/hl_memory.h"
#include "hartallo/hl_debug.h"
#include "hartallo/h264/hl_codec_264.h"
#include "hartallo/h264/hl_codec_264_mb.h"
#include "hartallo/h264/hl_codec_264_layer.h"
#include "hartallo/h264/hl_codec_264_encode.h"
#include "hartallo/h264/hl_codec_264_slice.h"
#include "hartallo/h264/hl_codec_264_sps.h"
#include "hartallo/h264/hl_codec_264_pps.h"
#include "hartallo/h264/hl_codec_264_pict.h"
#include "hartallo/h264/hl_codec_264_dpb.h"
#include "hartallo/h264/hl_codec_264_bits.h"
#include "hartallo/h264/hl_codec_264_residual.h"
#include "hartallo/h264/hl_codec_264_interpol.h"
#include "hartallo/h264/hl_codec_264_transf.h"
#include "hartallo/h264/hl_codec_264_quant.h"
#include "hartallo/h264/hl_codec_264_macros.h"

extern HL_ERROR_T hl_codec_264_interpol_luma(hl_codec_264_t*, hl_codec_264_mb_t*, int32_t, int32_t, const hl_codec_264_mv_xt*, const hl_pixel_t*, void*, int32_t);

static struct {
    int ready;
    hl_codec_264_t* codec;
    hl_codec_264_layer_t* layer;
    hl_codec_264_dpb_t* dpb;
    hl_codec_264_slice_t* slice;
    hl_codec_264_nal_slice_header_t* hdr;
    hl_codec_264_nal_pps_t* pps;
    hl_codec_264_nal_sps_t* sps;
    hl_codec_264_encode_slice_data_t* esd;
    hl_codec_264_mb_t* mb[3]; /* 0 = A, 1 = B, 2 = current */
    int idx_w, idx_h;
} G;

static const int kNormAdjust[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};

int ref_init(void)
{
    int a, b, m, i, j;
    if (G.ready) return 0;
    hl_debug_set_level(HL_DEBUG_LEVEL_ERROR);
    hl_engine_set_cpu_flags(0);
    if (hl_engine_init()) return -1;
    G.codec = (hl_codec_264_t*)calloc(1, sizeof(*G.codec));
    G.layer = (hl_codec_264_layer_t*)calloc(1, sizeof(*G.layer));
    G.dpb = (hl_codec_264_dpb_t*)calloc(1, sizeof(*G.dpb));
    G.slice = (hl_codec_264_slice_t*)calloc(1, sizeof(*G.slice));
    G.hdr = (hl_codec_264_nal_slice_header_t*)calloc(1, sizeof(*G.hdr));
    G.pps = (hl_codec_264_nal_pps_t*)calloc(1, sizeof(*G.pps));
    G.sps = (hl_codec_264_nal_sps_t*)calloc(1, sizeof(*G.sps));
    G.esd = (hl_codec_264_encode_slice_data_t*)calloc(1, sizeof(*G.esd));
    for (i = 0; i < 3; ++i) G.mb[i] = (hl_codec_264_mb_t*)calloc(1, sizeof(hl_codec_264_mb_t));
    G.sps->BitDepthY = 8; G.sps->BitDepthC = 8; G.sps->ChromaArrayType = 1;
    G.sps->SubWidthC_TrailingZeros = 1; G.sps->SubHeightC_TrailingZeros = 1; G.sps->MbWidthC = 8; G.sps->MbHeightC = 8;
    G.pps->pc_sps = G.sps;
    /* flat scaling lists (Flat_4x4_16): LevelScale4x4 = 16 * normAdjust4x4, 8.5.9 -- what pps.c:38-80 computes when no
       scaling matrix is present; the encoder-level traces pin the real table */
    for (a = 0; a < 2; ++a) for (b = 0; b < 3; ++b) for (m = 0; m < 6; ++m) for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j)
        G.pps->LevelScale4x4[a][b][m][i][j] = 16 * kNormAdjust[m][((i | j) & 1) == 0 ? 0 : ((i & j & 1) ? 1 : 2)];
    G.hdr->pc_pps = G.pps;
    G.hdr->SliceTypeModulo5 = HL_CODEC_264_SLICE_TYPE_P;
    G.slice->p_header = G.hdr; G.slice->u_idx = 0;
    G.layer->pc_slice_hdr = G.hdr; G.layer->p_list_slices[0] = G.slice; G.layer->pc_slice_curr = G.slice;
    G.layer->pp_list_macroblocks = (hl_codec_264_mb_t**)calloc(3, sizeof(void*));
    for (i = 0; i < 3; ++i) { G.layer->pp_list_macroblocks[i] = G.mb[i]; G.mb[i]->u_addr = (uint32_t)i; }
    G.layer->u_list_macroblocks_count = 3;
    G.esd->pc_slice = G.slice;
    if (hl_memory_blocks_create(&G.esd->pc_mem_blocks)) return -2;
    if (hl_codec_264_bits_create_2(&G.esd->rdo.pobj_bits)) return -3;
    G.layer->encoder.p_list_esd[0] = G.esd;
    G.codec->layers.pc_active = G.layer;
    G.codec->sps.pc_active = G.sps; G.codec->pps.pc_active = G.pps;
    G.codec->pc_dpb = G.dpb;
    if (hl_memory_blocks_create(&G.codec->pobj_mem_blocks)) return -4;
    for (i = 0; i < 4; ++i) { G.codec->PixelMaxValueY[i] = 255; G.codec->PixelMaxValueC[i] = 255; }
    G.ready = 1;
    return 0;
}

static int set_picture(int W, int H)
{
    if (G.idx_w != W || G.idx_h != H) {
        HL_OBJECT_SAFE_FREE(G.dpb->p_list_interpol_indices[0]);
        if (hl_codec_264_interpol_indices_create(&G.dpb->p_list_interpol_indices[0], (uint32_t)W, (uint32_t)H, HL_FALSE)) return -1;
        G.idx_w = W; G.idx_h = H;
    }
    G.hdr->PicWidthInSamplesL = (uint32_t)W; G.hdr->PicHeightInSamplesL = (uint32_t)H;
    G.hdr->PicWidthInSamplesC = (uint32_t)(W >> 1); G.hdr->PicHeightInSamplesC = (uint32_t)(H >> 1);
    return 0;
}

/* hl_codec_264_interpol_luma, source/h264/hl_codec_264_pred_inter.c:339 (u8 output, the path the encoder uses) */
int ref_interp_luma(const uint8_t* ref, int W, int H, int xL, int yL, int pw, int ph, int mvx, int mvy, uint8_t* out /*[16][16]*/)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    hl_codec_264_mv_xt mv;
    HL_ALIGN(HL_ALIGN_V) uint8_t pred[16][16];
    if (ref_init() || set_picture(W, H)) return -1;
    mb->u_slice_idx = 0; mb->mb_field_decoding_flag = 0;
    mb->partWidth[0][0] = pw; mb->partHeight[0][0] = ph; mb->xL_Idx = xL; mb->yL_Idx = yL;
    mv.x = mvx; mv.y = mvy;
    memset(pred, 0, sizeof(pred));
    if (hl_codec_264_interpol_luma(G.codec, mb, 0, 0, &mv, ref, &pred[0][0], 1)) return -2;
    memcpy(out, pred, 256);
    return 0;
}

/* hl_codec_264_interpol_chroma (-> hl_codec_264_interpol_chroma_cpp, pred_inter.c:888) */
int ref_interp_chroma(const uint8_t* refU, const uint8_t* refV, int W, int H, int xL, int yL, int pwc, int phc, int mvx, int mvy,
                      int32_t* outU /*[16][16]*/, int32_t* outV)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    hl_codec_264_mv_xt mv;
    hl_codec_264_pict_t pict;
    HL_ALIGN(HL_ALIGN_V) int32_t cb[16][16], cr[16][16];
    if (ref_init() || set_picture(W, H)) return -1;
    memset(&pict, 0, sizeof(pict));
    pict.pc_data_u = (hl_pixel_t*)refU; pict.pc_data_v = (hl_pixel_t*)refV;
    pict.uWidthL = (uint32_t)W; pict.uHeightL = (uint32_t)H; pict.uWidthC = (uint32_t)(W >> 1); pict.uHeightC = (uint32_t)(H >> 1);
    mb->u_slice_idx = 0; mb->mb_field_decoding_flag = 0;
    mb->partWidthC[0][0] = pwc; mb->partHeightC[0][0] = phc; mb->xL_Idx = xL; mb->yL_Idx = yL;
    mv.x = mvx; mv.y = mvy;
    memset(cb, 0, sizeof(cb)); memset(cr, 0, sizeof(cr));
    if (hl_codec_264_interpol_chroma(G.codec, mb, 0, 0, &mv, &mv, &pict, &pict, cb, cr)) return -2;
    memcpy(outU, cb, sizeof(cb)); memcpy(outV, cr, sizeof(cr));
    return 0;
}

/* hl_codec_264_transf_frw_residual4x4 (transf.c:716) */
void ref_fwd4x4(const int32_t* in, int32_t* out)
{
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, in, 64);
    hl_codec_264_transf_frw_residual4x4(a, b);
    memcpy(out, b, 64);
}
/* hl_codec_264_quant_frw4x4_scale_ac (quant.c:116) */
void ref_quant4x4(int qp, int intra, const int32_t* in, int32_t* out)
{
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, in, 64);
    hl_codec_264_quant_frw4x4_scale_ac(qp, intra ? HL_TRUE : HL_FALSE, a, b);
    memcpy(out, b, 64);
}
/* hl_codec_264_transf_scale_residual4x4 (transf.c:376): dequant (quant.c:68) + inverse transform (transf.c:420) */
void ref_dequant_inv4x4(int qp, int mb_is_inter, int luma, int intra16x16, int cbcr, const int32_t* c, int32_t* r)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, c, 64);
    mb->e_type = mb_is_inter ? HL_CODEC_264_MB_TYPE_P_L0_16X16 : HL_CODEC_264_MB_TYPE_I_NXN;
    mb->flags_type = mb_is_inter ? HL_CODEC_264_MB_TYPE_FLAGS_INTER_P : HL_CODEC_264_MB_TYPE_FLAGS_INTRA_4x4;
    mb->QPyprime = qp; mb->QPy = qp; mb->QPprimeC[0] = mb->QPprimeC[1] = qp; mb->TransformBypassModeFlag = 0;
    hl_codec_264_transf_scale_residual4x4(G.codec, mb, (const int32_t(*)[4])a, b, luma ? HL_TRUE : HL_FALSE, intra16x16 ? HL_TRUE : HL_FALSE, cbcr);
    memcpy(r, b, 64);
}
/* hl_codec_264_transf_frw_hadamard4x4_dc_luma (transf.c:774) */
void ref_hadamard4x4_dc_luma(const int32_t* in, int32_t* out)
{
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, in, 64);
    hl_codec_264_transf_frw_hadamard4x4_dc_luma(a, b);
    memcpy(out, b, 64);
}
/* hl_codec_264_quant_frw4x4_scale_dc_luma (quant.c:141) */
void ref_quant_dc_luma(int qp, int intra, const int32_t* in, int32_t* out)
{
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, in, 64);
    hl_codec_264_quant_frw4x4_scale_dc_luma(qp, intra ? HL_TRUE : HL_FALSE, a, b);
    memcpy(out, b, 64);
}
/* hl_codec_264_transf_scale_luma_dc_coeff_intra16x16 (transf.c:498) */
void ref_scale_luma_dc(int qp, const int32_t* c, int32_t* dcY)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    HL_ALIGN(HL_ALIGN_V) int32_t a[4][4], b[4][4];
    ref_init(); memcpy(a, c, 64);
    mb->QPy = qp; mb->QPyprime = qp; mb->TransformBypassModeFlag = 0;
    hl_codec_264_transf_scale_luma_dc_coeff_intra16x16(G.codec, mb, qp, 8, a, b);
    memcpy(dcY, b, 64);
}
/* hl_codec_264_transf_frw_hadamard2x2_dc_chroma (transf.c:843) + hl_codec_264_quant_frw2x2_scale_dc_chroma (quant.c:168) */
void ref_hadamard2x2_quant_dc_chroma(int qp, int intra, const int32_t* in /*4*/, int32_t* had /*4*/, int32_t* q /*4*/)
{
    HL_ALIGN(HL_ALIGN_V) int32_t a[2][2], b[2][2], c[2][2];
    ref_init(); memcpy(a, in, 16);
    hl_codec_264_transf_frw_hadamard2x2_dc_chroma(a, b);
    hl_codec_264_quant_frw2x2_scale_dc_chroma(qp, intra ? HL_TRUE : HL_FALSE, b, c);
    memcpy(had, b, 16); memcpy(q, c, 16);
}
/* hl_codec_264_transf_scale_chroma_dc_coeff (transf.c:612) */
void ref_scale_chroma_dc(int qpc, const int32_t* c4 /*c00,c01,c10,c11*/, int32_t* dc4)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    int32_t c[4][4], dcC[4][2];
    ref_init(); memset(c, 0, sizeof(c)); memset(dcC, 0, sizeof(dcC));
    c[0][0] = c4[0]; c[0][1] = c4[1]; c[1][0] = c4[2]; c[1][1] = c4[3];
    mb->TransformBypassModeFlag = 0;
    hl_codec_264_transf_scale_chroma_dc_coeff(G.codec, mb, (const int32_t(*)[4])c, 8, qpc, 0, dcC);
    dc4[0] = dcC[0][0]; dc4[1] = dcC[0][1]; dc4[2] = dcC[1][0]; dc4[3] = dcC[1][1];
}
/* hl_math_sad4x4_u8 (hl_math.c:239), hl_math_satd4x4_u8 (hl_math.c:283) */
int ref_sad4x4(const uint8_t* a, int sa, const uint8_t* b, int sb) { ref_init(); return hl_math_sad4x4_u8(a, sa, b, sb); }
int ref_satd4x4(const uint8_t* a, int sa, const uint8_t* b, int sb)
{
    HL_ALIGN(HL_ALIGN_V) uint8_t x[16], y[16]; int i;
    ref_init();
    for (i = 0; i < 4; ++i) { memcpy(x + 4 * i, a + i * sa, 4); memcpy(y + 4 * i, b + i * sb, 4); }
    return hl_math_satd4x4_u8(x, 4, y, 4);
}
/* hl_math_ssd4x4_u8 (hl_math.c:360), hl_math_homogeneousity8x8_u8 (hl_math.c:470) */
int ref_ssd4x4(const uint8_t* a, int sa, const uint8_t* b, int sb)
{
    HL_ALIGN(HL_ALIGN_V) uint8_t x[16], y[16]; int i;
    ref_init();
    for (i = 0; i < 4; ++i) { memcpy(x + 4 * i, a + i * sa, 4); memcpy(y + 4 * i, b + i * sb, 4); }
    return hl_math_ssd4x4_u8(x, 4, y, 4);
}
int ref_homogeneity8x8(const uint8_t* p, int stride) { ref_init(); return hl_math_homogeneousity8x8_u8(p, stride); }
/* hl_math_addclip_4x4_u8xi32 (hl_math.h:303, wraps) and hl_math_addclip_4x4 (hl_math.h:278, clips) */
void ref_addclip_u8xi32(const uint8_t* pred /*16*/, const int32_t* res /*16*/, uint8_t* out /*16*/)
{
    HL_ALIGN(HL_ALIGN_V) uint8_t p[16], o[16]; HL_ALIGN(HL_ALIGN_V) int32_t r[16];
    ref_init(); memcpy(p, pred, 16); memcpy(r, res, 64);
    hl_math_addclip_4x4_u8xi32(p, 4, r, 4, o, 4);
    memcpy(out, o, 16);
}
void ref_addclip_i32(const int32_t* pred, const int32_t* res, int32_t* out)
{
    HL_ALIGN(HL_ALIGN_V) int32_t p[16], r[16], o[16];
    ref_init(); memcpy(p, pred, 64); memcpy(r, res, 64);
    hl_math_addclip_4x4(p, 4, r, 4, G.codec->PixelMaxValueY, o, 4);
    memcpy(out, o, 64);
}

/* hl_codec_264_residual_write_block_cavlc (residual.c:587) in RDO mode for a luma 4x4 block whose neighbours A/B
 * have nA/nB coefficients (pass -1 for "not available").  Returns bit count; *single_ctr, *total_coeff as stored. */
int ref_cavlc_luma_bits(const int32_t* lv16, int nA, int nB, int32_t* single_ctr, int32_t* total_coeff)
{
    hl_codec_264_mb_t* mb = G.mb[2];
    hl_codec_264_residual_inv_xt inv;
    HL_ALIGN(HL_ALIGN_V) int32_t lv[16];
    int bits;
    ref_init();
    memset(&inv, 0, sizeof(inv));
    inv.e_type = HL_CODEC_264_RESISUAL_INV_TYPE_LUMA_LEVEL; inv.b_rdo = HL_TRUE; inv.i_luma4x4BlkIdx = 0;
    memcpy(lv, lv16, 64);
    mb->u_slice_idx = 0;
    mb->e_type = HL_CODEC_264_MB_TYPE_P_L0_16X16; mb->flags_type = HL_CODEC_264_MB_TYPE_FLAGS_INTER_P;
    mb->neighbouringLumaBlock4x4[0].i_addr_A = nA >= 0 ? 0 : HL_CODEC_264_MB_ADDR_NOT_AVAIL;
    mb->neighbouringLumaBlock4x4[0].i_blk_idx_A = 5;
    mb->neighbouringLumaBlock4x4[0].i_addr_B = nB >= 0 ? 1 : HL_CODEC_264_MB_ADDR_NOT_AVAIL;
    mb->neighbouringLumaBlock4x4[0].i_blk_idx_B = 10;
    G.mb[0]->e_type = HL_CODEC_264_MB_TYPE_P_L0_16X16; G.mb[0]->flags_type = HL_CODEC_264_MB_TYPE_FLAGS_INTER_P;
    G.mb[1]->e_type = HL_CODEC_264_MB_TYPE_P_L0_16X16; G.mb[1]->flags_type = HL_CODEC_264_MB_TYPE_FLAGS_INTER_P;
    G.mb[0]->CodedBlockPatternLuma = 15; G.mb[1]->CodedBlockPatternLuma = 15;
    G.mb[0]->TotalCoeffsLuma[5] = nA; G.mb[1]->TotalCoeffsLuma[10] = nB;
    hl_codec_264_bits_reset(G.esd->rdo.pobj_bits, G.esd->rdo.bits_buff, HL_CODEC_264_RDO_BUFFER_MAX_SIZE);
    if (hl_codec_264_residual_write_block_cavlc(&inv, G.codec, mb, G.esd->rdo.pobj_bits, lv, 0, 15, 16)) return -1;
    bits = (int)hl_codec_264_bits_get_stream_index(G.esd->rdo.pobj_bits);
    *single_ctr = G.esd->rdo.Single_ctr;
    *total_coeff = mb->TotalCoeffsLuma[0];
    return bits;
}
