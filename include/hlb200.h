/*
 * hlb200.h -- C-ABI of libhl_b200.so: the B200 (sm_100a) implementation of hartallo's H.264 encoder pixel hot path.
 *
 * This is the drop-in boundary.  Host code stays C (the reference's own host code, unchanged) and reaches CUDA only
 * through these entry points: plain pointers and sizes, no C++/torch types.  Every function returns an `int` that
 * maps 1:1 onto the reference's HL_ERROR_T (include/hartallo/hl_types.h:101-122); 0 = HL_ERROR_SUCCESS.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns HLB200_ERR_SYSTEM.
 *
 * What each group replaces in the reference (file:line under the reference tree):
 *
 *   hlb200_slice_encode        hl_codec_264_nal_slice_data_encode            source/h264/hl_codec_264_slice.c:1701
 *                              (the per-MB decide+reconstruct part of its loop :1786-1894, i.e.
 *                               hl_codec_264_rdo_mb_guess_best_inter_pred_avc source/h264/hl_codec_264_rdo.c:678,
 *                               hl_codec_264_rdo_mb_guess_best_intra_pred_avc rdo.c:99,
 *                               hl_codec_264_me_ds_mb_find_best_cost          source/h264/hl_codec_264_me_ds.c:104);
 *                              with hlb200_slice_params_t::me_early_term_flag the homogeneous-block mode mask of rdo.c:889-935, with ::deblock_flag the
 *                              loop filter hl_codec_264_deblock_avc source/h264/hl_codec_264_deblock.c:192 over the finished picture (slice.c:1897)
 *   hlb200_slice_bits_*        slice_data() of the pictures just encoded, written on the device: _hl_codec_264_mb_write_no_pcm
 *                              source/h264/hl_codec_264_mb.c:543-900, hl_codec_264_residual_write residual.c:903-1094, cavlc.c:652-836,
 *                              mb_skip_run slice.c:1840-1868.  (hlb200_records_download hands the decision records to a host writer instead:
 *                              streams with SVC layers, whose inter-layer derivation reads them on the host.)
 *   hlb200_interp_luma/chroma  hl_codec_264_interpol_luma  source/h264/hl_codec_264_pred_inter.c:339,
 *                              hl_codec_264_interpol_chroma_cpp pred_inter.c:888 (whole-frame batch, one MV set per MB)
 *   hlb200_tq_recon            _hl_codec_264_rdo_mb_reconstruct_inter rdo.c:2274 (residual part :2428-2478) and
 *                              _hl_codec_264_rdo_mb_reconstruct_chroma rdo.c:2502 (whole-frame batch)
 *   hlb200_dev_svc_inter_recon_batch  hl_codec_264_rdo_mb_guess_best_inter_pred_svc rdo.c:1273 (SVC enhancement layer, base-mode inter
 *                              macroblocks: prediction + residual coding + reconstruction fused, whole-picture batch)
 *   hlb200_dev_svc_derive_motion_batch   hl_codec_264_utils_derivation_process_initialisation_svc utils.c:1225 + ..._for_mv_comps_and_ref_indices_svc utils.c:1498
 *                                        (inter-layer motion derivation, P pictures with base_mode_flag = 1)
 *   hlb200_dev_svc_resample_intra_batch  _hl_codec_264_decode_svc_resample_intra_colour_comps decode_svc.c:2864 (Intra_Base resampling, I pictures)
 *   hlb200_dev_svc_bl_recon_batch     hl_codec_264_rdo_mb_guess_best_intra_pred_svc rdo.c:301 (I_BL macroblocks: residual coding + reconstruction
 *                              against the host-resampled base layer)
 *   hlb200_sad4x4              hl_math_sad4x4_u8 source/hl_math.c:239, hl_math_satd4x4_u8 hl_math.c:283, hl_math_ssd4x4_u8 hl_math.c:360 (whole-frame batch)
 *   hlb200_homogeneity8x8      hl_math_homogeneousity8x8_u8 hl_math.c:470 (whole-frame batch)
 *   hlb200_me_cost             hl_codec_264_me_ds_mb_compute_cost_mode me_ds.c:527 (batch of independent candidates)
 *
 * Ownership: device memory belongs to the stream context; host buffers are caller-owned.  One CUDA stream per
 * context.  Unlike the reference (rdo.c:97 `static double last_best_intra_cost`) many contexts may live in one process; the
 * library's process-wide state is read-mostly (resident-CTA counts and one launch-ordering event per device, the kernel-variant
 * override) plus a thread-local error string.
 *
 * ORDERING RULE of the batch entry points (hlb200_slice_encode_batch_async, hlb200_slice_bits_batch_async): the kernels run on the
 * stream of ctxs[0].  Work queued earlier on the other contexts' streams (uploads) is made a dependency of the launch, and everything
 * queued later on them (downloads, the next upload) waits for it -- events, no host synchronisation.  Batch launches of one device run
 * back to back in submission order.  A context must stay alive until the launches that covered it have completed (it may own their
 * scheduler storage); hlb200_stream_destroy synchronises its stream first.
 */
#ifndef HLB200_H_
#define HLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HLB200_API __attribute__((visibility("default")))
#else
#define HLB200_API
#endif

/* same numeric values as HL_ERROR_T */
enum {
    HLB200_OK = 0,
    HLB200_ERR_INVALID_PARAMETER = 1,
    HLB200_ERR_INVALID_STATE = 3,
    HLB200_ERR_NOT_IMPLEMENTED = 7,
    HLB200_ERR_OUTOFMEMORY = 8,
    HLB200_ERR_SYSTEM = 13
};

#define HLB200_MAX_REFS 16

typedef struct hlb200_ctx hlb200_ctx_t; /* one per encoded stream (or SVC layer) */

/* Partition layout + motion of one macroblock: mirrors NumMbPart / sub_mb_type / mvL0[4][4] / refIdxL0[4] of
 * hl_codec_264_mb_t (include/hartallo/h264/hl_codec_264_mb.h:98-269). */
typedef struct hlb200_mb_motion {
    uint8_t part_mode;   /* 0: 16x16, 1: 16x8, 2: 8x16, 3: 8x8 */
    uint8_t sub_mode[4]; /* for 8x8: 0: 8x8, 1: 8x4, 2: 4x8, 3: 4x4 */
    int8_t ref_idx[4];   /* per mbPartIdx */
    uint8_t pad[3];
    int16_t mv[4][4][2]; /* [mbPartIdx][subMbPartIdx][x,y], quarter-pel */
} hlb200_mb_motion_t;

enum { HLB200_MB_P_SKIP = 0, HLB200_MB_P_INTER = 1, HLB200_MB_I16x16 = 2, HLB200_MB_I4x4 = 3 };

/* Per-macroblock decision record: exactly the fields the reference's writer reads (mb.c:584-860, residual.c:987-1094). */
typedef struct hlb200_mb_record {
    uint8_t mb_class;                 /* HLB200_MB_* */
    uint8_t mb_type;                  /* syntax element value, Table 7-11/7-13 (P slices: already +5 for intra) */
    uint8_t part_mode;                /* as hlb200_mb_motion_t */
    uint8_t sub_mode[4];
    uint8_t i16_pred_mode;            /* Intra16x16PredMode */
    uint8_t intra_chroma_pred_mode;
    uint8_t coded_block_pattern;      /* coded_block_pattern (after the ">47 => -16" quirk, rdo.c:2776) */
    uint8_t cbp_luma, cbp_chroma;     /* CodedBlockPatternLuma / Chroma */
    uint8_t cbp_chroma_dc4x4[2], cbp_chroma_ac4x4[2];
    uint16_t cbp_luma4x4;             /* CodedBlockPatternLuma4x4 */
    int8_t mb_qp_delta;
    uint8_t qp_y, qp_c[2];
    int8_t ref_idx[4];
    uint8_t i4_pred_mode[16];         /* Intra4x4PredMode */
    uint8_t prev_intra4x4_pred_mode_flag[16];
    uint8_t rem_intra4x4_pred_mode[16];
    uint8_t pad[3];
    int16_t mv[4][4][2];              /* mvL0 */
    int16_t mvd[4][4][2];             /* mvd_l0 */
    int32_t mad;                      /* best distortion (rate-control hook, rdo.c:1265) */
    int16_t luma_level[16][16];       /* LumaLevel (zig-zag) */
    int16_t i16_dc_level[16];         /* Intra16x16DCLevel */
    int16_t i16_ac_level[16][16];     /* Intra16x16ACLevel (15 used) */
    int16_t chroma_dc_level[2][4];    /* ChromaDCLevel */
    int16_t chroma_ac_level[2][4][16];/* ChromaACLevel (15 used) */
    /* work the reference trajectory performed for this macroblock (roofline accounting) */
    uint32_t me_trials;               /* 4x4 trial encodes of the motion search */
    uint32_t me_interp_ops;           /* integer ops of the interpolation of those blocks, by fractional class */
    uint16_t me_candidates;           /* candidate motion vectors costed */
    uint16_t intra_trials;            /* 4x4 trial encodes of the intra decision */
    uint32_t t_start_ns, t_end_ns;    /* device globaltimer (low 32 bits) when the macroblock started / finished: scheduling analysis */
} hlb200_mb_record_t;

typedef struct hlb200_slice_params {
    int32_t slice_type;  /* 0: I (intra only), 1: P */
    int32_t qp;          /* SliceQPY; fixed QP (rate control stays on the host, rc_bitrate < 0) */
    int32_t me_range;    /* codec->me_range, clipped to [1,64] as rdo.c:847 */
    int32_t num_refs;    /* num_ref_idx_l0_active_minus1 + 1 */
    int32_t chroma_qp_index_offset;
    int32_t cur_slot;    /* frame-store slot receiving the reconstruction */
    int32_t ref_slot[HLB200_MAX_REFS]; /* RefPicList0[i] -> frame-store slot */
    int32_t me_early_term_flag; /* codec->me_early_term_flag: homogeneous-block detection masks partition modes (rdo.c:889-935, hl_math.c:470) */
    int32_t deblock_flag;       /* codec->deblock_flag: loop filter over the finished picture before it becomes a reference (slice.c:1897, deblock.c:192) */
} hlb200_slice_params_t;

/* Levels / flags produced by the batch transform-quant-reconstruct kernel, per macroblock */
typedef struct hlb200_mb_coeffs {
    int16_t luma_level[16][16];
    int16_t chroma_dc_level[2][4];
    int16_t chroma_ac_level[2][4][16];
    uint16_t cbp_luma4x4;
    uint8_t cbp_chroma_dc4x4[2], cbp_chroma_ac4x4[2];
    uint8_t pad[2];
} hlb200_mb_coeffs_t;

/* What the reference's macroblock object carries from one picture of an SVC layer to the next and reads again:
 * ChromaACLevel is only rewritten for blocks with a non-zero residual but used whenever a block's de-quantised DC is non-zero
 * (source/h264/hl_codec_264_transf.c:236-245); ChromaDCLevel is only rewritten when a plane has a non-zero DC coefficient (rdo.c:2653). */
typedef struct hlb200_svc_mb_state {
    int16_t chroma_ac_level[2][4][16];
    int16_t chroma_dc_level[2][4];
} hlb200_svc_mb_state_t;

/* What the SVC inter-layer motion derivation reads of one REFERENCE-LAYER macroblock: the fields of hl_codec_264_mb_t that utils.c:1701 (intra test),
 * mb.h:313-339 (partition indices of a luma location) and utils.c:1807-1834 (predFlagL0 / refIdxL0 / mvL0) touch, copied verbatim -- stale values included, the
 * reference's macroblock objects are never reset (utils.c:73-89). */
typedef struct hlb200_svc_base_mb {
    uint8_t flags;                /* bit 0: e_type is I_PCM / I_16X16 / I_8X8 / I_4X4 / I_BL (utils.c:1701); bit 1: HL_CODEC_264_MB_TYPE_IS_INTRA (flags_type, mb.h:322);
                                     bit 2: e_type is P_8X8 / P_8X8REF0 (mb.h:329) */
    uint8_t part_w, part_h;       /* MbPartWidth, MbPartHeight */
    uint8_t sub_w[4], sub_h[4];   /* SubMbPartWidth, SubMbPartHeight */
    int8_t pred_flag[4];          /* predFlagL0 */
    int8_t ref_idx[4];            /* refIdxL0 */
    uint8_t pad;
    int16_t mv[4][4][2];          /* mvL0 */
} hlb200_svc_base_mb_t;

/* Geometry of one enhancement-layer picture relative to its reference layer (layer.c:104-157, slice header SVC extension) */
typedef struct hlb200_svc_layer_geom {
    int32_t ref_width, ref_height;         /* RefLayerPicWidthInSamplesL / RefLayerPicHeightInSamplesL */
    int32_t scaled_width, scaled_height;   /* ScaledRefLayerPicWidthInSamplesL / ScaledRefLayerPicHeightInSamplesL */
    int32_t left_offset, top_offset;       /* ScaledRefLayerLeftOffset / ScaledRefLayerTopOffset */
    int32_t level_idc;                     /* of the SPS utils.c:989 reads */
    int32_t restricted;                    /* RestrictedSpatialResolutionChangeFlag (layer.c:143); 0 = the general case with its replacement / merging steps */
    int32_t cropping_change;               /* CroppingChangeFlag (layer.c:104); 1 is refused */
} hlb200_svc_layer_geom_t;

/* status bits of a derived picture (hlb200_dev_svc_derive_motion_batch); any bit = the picture has no reproduced reference behaviour */
enum { HLB200_SVC_DERIVE_BAD_REF = 1, HLB200_SVC_DERIVE_UNSUPPORTED = 2, HLB200_SVC_DERIVE_STALE_PARTS = 4, HLB200_SVC_DERIVE_NO_PRED_SOURCE = 8 };

/* One independent ME candidate for hlb200_me_cost (me_ds.c:527): partition rectangle inside MB (mb_x, mb_y) */
typedef struct hlb200_me_cand {
    int16_t mb_x, mb_y;     /* macroblock coordinates */
    uint8_t part_x, part_y; /* partition origin inside the MB (luma samples) */
    uint8_t part_w, part_h; /* 16,8,4 */
    int16_t mv_x, mv_y;     /* quarter-pel */
} hlb200_me_cand_t;
typedef struct hlb200_me_cost {
    int32_t dist;               /* SAD of the trial reconstruction */
    int32_t bits_rest;          /* residual bits without coeff_token */
    int32_t single_ctr;         /* sum over non-zero blocks */
    uint16_t cbp_luma4x4;       /* non-zero 4x4 blocks (bit = luma4x4BlkIdx) */
    uint8_t total_coeff[16];    /* per luma4x4BlkIdx of the partition (0 where zero / outside) */
    uint8_t trailing_ones[16];
    uint16_t pad;
} hlb200_me_cost_t;

/* ---- library / device ---- */
HLB200_API int hlb200_init(int device);            /* cudaSetDevice + sanity; HLB200_ERR_SYSTEM when no GPU */
HLB200_API int hlb200_device_count(void);
HLB200_API const char* hlb200_last_error(void);    /* thread-local text of the last CUDA failure */
HLB200_API int hlb200_version(void);

/* ---- stream context ---- */
HLB200_API int hlb200_stream_create(int width, int height, int max_refs, hlb200_ctx_t** out);
HLB200_API int hlb200_stream_destroy(hlb200_ctx_t* ctx);
HLB200_API int hlb200_stream_set_cuda_stream(hlb200_ctx_t* ctx, void* cuda_stream); /* run on a caller-owned cudaStream_t */
HLB200_API int hlb200_stream_sync(hlb200_ctx_t* ctx);
/* source frame (hl_frame_video_t::data_ptr[0..2], include/hartallo/hl_frame.h:28-41) -> device */
HLB200_API int hlb200_frame_upload(hlb200_ctx_t* ctx, const uint8_t* y, const uint8_t* u, const uint8_t* v, int stride_y, int stride_c);
/* source frame already resident on the device (tight planes, pitch = width, plane bases 4-byte aligned); caller-owned, must stay valid until the slice has run */
HLB200_API int hlb200_frame_set_device(hlb200_ctx_t* ctx, const uint8_t* d_y, const uint8_t* d_u, const uint8_t* d_v);
/* frame-store planes (DPB layout source/h264/hl_codec_264_dpb.c:88-166: tight Y|U|V, stride = width) */
HLB200_API int hlb200_slot_upload(hlb200_ctx_t* ctx, int slot, const uint8_t* y, const uint8_t* u, const uint8_t* v);
HLB200_API int hlb200_slot_download(hlb200_ctx_t* ctx, int slot, uint8_t* y, uint8_t* u, uint8_t* v);
HLB200_API int hlb200_state_reset(hlb200_ctx_t* ctx); /* forget the per-MB state carried across frames (new stream) */

/* ---- the hot path: one slice (= one picture) ---- */
HLB200_API int hlb200_slice_encode(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params, hlb200_mb_record_t* out_records /* PicSizeInMbs, host */);
/* asynchronous form: records stay on the device until fetched */
HLB200_API int hlb200_slice_encode_async(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params);
HLB200_API int hlb200_records_download(hlb200_ctx_t* ctx, hlb200_mb_record_t* out_records);
/* one picture of each of n independent streams in ONE launch (the contexts may differ in size); this is how a B200 is kept busy:
 * a single 1080p picture exposes at most 60 macroblocks of wavefront parallelism */
HLB200_API int hlb200_slice_encode_batch_async(hlb200_ctx_t** ctxs, const hlb200_slice_params_t* params, int n);
/* watchdog words of the last slice launch: out16[3] != 0 means a wait inside the kernel gave up (HLB200_ERR_INVALID_STATE) */
HLB200_API int hlb200_slice_status(hlb200_ctx_t* ctx, int* out16);
HLB200_API int hlb200_slice_grid_size(void); /* CTAs the slice kernel variant last launched (or selected) keeps resident on the current device */
/* slice kernel variant: 0 = one CTA per macroblock (lowest latency per picture), 1 = one warp per macroblock (highest throughput for
 * large batches), -1 = chosen per launch from the batch size (default).  Results are identical.  Returns the previous setting.
 * The environment variable HLB200_SLICE_KERNEL=cta|warp sets the initial value. */
HLB200_API int hlb200_slice_set_variant(int variant);
HLB200_API int hlb200_slice_last_variant(void);   /* variant the most recent slice launch used (0 / 1) */

/* ---- SVC enhancement layers (currDQId > 0): one picture of one layer; the context IS the layer (created with the layer's size; source picture uploaded with
 * hlb200_frame_upload; frame stores = the layer's own reference pictures).  Replaces, for all macroblocks of the picture at once, the part of
 * hl_codec_264_rdo_mb_guess_best_inter_pred_svc (rdo.c:1273; ref_slot >= 0, `motion` = what the inter-layer derivation inferred) or of
 * hl_codec_264_rdo_mb_guess_best_intra_pred_svc (rdo.c:301; ref_slot < 0, pred_* = host-resampled base layer) that follows the derivation.  The reconstruction
 * goes to frame store cur_slot (fetch it with hlb200_slot_download: the next layer's resampling runs on the host); the per-macroblock state the reference
 * carries from picture to picture lives in the context (hlb200_state_reset clears it). ---- */
HLB200_API int hlb200_svc_layer_picture(hlb200_ctx_t* ctx, int ref_slot, int cur_slot, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* motion,
                                        const uint8_t* pred_y, const uint8_t* pred_u, const uint8_t* pred_v, hlb200_mb_coeffs_t* out_coeffs);
/* The P picture of an enhancement layer with the inter-layer motion derivation on the device as well (second half; replaces the per-macroblock calls of
 * hl_codec_264_utils_derivation_process_initialisation_svc utils.c:1225 + ..._for_mv_comps_and_ref_indices_svc utils.c:1498 that rdo.c:1318-1346 makes before it predicts):
 * `base` = the reference layer's macroblock fields (host array, (ref_width / 16) x (ref_height / 16) entries).  k_svc_derive turns them into the layer's motion field on the
 * device, k_svc_inter_recon codes the picture with it.  out_motion (optional) receives the derived field, *out_status the HLB200_SVC_DERIVE_* bits; when any is set the
 * call returns HLB200_ERR_NOT_IMPLEMENTED and cur_slot / out_coeffs hold nothing usable.  The context keeps, per macroblock, whether an earlier picture left partitions
 * in the reference's macroblock object (hlb200_state_reset clears it). */
HLB200_API int hlb200_svc_layer_picture_derived(hlb200_ctx_t* ctx, int ref_slot, int cur_slot, int qp, int chroma_qp_index_offset, const hlb200_svc_base_mb_t* base,
                                                const hlb200_svc_layer_geom_t* geom, hlb200_mb_motion_t* out_motion, int32_t* out_status, hlb200_mb_coeffs_t* out_coeffs);
/* The same for an I picture with the Intra_Base resampling on the device too (hlb200_dev_svc_resample_intra_batch below; decode_svc.c:2864-3200): the caller passes
 * the reference layer's reconstruction (host planes, ref_width x ref_height) instead of full-size prediction planes.  Restrictions of that entry point apply
 * (no cropping offsets, the reference's chroma phases, no power-of-two reference dimension when level_idc > 30): HLB200_ERR_INVALID_PARAMETER otherwise. */
HLB200_API int hlb200_svc_layer_picture_resampled(hlb200_ctx_t* ctx, int cur_slot, int qp, int chroma_qp_index_offset, const uint8_t* ref_y, const uint8_t* ref_u,
                                                  const uint8_t* ref_v, int ref_width, int ref_height, int level_idc, hlb200_mb_coeffs_t* out_coeffs);

/* ---- whole-frame batch kernels, host buffers (copies inside) ---- */
/* ... and with the reference layer's reconstruction read from frame store `ref_ctx_slot` of the lower layer's context `ref_ctx` (its width x height are the reference
 * layer's), on the same GPU or another one: the hand-off between the layers of an access unit stays on the device(s) -- cudaMemcpyPeerAsync between GPUs.
 * Replaces the same call sites as hlb200_svc_layer_picture_resampled (rdo.c:363-377 -> decode_svc.c:2864). */
HLB200_API int hlb200_svc_layer_picture_resampled_from(hlb200_ctx_t* ctx, int cur_slot, int qp, int chroma_qp_index_offset, hlb200_ctx_t* ref_ctx, int ref_ctx_slot,
                                                       int level_idc, hlb200_mb_coeffs_t* out_coeffs);
HLB200_API int hlb200_interp_luma(hlb200_ctx_t* ctx, int ref_slot, const hlb200_mb_motion_t* motion, uint8_t* pred_y);
HLB200_API int hlb200_interp_chroma(hlb200_ctx_t* ctx, int ref_slot, const hlb200_mb_motion_t* motion, uint8_t* pred_u, uint8_t* pred_v);
HLB200_API int hlb200_tq_recon(hlb200_ctx_t* ctx, int qp, int chroma_qp_index_offset, const uint8_t* pred_y, const uint8_t* pred_u, const uint8_t* pred_v,
                               hlb200_mb_coeffs_t* coeffs, uint8_t* recon_y, uint8_t* recon_u, uint8_t* recon_v);
/* use_satd: 0 SAD (hl_math.c:239), 1 SATD (hl_math.c:283), 2 SSD (hl_math_ssd4x4_u8 hl_math.c:360) */
HLB200_API int hlb200_sad4x4(hlb200_ctx_t* ctx, const uint8_t* pred_y, int use_satd, int32_t* out_per_blk /* (H/4)*(W/4) */);
/* hl_math_homogeneousity8x8_u8 (hl_math.c:470) of every 8x8 block of the uploaded source luma: sum |dx| + |dy| of the Sobel pair; blocks touching the
 * plane's border (whose 3x3 support leaves the plane) report -1.  Feeds the early-termination mode mask (rdo.c:889-935; inside the slice kernel: me_mode_mask). */
HLB200_API int hlb200_homogeneity8x8(hlb200_ctx_t* ctx, int32_t* out_per_blk /* (H/8)*(W/8) */);
HLB200_API int hlb200_me_cost(hlb200_ctx_t* ctx, int ref_slot, int qp, const hlb200_me_cand_t* cands, int n, hlb200_me_cost_t* out);

/* ---- the same kernels on device pointers (inputs already resident in HBM; asynchronous on `cuda_stream`) ---- */
HLB200_API int hlb200_dev_interp_luma(const uint8_t* d_ref_y, int width, int height, const hlb200_mb_motion_t* d_motion, uint8_t* d_pred_y, void* cuda_stream);
HLB200_API int hlb200_dev_interp_chroma(const uint8_t* d_ref_u, const uint8_t* d_ref_v, int width, int height, const hlb200_mb_motion_t* d_motion,
                                        uint8_t* d_pred_u, uint8_t* d_pred_v, void* cuda_stream);
HLB200_API int hlb200_dev_tq_recon(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                                   const uint8_t* d_pred_v, int width, int height, int qp, int chroma_qp_index_offset, hlb200_mb_coeffs_t* d_coeffs,
                                   uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream);
/* picture batches of the three kernels above: n_pics pictures whose planes lie frame_stride bytes apart (consecutive tight Y|U|V frames:
 * frame_stride = width * height * 3 / 2; every plane pointer addresses picture 0) and whose motion / coefficient arrays are contiguous
 * (n_pics x macroblocks).  One launch per batch -- a single 1080p picture is far too small to load HBM3e. */
HLB200_API int hlb200_dev_interp_luma_batch(const uint8_t* d_ref_y, int width, int height, int n_pics, size_t frame_stride, const hlb200_mb_motion_t* d_motion,
                                            uint8_t* d_pred_y, void* cuda_stream);
HLB200_API int hlb200_dev_interp_chroma_batch(const uint8_t* d_ref_u, const uint8_t* d_ref_v, int width, int height, int n_pics, size_t frame_stride,
                                              const hlb200_mb_motion_t* d_motion, uint8_t* d_pred_u, uint8_t* d_pred_v, void* cuda_stream);
HLB200_API int hlb200_dev_tq_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                                         const uint8_t* d_pred_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                                         hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream);
/* SVC enhancement-layer inter macroblocks with base_mode_flag = 1 -- hl_codec_264_rdo_mb_guess_best_inter_pred_svc, source/h264/hl_codec_264_rdo.c:1273-1521
 * (prediction :1340-1424, luma residual coding with the INTRA rounding offset :1428-1496, chroma :1500, CBP :1506): one fused launch over n_pics layer
 * pictures (prediction from the layer's own reference picture + residual coding + reconstruction; the prediction never goes through HBM).
 * d_motion = the partitions / motion vectors the inter-layer derivation (host, utils.c:966-2439) inferred from the base layer; d_state (in/out,
 * n_pics x macroblocks, zero for a new layer) = hlb200_svc_mb_state_t; layout of planes and arrays as for the *_batch entry points above.
 * The reference calls this function for every macroblock of an enhancement P picture, also for those whose base macroblock is intra: these arrive with
 * predFlagL0 = 0 and no partition, and the reference codes them against the prediction its scratch blocks still hold from the last macroblock that had
 * partitions (with the intra offset in the chroma DC quantisation).  Mark such a macroblock with pad[0] = 1, pad[1] | pad[2] << 8 = address of the macroblock
 * (same picture, itself not marked) whose prediction it inherits; its own part_mode / mv are not read. */
HLB200_API int hlb200_dev_svc_inter_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_ref_y, const uint8_t* d_ref_u,
                                                const uint8_t* d_ref_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                                                const hlb200_mb_motion_t* d_motion, hlb200_svc_mb_state_t* d_state, hlb200_mb_coeffs_t* d_coeffs,
                                                uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v, void* cuda_stream);
/* Inter-layer motion derivation for enhancement-layer P pictures with base_mode_flag = 1 -- hl_codec_264_utils_derivation_process_initialisation_svc
 * utils.c:1225 (G.8.6.1.1 utils.c:1677, G.8.6.1.2 utils.c:1779, G.8.6.1.3 utils.c:1986) and ..._for_mv_comps_and_ref_indices_svc utils.c:1498 (G.8.4.1), with the
 * reference-layer lookups of utils.c:966-1059 / mb.h:313-339: d_base = n_pics x reference-layer macroblocks, d_motion = n_pics x (width / 16) x (height / 16) derived
 * macroblocks (partition layout, refIdxL0, mvL0; macroblocks whose base macroblock is intra carry the address of the macroblock whose prediction they inherit in pad[],
 * as hlb200_dev_svc_inter_recon_batch reads it), d_had_parts = one byte per derived macroblock carried from picture to picture of a layer (zero for a new layer; bit 0 = the reference's
 * macroblock object holds partitions of an earlier picture, bits 1-2 = scratch of the last call),
 * d_status = one int32 per picture, HLB200_SVC_DERIVE_* bits OR-ed in (the caller zeroes it).  Frame macroblocks, EP slices, CroppingChangeFlag = 0 (1 returns
 * HLB200_ERR_NOT_IMPLEMENTED); RestrictedSpatialResolutionChangeFlag 1 and 0. */
HLB200_API int hlb200_dev_svc_derive_motion_batch(const hlb200_svc_base_mb_t* d_base, const hlb200_svc_layer_geom_t* geom, int width, int height, int n_pics,
                                                  uint8_t* d_had_parts, hlb200_mb_motion_t* d_motion, int32_t* d_status, void* cuda_stream);
/* I_BL macroblocks (enhancement-layer I pictures) -- hl_codec_264_rdo_mb_guess_best_intra_pred_svc, rdo.c:301-461: the prediction is the base-layer reconstruction
 * resampled by the host (G.8.6.2.1, _hl_codec_264_decode_svc_resample_intra_colour_comps, source/h264/hl_codec_264_decode_svc.c:216) and arrives as
 * planes; residual coding and reconstruction are those of the base-mode inter macroblock above (rdo.c:387-446). */
HLB200_API int hlb200_dev_svc_bl_recon_batch(const uint8_t* d_src_y, const uint8_t* d_src_u, const uint8_t* d_src_v, const uint8_t* d_pred_y, const uint8_t* d_pred_u,
                                             const uint8_t* d_pred_v, int width, int height, int n_pics, size_t frame_stride, int qp, int chroma_qp_index_offset,
                                             hlb200_svc_mb_state_t* d_state, hlb200_mb_coeffs_t* d_coeffs, uint8_t* d_recon_y, uint8_t* d_recon_u, uint8_t* d_recon_v,
                                             void* cuda_stream);
/* Intra_Base resampling for enhancement-layer I pictures (first half) -- _hl_codec_264_decode_svc_resample_intra_colour_comps,
 * source/h264/hl_codec_264_decode_svc.c:2864-3200 with the sample locations of utils.c:1064-1157 (G.6.3, G.8.6.2): the reference layer's reconstruction
 * (ref_width x ref_height, tight planes, n_pics pictures ref_frame_stride bytes apart) is resampled into the prediction planes of the current layer
 * (width x height, frame_stride bytes apart, multiple of 4) that hlb200_dev_svc_bl_recon_batch consumes -- the planes the reference builds macroblock by macroblock
 * on the host.  Valid for pictures whose reference-layer macroblocks are all intra (I pictures: the only case the reference's encoder uses it for), frame
 * macroblocks, no cropping offsets, chroma phases as the reference's SPS writes them (sps.c:810-813).  level_idc selects the fixed-point precision of (G-43)
 * (16 bits up to level 3.0, 31 - ceil(log2(dimension)) above); with level_idc > 30 a reference dimension (luma or chroma) that is a power of two makes the
 * reference's own int32 arithmetic overflow -- no behaviour to pin, HLB200_ERR_INVALID_PARAMETER. */
HLB200_API int hlb200_dev_svc_resample_intra_batch(const uint8_t* d_ref_y, const uint8_t* d_ref_u, const uint8_t* d_ref_v, int ref_width, int ref_height,
                                                   uint8_t* d_pred_y, uint8_t* d_pred_u, uint8_t* d_pred_v, int width, int height, int n_pics, size_t ref_frame_stride,
                                                   size_t frame_stride, int level_idc, void* cuda_stream);
HLB200_API int hlb200_dev_sad4x4(const uint8_t* d_a, const uint8_t* d_b, int width, int height, int use_satd, int32_t* d_out, void* cuda_stream);
HLB200_API int hlb200_dev_homogeneity8x8(const uint8_t* d_plane, int width, int height, int32_t* d_out, void* cuda_stream);
HLB200_API int hlb200_dev_me_cost(const uint8_t* d_src_y, const uint8_t* d_ref_y, int width, int height, int qp, const hlb200_me_cand_t* d_cands, int n,
                                  hlb200_me_cost_t* d_out, void* cuda_stream);

/* ---- measurement aid: dependency-free 32-bit integer-ALU micro-kernel (the "measured integer peak" the ME roofline is quoted against).
 * Launches `blocks` x 256 threads, each doing `iters` x 32 independent IADD3/LOP3; writes one word per thread to d_sink; *ops_out = integer
 * operations executed. */
HLB200_API int hlb200_dev_int_alu_probe(int blocks, int iters, uint32_t* d_sink, void* cuda_stream, uint64_t* ops_out);

/* page-locks / releases a caller-owned host buffer (source pictures): uploads from it become asynchronous DMA transfers */
HLB200_API int hlb200_host_register(void* p, size_t bytes);
HLB200_API int hlb200_host_unregister(void* p);

/* ---- device-side CAVLC serialisation: slice_data() of the picture each context encoded last, written on the device from its decision records.
 * Replaces the WRITING half of the reference's macroblock loop: mb_skip_run bookkeeping + _hl_codec_264_mb_write_no_pcm (source/h264/hl_codec_264_mb.c:543-860) +
 * hl_codec_264_residual_write (source/h264/hl_codec_264_residual.c:903-1094) + the VLC writers (source/h264/hl_codec_264_cavlc.c:652-836), for the syntax the
 * reference's encoder emits (Baseline, CAVLC, one slice per picture, no ref_idx).  slice_types[i]: 0 = I, 1 = P (as hlb200_slice_params_t).  One launch sequence
 * for the whole batch on the stream of ctxs[0]; same ordering rule as hlb200_slice_encode_batch_async.  The output excludes rbsp_trailing_bits(): the host
 * appends the bits to the slice header it wrote (e.g. hl_codec_264_bits_write_u, 32 bits per word) and finishes the NAL as before. */
#define HLB200_BITS_WORDS_PER_MB 448   /* capacity of the per-picture bit buffer, 32-bit words per macroblock: above the CAVLC worst case (~1.7 KB per macroblock) */
HLB200_API int hlb200_slice_bits_batch_async(hlb200_ctx_t** ctxs, const int32_t* slice_types, int n);
/* *nbits_out = length of slice_data() in bits; out_words[k] holds bits 32k .. 32k+31, bit 32k in the most significant position, host byte order.
 * HLB200_ERR_OUTOFMEMORY when the picture does not fit cap_words (nmb * HLB200_BITS_WORDS_PER_MB + 64 words always suffice). */
HLB200_API int hlb200_slice_bits_download(hlb200_ctx_t* ctx, uint32_t* out_words, size_t cap_words, uint32_t* nbits_out);

/* ---- device self-test: the packed-instruction formulations of the search's 4x4 primitives (prediction at all 16 fractional positions, trial encode at
 * QP 12..51) against the plain formulations, both run on the device over `blocks` x 64 threads of pseudo-random inputs; *mismatches_out = 0 when they agree. */
HLB200_API int hlb200_dev_selftest(int blocks, unsigned seed, int* mismatches_out);

/* ---- TMA probe: fetches the 64x40 tile at (x0 rounded down to a multiple of 16, y0) of a device plane (width multiple of 16) with one cp.async.bulk.tensor load, as the slice kernel does, and
 * compares it with plain loads (out-of-picture samples must arrive as zeros).  mode 0: descriptor in global memory; 1: the same after a tensormap-proxy acquire
 * fence; 2: descriptor passed as a __grid_constant__ kernel parameter. */
HLB200_API int hlb200_dev_tma_probe(const uint8_t* d_plane, int width, int height, int x0, int y0, int mode, int* mismatches_out);

#ifdef __cplusplus
}
#endif
#endif /* HLB200_H_ */
