// hlb_api.cu -- C-ABI plumbing of libhl_b200.so: device selection, stream contexts, frame-store slots, host-buffer
// wrappers around the batch kernels.  No CPU fallback anywhere: a failing CUDA call is reported as an HL_ERROR_T value.
#include <string.h>

#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is resolved through the runtime, libcuda is not linked

#include "hlb_common.cuh"

namespace hlb {
thread_local char g_err[512] = "";
void set_last_error(const char* what, cudaError_t e, const char* file, int line)
{
    snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    cudaGetLastError();  // clear the sticky-less error state
}
int launch_me_cost(const uint8_t* d_src, const uint8_t* d_ref, int W, int H, int qp, const hlb200_me_cand_t* d_cands, int n, hlb200_me_cost_t* d_out, cudaStream_t st);
size_t mbstate_bytes(int nmb);
int slice_reset_state(hlb200_ctx* ctx);
}  // namespace hlb
using namespace hlb;

static int ensure_scratch(hlb200_ctx* c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return HLB200_OK;
    if (c->d_scratch) HLB_CUDA(cudaFree(c->d_scratch));
    c->d_scratch = nullptr; c->scratch_bytes = 0;
    HLB_CUDA(cudaMalloc(&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    return HLB200_OK;
}
// TMA descriptors of the frame stores' luma planes: u8 tensor {W, H}, row pitch W, box HLB tile side x tile side, out-of-picture samples read as zero
// (the kernel replicates the border samples itself, hlb_mbcore.cuh: phase_tile_load)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
namespace hlb {
int encode_tile_map(void* out128, const uint8_t* d_plane, int width, int height)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    HLB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled is not available in this driver"); return HLB200_ERR_SYSTEM; }
    const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)height}, strides[1] = {(cuuint64_t)width};
    const cuuint32_t box[2] = {64, 40}, estr[2] = {1, 1};   // HLB_TILE_W x HLB_TILE_H of hlb_mbcore.cuh
    const CUresult r = ((tmap_encode_fn)fn)((CUtensorMap*)out128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)d_plane, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d) for a %dx%d plane", (int)r, width, height); return HLB200_ERR_SYSTEM; }
    return HLB200_OK;
}
}  // namespace hlb
static int make_tile_maps(hlb200_ctx* c)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    HLB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { snprintf(hlb::g_err, sizeof(hlb::g_err), "cuTensorMapEncodeTiled is not available in this driver"); return HLB200_ERR_SYSTEM; }
    CUtensorMap maps[HLB200_MAX_REFS + 1];
    for (int s = 0; s < c->nslots; ++s) {
        const cuuint64_t dims[2] = {(cuuint64_t)c->width, (cuuint64_t)c->height}, strides[1] = {(cuuint64_t)c->width};
        const cuuint32_t box[2] = {64, 40}, estr[2] = {1, 1};   // HLB_TILE_W x HLB_TILE_H of hlb_mbcore.cuh
        const CUresult r = ((tmap_encode_fn)fn)(&maps[s], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->d_slot[s][0], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(hlb::g_err, sizeof(hlb::g_err), "cuTensorMapEncodeTiled failed (%d) for a %dx%d plane", (int)r, c->width, c->height); return HLB200_ERR_SYSTEM; }
    }
    HLB_CUDA(cudaMalloc(&c->d_tmaps, sizeof(CUtensorMap) * (size_t)c->nslots));
    HLB_CUDA(cudaMemcpy(c->d_tmaps, maps, sizeof(CUtensorMap) * (size_t)c->nslots, cudaMemcpyHostToDevice));
    return HLB200_OK;
}
static size_t plane_bytes(const hlb200_ctx* c, int p) { return p == 0 ? (size_t)c->width * c->height : (size_t)(c->width >> 1) * (c->height >> 1); }

// host -> device through the pinned staging buffer (so the copy is truly asynchronous and ordered on the stream)
static int h2d(hlb200_ctx* c, void* d, const void* h, size_t bytes)
{
    HLB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return HLB200_OK;
}
static int d2h(hlb200_ctx* c, void* h, const void* d, size_t bytes)
{
    HLB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    return HLB200_OK;
}

extern "C" {

int hlb200_version(void) { return 100; }
const char* hlb200_last_error(void) { return g_err; }

int hlb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hlb200_init(int device)
{
    int n = 0;
    HLB_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0 || device < 0 || device >= n) { snprintf(g_err, sizeof(g_err), "no CUDA device %d (count %d): this library has no CPU fallback", device, n); return HLB200_ERR_SYSTEM; }
    HLB_CUDA(cudaSetDevice(device));
    HLB_CUDA(cudaFree(0));
    return HLB200_OK;
}

// page-locks a caller-owned host buffer so that hlb200_frame_upload from it is a true asynchronous DMA (a host program in C has no other way to ask the CUDA runtime)
int hlb200_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) return HLB200_ERR_INVALID_PARAMETER;
    HLB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return HLB200_OK;
}
int hlb200_host_unregister(void* p)
{
    if (!p) return HLB200_ERR_INVALID_PARAMETER;
    HLB_CUDA(cudaHostUnregister(p));
    return HLB200_OK;
}

int hlb200_stream_create(int width, int height, int max_refs, hlb200_ctx_t** out)
{
    if (!out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || max_refs < 1 || max_refs > HLB200_MAX_REFS) return HLB200_ERR_INVALID_PARAMETER;
    hlb200_ctx* c = new hlb200_ctx();
    memset(c, 0, sizeof(*c));
    c->width = width; c->height = height; c->mbw = width >> 4; c->mbh = height >> 4; c->nmb = c->mbw * c->mbh;
    c->max_refs = max_refs; c->nslots = max_refs + 1;
    *out = c;
    HLB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    HLB_CUDA(cudaGetDevice(&c->device));
    for (int p = 0; p < 3; ++p) {
        HLB_CUDA(cudaMalloc(&c->d_src[p], plane_bytes(c, p)));
        c->d_src_cur[p] = c->d_src[p];
        HLB_CUDA(cudaMalloc(&c->d_pred[p], plane_bytes(c, p)));
        HLB_CUDA(cudaMalloc(&c->d_tmp[p], plane_bytes(c, p)));
        for (int s = 0; s < c->nslots; ++s) {
            HLB_CUDA(cudaMalloc(&c->d_slot[s][p], plane_bytes(c, p)));
            HLB_CUDA(cudaMemsetAsync(c->d_slot[s][p], 0, plane_bytes(c, p), c->stream));
        }
    }
    HLB_CUDA(cudaMalloc(&c->d_records, sizeof(hlb200_mb_record_t) * c->nmb));
    HLB_CUDA(cudaMalloc(&c->d_mbstate, mbstate_bytes(c->nmb)));
    int rc = slice_reset_state(c);
    if (rc) return rc;
    if ((rc = make_tile_maps(c))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_stream_destroy(hlb200_ctx_t* c)
{
    if (!c) return HLB200_ERR_INVALID_PARAMETER;
    cudaStreamSynchronize(c->stream);
    for (int p = 0; p < 3; ++p) {
        cudaFree(c->d_src[p]); cudaFree(c->d_pred[p]); cudaFree(c->d_tmp[p]);
        for (int s = 0; s < c->nslots; ++s) cudaFree(c->d_slot[s][p]);
    }
    cudaFree(c->d_records); cudaFree(c->d_mbstate); cudaFree(c->d_svc_state); cudaFree(c->d_svc_had_parts); cudaFree(c->d_tmaps); cudaFree(c->d_sched); cudaFree(c->d_scratch);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_jobs) cudaFreeHost(c->h_jobs);
    cudaFree(c->d_bits); cudaFree(c->d_bits_jobs); cudaFree(c->d_dbk_bs);
    if (c->h_bits_jobs) cudaFreeHost(c->h_bits_jobs);
    if (c->ev_bits) cudaEventDestroy(c->ev_bits);
    if (c->ev_jobs) cudaEventDestroy(c->ev_jobs);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return HLB200_OK;
}

int hlb200_stream_set_cuda_stream(hlb200_ctx_t* c, void* cuda_stream)
{
    if (!c) return HLB200_ERR_INVALID_PARAMETER;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) HLB_CUDA(cudaStreamDestroy(c->stream));
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return HLB200_OK;
}

int hlb200_stream_sync(hlb200_ctx_t* c)
{
    if (!c) return HLB200_ERR_INVALID_PARAMETER;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_frame_upload(hlb200_ctx_t* c, const uint8_t* y, const uint8_t* u, const uint8_t* v, int stride_y, int stride_c)
{
    if (!c || !y || !u || !v || stride_y < c->width || stride_c < (c->width >> 1)) return HLB200_ERR_INVALID_PARAMETER;
    const uint8_t* h[3] = {y, u, v};
    for (int p = 0; p < 3; ++p) {
        const int w = p ? c->width >> 1 : c->width, hh = p ? c->height >> 1 : c->height, st = p ? stride_c : stride_y;
        // tight planes go as ONE linear copy: a copy-engine DMA that overlaps a running slice kernel (the pitched form was observed not to, with every SM held by
        // the persistent kernel -- 32 ms of idle device per 256-stream launch)
        if (st == w) HLB_CUDA(cudaMemcpyAsync(c->d_src[p], h[p], (size_t)w * hh, cudaMemcpyHostToDevice, c->stream));
        else HLB_CUDA(cudaMemcpy2DAsync(c->d_src[p], w, h[p], st, w, hh, cudaMemcpyHostToDevice, c->stream));
        c->d_src_cur[p] = c->d_src[p];
    }
    return HLB200_OK;
}

int hlb200_frame_set_device(hlb200_ctx_t* c, const uint8_t* d_y, const uint8_t* d_u, const uint8_t* d_v)
{
    if (!c || !d_y || !d_u || !d_v) return HLB200_ERR_INVALID_PARAMETER;
    if ((((uintptr_t)d_y) | ((uintptr_t)d_u) | ((uintptr_t)d_v)) & 3) return HLB200_ERR_INVALID_PARAMETER;   // the kernels move samples as aligned 32-bit words
    c->d_src_cur[0] = d_y; c->d_src_cur[1] = d_u; c->d_src_cur[2] = d_v;
    return HLB200_OK;
}

int hlb200_slot_upload(hlb200_ctx_t* c, int slot, const uint8_t* y, const uint8_t* u, const uint8_t* v)
{
    if (!c || slot < 0 || slot >= c->nslots || !y || !u || !v) return HLB200_ERR_INVALID_PARAMETER;
    const uint8_t* h[3] = {y, u, v};
    for (int p = 0; p < 3; ++p) { int rc = h2d(c, c->d_slot[slot][p], h[p], plane_bytes(c, p)); if (rc) return rc; }
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_slot_download(hlb200_ctx_t* c, int slot, uint8_t* y, uint8_t* u, uint8_t* v)
{
    if (!c || slot < 0 || slot >= c->nslots || !y || !u || !v) return HLB200_ERR_INVALID_PARAMETER;
    uint8_t* h[3] = {y, u, v};
    for (int p = 0; p < 3; ++p) { int rc = d2h(c, h[p], c->d_slot[slot][p], plane_bytes(c, p)); if (rc) return rc; }
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_state_reset(hlb200_ctx_t* c)
{
    if (!c) return HLB200_ERR_INVALID_PARAMETER;
    if (c->d_svc_state) HLB_CUDA(cudaMemsetAsync(c->d_svc_state, 0, sizeof(hlb200_svc_mb_state_t) * c->nmb, c->stream));
    if (c->d_svc_had_parts) HLB_CUDA(cudaMemsetAsync(c->d_svc_had_parts, 0, c->nmb, c->stream));
    return slice_reset_state(c);
}

// One picture of an SVC enhancement layer (the context is the layer: its source picture was uploaded with hlb200_frame_upload, its frame stores hold
// the layer's own reference pictures).  ref_slot >= 0: P picture, base-mode inter macroblocks predicted from that slot with `motion`;
// ref_slot < 0: I picture, I_BL macroblocks predicted by the host-resampled planes pred_y/u/v.  The reconstruction is written to cur_slot.
int hlb200_svc_layer_picture(hlb200_ctx_t* c, int ref_slot, int cur_slot, int qp, int chroma_qp_index_offset, const hlb200_mb_motion_t* motion, const uint8_t* pred_y,
                             const uint8_t* pred_u, const uint8_t* pred_v, hlb200_mb_coeffs_t* out_coeffs)
{
    const bool bl = ref_slot < 0;
    if (!c || ref_slot >= c->nslots || cur_slot < 0 || cur_slot >= c->nslots || cur_slot == ref_slot || !out_coeffs || qp < 0 || qp > 51) return HLB200_ERR_INVALID_PARAMETER;
    if (bl ? (!pred_y || !pred_u || !pred_v) : !motion) return HLB200_ERR_INVALID_PARAMETER;
    const size_t mbytes = (sizeof(hlb200_mb_motion_t) * c->nmb + 255) & ~(size_t)255, cbytes = sizeof(hlb200_mb_coeffs_t) * c->nmb;
    int rc = ensure_scratch(c, mbytes + cbytes);
    if (rc) return rc;
    if (!c->d_svc_state) {
        HLB_CUDA(cudaMalloc(&c->d_svc_state, sizeof(hlb200_svc_mb_state_t) * c->nmb));
        HLB_CUDA(cudaMemsetAsync(c->d_svc_state, 0, sizeof(hlb200_svc_mb_state_t) * c->nmb, c->stream));
    }
    hlb200_mb_motion_t* d_motion = (hlb200_mb_motion_t*)c->d_scratch;
    hlb200_mb_coeffs_t* d_coeffs = (hlb200_mb_coeffs_t*)((char*)c->d_scratch + mbytes);
    if (bl) {
        const uint8_t* hp[3] = {pred_y, pred_u, pred_v};
        for (int p = 0; p < 3; ++p) if ((rc = h2d(c, c->d_pred[p], hp[p], plane_bytes(c, p)))) return rc;
        rc = hlb200_dev_svc_bl_recon_batch(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, 1, 0, qp,
                                           chroma_qp_index_offset, (hlb200_svc_mb_state_t*)c->d_svc_state, d_coeffs, c->d_slot[cur_slot][0], c->d_slot[cur_slot][1],
                                           c->d_slot[cur_slot][2], c->stream);
    } else {
        if ((rc = h2d(c, d_motion, motion, sizeof(hlb200_mb_motion_t) * c->nmb))) return rc;
        rc = hlb200_dev_svc_inter_recon_batch(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_slot[ref_slot][0], c->d_slot[ref_slot][1], c->d_slot[ref_slot][2],
                                              c->width, c->height, 1, 0, qp, chroma_qp_index_offset, d_motion, (hlb200_svc_mb_state_t*)c->d_svc_state, d_coeffs,
                                              c->d_slot[cur_slot][0], c->d_slot[cur_slot][1], c->d_slot[cur_slot][2], c->stream);
    }
    if (rc) return rc;
    if ((rc = d2h(c, out_coeffs, d_coeffs, cbytes))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

// A P picture of an SVC enhancement layer with the inter-layer motion derivation on the device as well (second half): the caller hands over the reference
// layer's macroblock fields instead of a motion field derived macroblock by macroblock on the host (rdo.c:1318-1346 -> utils.c:1225, :1498); k_svc_derive /
// k_svc_derive_inherit build the field in the context's scratch, k_svc_inter_recon codes the picture with it.  Same results as hlb200_svc_layer_picture with the host's field.
int hlb200_svc_layer_picture_derived(hlb200_ctx_t* c, int ref_slot, int cur_slot, int qp, int chroma_qp_index_offset, const hlb200_svc_base_mb_t* base,
                                     const hlb200_svc_layer_geom_t* geom, hlb200_mb_motion_t* out_motion, int32_t* out_status, hlb200_mb_coeffs_t* out_coeffs)
{
    if (!c || ref_slot < 0 || ref_slot >= c->nslots || cur_slot < 0 || cur_slot >= c->nslots || cur_slot == ref_slot || !out_coeffs || !out_status || !base || !geom || qp < 0 || qp > 51 ||
        geom->ref_width < 16 || geom->ref_height < 16 || (geom->ref_width & 15) || (geom->ref_height & 15) || geom->ref_width > 16384 || geom->ref_height > 16384)
        return HLB200_ERR_INVALID_PARAMETER;
    const size_t nref = (size_t)(geom->ref_width >> 4) * (geom->ref_height >> 4);
    const size_t mbytes = (sizeof(hlb200_mb_motion_t) * c->nmb + 255) & ~(size_t)255, cbytes = (sizeof(hlb200_mb_coeffs_t) * c->nmb + 255) & ~(size_t)255;
    const size_t bbytes = (sizeof(hlb200_svc_base_mb_t) * nref + 255) & ~(size_t)255;
    int rc = ensure_scratch(c, mbytes + cbytes + bbytes + 256);
    if (rc) return rc;
    if (!c->d_svc_state) {
        HLB_CUDA(cudaMalloc(&c->d_svc_state, sizeof(hlb200_svc_mb_state_t) * c->nmb));
        HLB_CUDA(cudaMemsetAsync(c->d_svc_state, 0, sizeof(hlb200_svc_mb_state_t) * c->nmb, c->stream));
    }
    if (!c->d_svc_had_parts) {
        HLB_CUDA(cudaMalloc((void**)&c->d_svc_had_parts, c->nmb));
        HLB_CUDA(cudaMemsetAsync(c->d_svc_had_parts, 0, c->nmb, c->stream));
    }
    char* s = (char*)c->d_scratch;
    hlb200_mb_motion_t* d_motion = (hlb200_mb_motion_t*)s;
    hlb200_mb_coeffs_t* d_coeffs = (hlb200_mb_coeffs_t*)(s + mbytes);
    hlb200_svc_base_mb_t* d_base = (hlb200_svc_base_mb_t*)(s + mbytes + cbytes);
    int32_t* d_status = (int32_t*)(s + mbytes + cbytes + bbytes);
    if ((rc = h2d(c, d_base, base, sizeof(hlb200_svc_base_mb_t) * nref))) return rc;
    HLB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t), c->stream));
    if ((rc = hlb200_dev_svc_derive_motion_batch(d_base, geom, c->width, c->height, 1, c->d_svc_had_parts, d_motion, d_status, c->stream))) return rc;
    // the status decides whether the picture may be coded at all: one small read-back before the big kernel (a refused picture leaves the frame stores and the carried
    // chroma levels alone; the per-macroblock flags are those after its derivation, as the reference's macroblock objects would be)
    int32_t st = 0;
    if ((rc = d2h(c, &st, d_status, sizeof(st)))) return rc;
    if (out_motion && (rc = d2h(c, out_motion, d_motion, sizeof(hlb200_mb_motion_t) * c->nmb))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    *out_status = st;
    if (st) {
        snprintf(g_err, sizeof(g_err), "hlb200: the derived motion of this enhancement-layer picture has no reproduced reference behaviour (HLB200_SVC_DERIVE_* status %d)", (int)st);
        return HLB200_ERR_NOT_IMPLEMENTED;
    }
    if ((rc = hlb200_dev_svc_inter_recon_batch(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_slot[ref_slot][0], c->d_slot[ref_slot][1], c->d_slot[ref_slot][2],
                                               c->width, c->height, 1, 0, qp, chroma_qp_index_offset, d_motion, (hlb200_svc_mb_state_t*)c->d_svc_state, d_coeffs,
                                               c->d_slot[cur_slot][0], c->d_slot[cur_slot][1], c->d_slot[cur_slot][2], c->stream)))
        return rc;
    if ((rc = d2h(c, out_coeffs, d_coeffs, sizeof(hlb200_mb_coeffs_t) * c->nmb))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

// An I picture of an SVC enhancement layer with the Intra_Base resampling on the device as well (first half): the caller hands over the reference
// layer's reconstruction (ref_width x ref_height, a quarter of the samples at the dyadic ratio) instead of the full-size prediction planes it would otherwise have to
// build macroblock by macroblock (rdo.c:363-377 -> decode_svc.c:2864); k_svc_resample_intra fills the context's prediction planes, k_svc_inter_recon<true> codes the
// picture against them.  Same results as hlb200_svc_layer_picture(ref_slot = -1) with host-resampled planes.
int hlb200_svc_layer_picture_resampled(hlb200_ctx_t* c, int cur_slot, int qp, int chroma_qp_index_offset, const uint8_t* ref_y, const uint8_t* ref_u, const uint8_t* ref_v,
                                       int ref_width, int ref_height, int level_idc, hlb200_mb_coeffs_t* out_coeffs)
{
    if (!c || cur_slot < 0 || cur_slot >= c->nslots || !out_coeffs || qp < 0 || qp > 51 || !ref_y || !ref_u || !ref_v || ref_width < 16 || ref_height < 16 || (ref_width & 15) ||
        (ref_height & 15) || ref_width > c->width || ref_height > c->height)
        return HLB200_ERR_INVALID_PARAMETER;
    const size_t cbytes = (sizeof(hlb200_mb_coeffs_t) * c->nmb + 255) & ~(size_t)255, rys = (size_t)ref_width * ref_height, rcs = rys >> 2;
    int rc = ensure_scratch(c, cbytes + rys + 2 * rcs);
    if (rc) return rc;
    if (!c->d_svc_state) {
        HLB_CUDA(cudaMalloc(&c->d_svc_state, sizeof(hlb200_svc_mb_state_t) * c->nmb));
        HLB_CUDA(cudaMemsetAsync(c->d_svc_state, 0, sizeof(hlb200_svc_mb_state_t) * c->nmb, c->stream));
    }
    hlb200_mb_coeffs_t* d_coeffs = (hlb200_mb_coeffs_t*)c->d_scratch;
    uint8_t* d_ref = (uint8_t*)c->d_scratch + cbytes;
    if ((rc = h2d(c, d_ref, ref_y, rys)) || (rc = h2d(c, d_ref + rys, ref_u, rcs)) || (rc = h2d(c, d_ref + rys + rcs, ref_v, rcs))) return rc;
    if ((rc = hlb200_dev_svc_resample_intra_batch(d_ref, d_ref + rys, d_ref + rys + rcs, ref_width, ref_height, c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, 1, 0, 0,
                                                  level_idc, c->stream)))
        return rc;
    if ((rc = hlb200_dev_svc_bl_recon_batch(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, 1, 0, qp,
                                            chroma_qp_index_offset, (hlb200_svc_mb_state_t*)c->d_svc_state, d_coeffs, c->d_slot[cur_slot][0], c->d_slot[cur_slot][1],
                                            c->d_slot[cur_slot][2], c->stream)))
        return rc;
    if ((rc = d2h(c, out_coeffs, d_coeffs, sizeof(hlb200_mb_coeffs_t) * c->nmb))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

// The same with the reference layer's reconstruction taken from ANOTHER CONTEXT's frame store instead of host planes (the per-access-unit hand-off between the
// layers of an SVC stream): the lower layer's context may live on another GPU, the planes then travel GPU to GPU (cudaMemcpyPeerAsync over NVLink; without peer access the
// runtime stages the copy itself).  The calling thread's current device must be ctx's.
int hlb200_svc_layer_picture_resampled_from(hlb200_ctx_t* c, int cur_slot, int qp, int chroma_qp_index_offset, hlb200_ctx_t* rc_ctx, int ref_ctx_slot, int level_idc,
                                            hlb200_mb_coeffs_t* out_coeffs)
{
    if (!c || !rc_ctx || rc_ctx == c || cur_slot < 0 || cur_slot >= c->nslots || ref_ctx_slot < 0 || ref_ctx_slot >= rc_ctx->nslots || !out_coeffs || qp < 0 || qp > 51 ||
        rc_ctx->width > c->width || rc_ctx->height > c->height)
        return HLB200_ERR_INVALID_PARAMETER;
    const int ref_width = rc_ctx->width, ref_height = rc_ctx->height;
    const size_t cbytes = (sizeof(hlb200_mb_coeffs_t) * c->nmb + 255) & ~(size_t)255, rys = (size_t)ref_width * ref_height, rcs = rys >> 2;
    int rc = ensure_scratch(c, cbytes + rys + 2 * rcs);
    if (rc) return rc;
    if (!c->d_svc_state) {
        HLB_CUDA(cudaMalloc(&c->d_svc_state, sizeof(hlb200_svc_mb_state_t) * c->nmb));
        HLB_CUDA(cudaMemsetAsync(c->d_svc_state, 0, sizeof(hlb200_svc_mb_state_t) * c->nmb, c->stream));
    }
    hlb200_mb_coeffs_t* d_coeffs = (hlb200_mb_coeffs_t*)c->d_scratch;
    uint8_t* d_ref = (uint8_t*)c->d_scratch + cbytes;
    // the lower layer's picture is complete once its stream has drained (its own picture call synchronised already; this also covers asynchronous producers)
    {
        int cur = 0;
        HLB_CUDA(cudaGetDevice(&cur));
        if (cur != rc_ctx->device) HLB_CUDA(cudaSetDevice(rc_ctx->device));
        const cudaError_t e = cudaStreamSynchronize(rc_ctx->stream);
        if (cur != rc_ctx->device) HLB_CUDA(cudaSetDevice(cur));
        HLB_CUDA(e);
    }
    const size_t off[3] = {0, rys, rys + rcs}, bytes[3] = {rys, rcs, rcs};
    for (int p = 0; p < 3; ++p) {
        if (rc_ctx->device == c->device) HLB_CUDA(cudaMemcpyAsync(d_ref + off[p], rc_ctx->d_slot[ref_ctx_slot][p], bytes[p], cudaMemcpyDeviceToDevice, c->stream));
        else HLB_CUDA(cudaMemcpyPeerAsync(d_ref + off[p], c->device, rc_ctx->d_slot[ref_ctx_slot][p], rc_ctx->device, bytes[p], c->stream));
    }
    if ((rc = hlb200_dev_svc_resample_intra_batch(d_ref, d_ref + rys, d_ref + rys + rcs, ref_width, ref_height, c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, 1, 0, 0,
                                                  level_idc, c->stream)))
        return rc;
    if ((rc = hlb200_dev_svc_bl_recon_batch(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, 1, 0, qp,
                                            chroma_qp_index_offset, (hlb200_svc_mb_state_t*)c->d_svc_state, d_coeffs, c->d_slot[cur_slot][0], c->d_slot[cur_slot][1],
                                            c->d_slot[cur_slot][2], c->stream)))
        return rc;
    if ((rc = d2h(c, out_coeffs, d_coeffs, sizeof(hlb200_mb_coeffs_t) * c->nmb))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

// ---- host-buffer batch wrappers -----------------------------------------------------------------------------------
int hlb200_interp_luma(hlb200_ctx_t* c, int ref_slot, const hlb200_mb_motion_t* motion, uint8_t* pred_y)
{
    if (!c || ref_slot < 0 || ref_slot >= c->nslots || !motion || !pred_y) return HLB200_ERR_INVALID_PARAMETER;
    int rc = ensure_scratch(c, sizeof(hlb200_mb_motion_t) * c->nmb);
    if (rc) return rc;
    if ((rc = h2d(c, c->d_scratch, motion, sizeof(hlb200_mb_motion_t) * c->nmb))) return rc;
    if ((rc = hlb200_dev_interp_luma(c->d_slot[ref_slot][0], c->width, c->height, (const hlb200_mb_motion_t*)c->d_scratch, c->d_pred[0], c->stream))) return rc;
    if ((rc = d2h(c, pred_y, c->d_pred[0], plane_bytes(c, 0)))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_interp_chroma(hlb200_ctx_t* c, int ref_slot, const hlb200_mb_motion_t* motion, uint8_t* pred_u, uint8_t* pred_v)
{
    if (!c || ref_slot < 0 || ref_slot >= c->nslots || !motion || !pred_u || !pred_v) return HLB200_ERR_INVALID_PARAMETER;
    int rc = ensure_scratch(c, sizeof(hlb200_mb_motion_t) * c->nmb);
    if (rc) return rc;
    if ((rc = h2d(c, c->d_scratch, motion, sizeof(hlb200_mb_motion_t) * c->nmb))) return rc;
    if ((rc = hlb200_dev_interp_chroma(c->d_slot[ref_slot][1], c->d_slot[ref_slot][2], c->width, c->height, (const hlb200_mb_motion_t*)c->d_scratch, c->d_pred[1],
                                       c->d_pred[2], c->stream)))
        return rc;
    if ((rc = d2h(c, pred_u, c->d_pred[1], plane_bytes(c, 1)))) return rc;
    if ((rc = d2h(c, pred_v, c->d_pred[2], plane_bytes(c, 2)))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_tq_recon(hlb200_ctx_t* c, int qp, int chroma_qp_index_offset, const uint8_t* pred_y, const uint8_t* pred_u, const uint8_t* pred_v,
                    hlb200_mb_coeffs_t* coeffs, uint8_t* recon_y, uint8_t* recon_u, uint8_t* recon_v)
{
    if (!c || !pred_y || !pred_u || !pred_v || !coeffs || !recon_y || !recon_u || !recon_v) return HLB200_ERR_INVALID_PARAMETER;
    int rc = ensure_scratch(c, sizeof(hlb200_mb_coeffs_t) * c->nmb);
    if (rc) return rc;
    const uint8_t* hp[3] = {pred_y, pred_u, pred_v};
    uint8_t* hr[3] = {recon_y, recon_u, recon_v};
    for (int p = 0; p < 3; ++p) if ((rc = h2d(c, c->d_pred[p], hp[p], plane_bytes(c, p)))) return rc;
    if ((rc = hlb200_dev_tq_recon(c->d_src_cur[0], c->d_src_cur[1], c->d_src_cur[2], c->d_pred[0], c->d_pred[1], c->d_pred[2], c->width, c->height, qp, chroma_qp_index_offset,
                                  (hlb200_mb_coeffs_t*)c->d_scratch, c->d_tmp[0], c->d_tmp[1], c->d_tmp[2], c->stream)))
        return rc;
    if ((rc = d2h(c, coeffs, c->d_scratch, sizeof(hlb200_mb_coeffs_t) * c->nmb))) return rc;
    for (int p = 0; p < 3; ++p) if ((rc = d2h(c, hr[p], c->d_tmp[p], plane_bytes(c, p)))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_sad4x4(hlb200_ctx_t* c, const uint8_t* pred_y, int use_satd, int32_t* out_per_blk)
{
    if (!c || !pred_y || !out_per_blk) return HLB200_ERR_INVALID_PARAMETER;
    const size_t nb = (size_t)(c->width >> 2) * (c->height >> 2);
    int rc = ensure_scratch(c, nb * sizeof(int32_t));
    if (rc) return rc;
    if ((rc = h2d(c, c->d_pred[0], pred_y, plane_bytes(c, 0)))) return rc;
    if ((rc = hlb200_dev_sad4x4(c->d_src_cur[0], c->d_pred[0], c->width, c->height, use_satd, (int32_t*)c->d_scratch, c->stream))) return rc;
    if ((rc = d2h(c, out_per_blk, c->d_scratch, nb * sizeof(int32_t)))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_homogeneity8x8(hlb200_ctx_t* c, int32_t* out_per_blk)
{
    if (!c || !out_per_blk) return HLB200_ERR_INVALID_PARAMETER;
    const size_t nb = (size_t)(c->width >> 3) * (c->height >> 3);
    int rc = ensure_scratch(c, nb * sizeof(int32_t));
    if (rc) return rc;
    if ((rc = hlb200_dev_homogeneity8x8(c->d_src_cur[0], c->width, c->height, (int32_t*)c->d_scratch, c->stream))) return rc;
    if ((rc = d2h(c, out_per_blk, c->d_scratch, nb * sizeof(int32_t)))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

int hlb200_me_cost(hlb200_ctx_t* c, int ref_slot, int qp, const hlb200_me_cand_t* cands, int n, hlb200_me_cost_t* out)
{
    if (!c || ref_slot < 0 || ref_slot >= c->nslots || !cands || !out || n < 0 || qp < 0 || qp > 51) return HLB200_ERR_INVALID_PARAMETER;
    if (n == 0) return HLB200_OK;
    for (int i = 0; i < n; ++i) {
        const hlb200_me_cand_t& d = cands[i];
        if (d.mb_x < 0 || d.mb_x >= c->mbw || d.mb_y < 0 || d.mb_y >= c->mbh || (d.part_w != 4 && d.part_w != 8 && d.part_w != 16) ||
            (d.part_h != 4 && d.part_h != 8 && d.part_h != 16) || d.part_x + d.part_w > 16 || d.part_y + d.part_h > 16 || (d.part_x & 3) || (d.part_y & 3))
            return HLB200_ERR_INVALID_PARAMETER;
    }
    const size_t in_b = sizeof(hlb200_me_cand_t) * (size_t)n, in_pad = (in_b + 255) & ~(size_t)255, out_b = sizeof(hlb200_me_cost_t) * (size_t)n;
    int rc = ensure_scratch(c, in_pad + out_b);
    if (rc) return rc;
    hlb200_me_cand_t* d_c = (hlb200_me_cand_t*)c->d_scratch;
    hlb200_me_cost_t* d_o = (hlb200_me_cost_t*)((char*)c->d_scratch + in_pad);
    if ((rc = h2d(c, d_c, cands, in_b))) return rc;
    if ((rc = launch_me_cost(c->d_src_cur[0], c->d_slot[ref_slot][0], c->width, c->height, qp, d_c, n, d_o, c->stream))) return rc;
    if ((rc = d2h(c, out, d_o, out_b))) return rc;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return HLB200_OK;
}

}  // extern "C"
