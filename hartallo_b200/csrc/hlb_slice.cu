// hlb_slice.cu -- slice-level hot path (ME + mode decision + reconstruction), wavefront-scheduled.  (under construction)
#include "hlb_common.cuh"

namespace hlb {
size_t mbstate_bytes(int nmb) { return (size_t)nmb * 1024; }
int slice_reset_state(hlb200_ctx* c)
{
    HLB_CUDA(cudaMemsetAsync(c->d_mbstate, 0, mbstate_bytes(c->nmb), c->stream));
    c->frame_count = 0;
    return HLB200_OK;
}
}  // namespace hlb

extern "C" {
int hlb200_slice_encode_async(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params) { (void)ctx; (void)params; return HLB200_ERR_NOT_IMPLEMENTED; }
int hlb200_records_download(hlb200_ctx_t* ctx, hlb200_mb_record_t* out_records) { (void)ctx; (void)out_records; return HLB200_ERR_NOT_IMPLEMENTED; }
int hlb200_slice_encode(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params, hlb200_mb_record_t* out_records)
{
    int rc = hlb200_slice_encode_async(ctx, params);
    if (rc) return rc;
    return hlb200_records_download(ctx, out_records);
}
}
