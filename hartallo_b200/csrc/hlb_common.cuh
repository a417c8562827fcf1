// hlb_common.cuh -- shared host/device declarations of the B200 library (context, error mapping, partition geometry)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hlb200.h"
#include "hlb_prims.cuh"

namespace hlb {

void set_last_error(const char* what, cudaError_t e, const char* file, int line);
extern thread_local char g_err[512];

#define HLB_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) {                                           \
            ::hlb::set_last_error(#expr, _e, __FILE__, __LINE__);          \
            return _e == cudaErrorMemoryAllocation ? HLB200_ERR_OUTOFMEMORY : HLB200_ERR_SYSTEM; \
        }                                                                  \
    } while (0)

// Geometry of the partition that contains luma position (bx,by) of a macroblock (6.4.2.1 / 6.4.2.2; partition
// tables source/h264/hl_codec_264_rdo.c:711-809).
struct PartGeom {
    int part, sub;  // mbPartIdx, subMbPartIdx
    int ox, oy;     // partition origin inside the MB
    int w, h;       // partition size
};
HLB_HD PartGeom part_of(int part_mode, const uint8_t sub_mode[4], int bx, int by)
{
    PartGeom g;
    g.sub = 0;
    switch (part_mode) {
    case 0: g.part = 0; g.ox = 0; g.oy = 0; g.w = 16; g.h = 16; break;
    case 1: g.part = by >> 3; g.ox = 0; g.oy = g.part * 8; g.w = 16; g.h = 8; break;
    case 2: g.part = bx >> 3; g.ox = g.part * 8; g.oy = 0; g.w = 8; g.h = 16; break;
    default: {
        g.part = ((by >> 3) << 1) | (bx >> 3);
        const int px = (g.part & 1) * 8, py = (g.part >> 1) * 8, lx = bx & 7, ly = by & 7;
        switch (sub_mode[g.part]) {
        case 0: g.sub = 0; g.ox = px; g.oy = py; g.w = 8; g.h = 8; break;
        case 1: g.sub = ly >> 2; g.ox = px; g.oy = py + g.sub * 4; g.w = 8; g.h = 4; break;
        case 2: g.sub = lx >> 2; g.ox = px + g.sub * 4; g.oy = py; g.w = 4; g.h = 8; break;
        default: g.sub = ((ly >> 2) << 1) | (lx >> 2); g.ox = px + (g.sub & 1) * 4; g.oy = py + (g.sub >> 1) * 4; g.w = 4; g.h = 4; break;
        }
    }
    }
    return g;
}

// Index division by a per-launch constant (macroblocks per row, words per row) as ONE multiply-high: rcp = ceil(2^32 / d) gives the exact quotient while
// n * (rcp * d - 2^32) < 2^32, which holds for every index of a plane up to 8192 x 8192 (n < 2^22, d <= 1024); 0 = "d is 1", 1 = "use the divider" (larger planes).
// The generic 32-bit division costs ~25 instructions per thread, which showed up as 4-6 % of the whole-picture kernels (profiles/r02v5_*).
inline uint32_t host_rcp32(int d, int w, int h) { return d <= 1 ? 0u : ((w > 8192 || h > 8192) ? 1u : (uint32_t)((0x100000000ull + (unsigned)d - 1) / (unsigned)d)); }
__device__ __forceinline__ int div_rcp(int n, int d, uint32_t rcp) { return rcp > 1u ? (int)__umulhi((uint32_t)n, rcp) : (rcp ? n / d : n); }

// chroma QP (8.5.8; source/h264/hl_codec_264_mb.c:375-418)
HLB_HD int chroma_qp(int qp_y, int offset) { return kQpc[clip3(0, 51, qp_y + offset)]; }
// host-side twin (the device tables are not readable from host code)
inline int host_chroma_qp(int qp_y, int offset)
{
    int q = qp_y + offset;
    q = q < 0 ? 0 : (q > 51 ? 51 : q);
    static const unsigned char t[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
    return q < 30 ? q : t[q - 30];
}

}  // namespace hlb

// context (opaque to C callers)
struct hlb200_ctx {
    int width, height, mbw, mbh, nmb;
    int max_refs, nslots;
    cudaStream_t stream;
    bool own_stream;
    uint8_t* d_src[3];                          // source frame planes (owned)
    const uint8_t* d_src_cur[3];                // planes the next slice reads: d_src, or caller-owned device planes (hlb200_frame_set_device)
    uint8_t* d_slot[HLB200_MAX_REFS + 1][3];    // frame stores: tight planes, pitch = width (dpb.c:88-166)
    uint8_t* d_pred[3];                         // scratch prediction planes (batch kernels)
    uint8_t* d_tmp[3];                          // scratch output planes (batch kernels)
    void* d_scratch; size_t scratch_bytes;      // generic device scratch (motion fields, coeffs, candidates)
    void* h_pinned; size_t pinned_bytes;        // pinned staging
    struct hlb200_mb_record* d_records;
    void* d_mbstate;                            // per-MB state carried across MBs and frames
    void* d_tmaps;                              // CUtensorMap[nslots] (device): 2D tile descriptors of the frame stores' luma planes for the slice kernel's TMA loads
    void* d_svc_state;                          // hlb200_svc_mb_state_t[nmb] when the context is an SVC enhancement layer (allocated on first use)
    uint8_t* d_svc_had_parts;                   // SVC enhancement layer with the motion derivation on the device: one byte per macroblock carried from picture to picture (hlb_svc_derive.cuh)
    void* d_sched; size_t sched_bytes;          // job descriptors + ready-queue scheduler words (hlb_slice.cu)
    void* h_jobs; int h_jobs_cap;               // pinned staging of the job descriptors
    int* last_sched;                            // scheduler words of the last launch (watchdog status)
    struct hlb200_ctx* batch_owner;             // first context of the batch launch that last covered this context (owns the scheduler words)
    int abort_state;                            // owner: -1 = not read back yet, 0 = the launch completed, > 0 = watchdog code
    cudaEvent_t ev_jobs, ev_done;               // owner: job-array upload / launch completion; other contexts: ev_done orders their stream before the launch
    void* d_bits; int bits_cap_words;           // device CAVLC output of the last picture: words | per-macroblock lengths / offsets | header (hlb_slice.cu)
    void* d_bits_jobs; void* h_bits_jobs; int bits_jobs_cap; cudaEvent_t ev_bits;   // owner of a serialisation batch: job descriptors
    void* d_dbk_bs;                             // deblocking: 32 boundary-strength bytes per macroblock (allocated on first use, hlb_deblock.cuh)
    int frame_count;
    int device;                                 // CUDA device the context's memory and stream live on (the device current when hlb200_stream_create ran)
};
