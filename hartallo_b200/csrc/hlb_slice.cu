// hlb_slice.cu -- slice-level hot path on the device: hlb200_slice_encode = the per-macroblock decide + reconstruct loop of
// hl_codec_264_nal_slice_data_encode (source/h264/hl_codec_264_slice.c:1786-1894) for one picture of one or many streams.
//
// Scheduling.  A macroblock may start when its neighbours A, B, C, D are final: the 2:1 wavefront.  One persistent
// kernel serves every picture of a batch (independent streams): CTAs pop ready macroblocks from a global queue; a finished
// macroblock decrements the dependency counters of its successors (right neighbour; the MB whose top-right it is) and pushes
// those that reach zero.  A CTA therefore never holds a blocked macroblock, so the scheme cannot deadlock whatever the grid
// size.  I pictures add the raster predecessor as a dependency (the Single_ctr chain of residual.c:882, see hlb_mbintra.cuh);
// in P pictures that chain is consumed almost never and is fetched lazily by waiting for the predecessor's done flag.
//
// One CTA per macroblock, warp-specialised: warp 0 (the master warp) runs the serial control flow of hlb_mbcore.cuh -- all 32 lanes
// redundantly on identical data, so the warp stays converged -- and posts commands; the worker warps (HLB_WORKERS threads; a command's logical lanes are strided
// over them) execute the command phases.  Master and workers meet at named barrier 1 from different code paths, which is legal for bar.sync as long
// as every WARP is converged at its own call site (bar.sync == barrier.sync.aligned).  Global loads are compiled .cg (-Xptxas -dlcm=cg) so that state and
// reconstruction written by other CTAs is read from L2; read-only planes go through __ldg.
#include <new>
#include <stdlib.h>
#include <string.h>

// ONE out-of-line copy of the packed luma prediction for all its callers (trial encodes, final prediction): the slice kernels are instruction-fetch bound
// (DESIGN.md 4.1), and the four packed rows travel in registers
#define HLB_FASTPRED_FN __device__ __noinline__
#define HLB_FASTPRED_SRC(t) __builtin_assume(__isShared(t))
#include "hlb_common.cuh"
#include "hlb_mbcore.cuh"
#include "hlb_bits.cuh"
#include "hlb_deblock.cuh"

namespace hlb {

struct SliceJob {
    FrameCtx f;
    DbkJob d;            // loop filter over the finished picture (d.enabled; hlb_deblock.cuh)
    int nmb;
    int base;            // first item index of this job in the scheduler arrays
    int prev_frame_sctr; // Single_ctr chain value entering the picture
    int* chain;          // per-stream carrier of that value across pictures (device)
};

struct Sched {
    int head, tail, total, abort;   // abort != 0: a watchdog fired (code below); every spin loop gives up and the kernel drains
    int dbg[12];                    // [0] code, [1] item, [2] queue index / neighbour, [3] blockIdx, [4..] spare
};
#define HLB_SCHED_WORDS 16
enum { WD_QUEUE = 1, WD_PREV = 2, WD_SEARCH = 3 };
// Longest sleep of an idle CTA between two looks at its queue slot.  Idle pollers are not free: every wake-up refetches its loop and
// competes with the working warps of the SM for the instruction caches.  Measured (r01d): the warp variant at 256 pictures per launch
// runs ~1,000 of its 2,368 warps idle; 2 us -> 131 us of back-off raised throughput 2.26 M -> 2.59 M MB/s (512 us: 2.51 M).  The CTA
// variant serves small batches where pick-up latency sits on the critical path of a picture: 8 us there (32 us cost 1.5 % at 32 pictures).
#ifndef HLB_SPIN_MAX_NS_CTA
#define HLB_SPIN_MAX_NS_CTA 8192
#endif
#ifndef HLB_SPIN_MAX_NS_WARP
#define HLB_SPIN_MAX_NS_WARP 131072
#endif
#define HLB_SPIN_LIMIT (1 << 22)    // x ~200 ns sleep: about a second of waiting before a spin loop declares the kernel stuck
// layout of the scheduler buffer: Sched | queue[total] | deps[total] | done[total]
__device__ __forceinline__ void watchdog_fire(Sched* s, int code, int a, int b)
{
    if (atomicCAS(&s->abort, 0, code) == 0) { s->dbg[0] = code; s->dbg[1] = a; s->dbg[2] = b; s->dbg[3] = (int)blockIdx.x; __threadfence(); }
}

#ifndef HLB_WORKERS
#define HLB_WORKERS 64            /* worker threads per CTA; a command's logical lanes are strided over them */
#endif
#define HLB_CTA_THREADS (HLB_WORKERS + 32)
#ifndef HLB_SLICE_MIN_CTAS
#define HLB_SLICE_MIN_CTAS 6   /* register budget: 65536 / (6 x 96) = 113 -> ptxas settles at 96; measured best of {3,6,8} x {32,64,128} workers (r01c sweep) */
#endif
__device__ __forceinline__ void cta_bar() { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"n"(HLB_CTA_THREADS) : "memory"); }
__device__ __forceinline__ int ld_volatile(const int* p) { return *(const volatile int*)p; }

struct GpuExec {
    MbWork* w;
    const FrameCtx* f;
    const SliceJob* job;
    const int* done;
    Sched* sched;
    __device__ __noinline__ void run(int cmd, int nlanes)
    {
#ifdef HLB_PROFILE_STEPS
        const long long t0 = clock64();
#endif
        w->arg0_lanes = nlanes;
        ((volatile int*)&w->cmd)[0] = cmd;
        cta_bar();
        const int np = cmd_phases(cmd);
        for (int p = 0; p < np; ++p) cta_bar();   // the worker warps run the phases
#ifdef HLB_PROFILE_STEPS
        w->prof_run_cycles += (unsigned)(clock64() - t0); w->prof_runs++;
        if (cmd == CMD_ME_EVAL) { w->prof_me_cycles += (unsigned)(clock64() - t0); }
#endif
    }
    __device__ __forceinline__ void trials(int nlanes) { run(CMD_ME_EVAL, nlanes); }   // the worker warps run the trial encodes of a search step
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int nlanes() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // cross-lane reductions of the (master) warp; every lane receives the result
    __device__ __forceinline__ int reduce_min(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_add(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    // best-MV argmin: the smallest cost wins, the smaller candidate index among equal costs (warp-shuffle butterfly)
    __device__ __forceinline__ void reduce_argmin(double& c, int& i) const
    {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, c, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (oc < c || (oc == c && oi < i)) { c = oc; i = oi; }
        }
    }
    __device__ __noinline__ int prev_sctr(int mb)
    {
        for (int a = mb - 1; a >= 0; --a) {
            int spins = 0;
            while (ld_volatile(done + job->base + a) == 0) {
                if (ld_volatile(&sched->abort)) return 0;
                if (++spins > HLB_SPIN_LIMIT) { watchdog_fire(sched, WD_PREV, job->base + mb, a); return 0; }
                __nanosleep(200);
            }
            __threadfence();
            const int v = *(volatile const uint8_t*)&f->st[a].last_sctr;
            if (v != 255) return v;
        }
        return job->prev_frame_sctr;
    }
};

// ---- loop filter of the pictures of a launch (hlb_deblock.cuh) -------------------------------------------------------------------------------------
// boundary strengths: they depend on the decision records only, one thread per value.  blockIdx.y = picture
__global__ void __launch_bounds__(256) k_dbk_bs(const SliceJob* __restrict__ jobs)
{
    const DbkJob& j = jobs[blockIdx.y].d;
    if (!j.enabled) return;
    const int t = blockIdx.x * 256 + threadIdx.x, mb = t >> 5;
    if (mb >= j.mbw * j.mbh) return;
    j.bs[t] = (uint8_t)dbk_bs_mb(j, mb % j.mbw, mb / j.mbw, t & 31);
}
// One CTA per picture, one warp per macroblock row in flight: row y filters macroblock x once row y-1 has finished macroblock x+1 (the raster order of 8.7
// only orders a macroblock after its left, top and top-right neighbours, whose samples it reads or rewrites).  Progress counters live in shared memory, every
// warp of the picture is resident in the same CTA, so the waits cannot deadlock; the samples travel through the L1 of the one SM all of them run on.
#define HLB_DBK_WARPS 16
__global__ void __launch_bounds__(HLB_DBK_WARPS * 32) k_dbk(const SliceJob* __restrict__ jobs)
{
    extern __shared__ int s_prog[];   // macroblocks finished per row
    __shared__ DbkJob j;
    __shared__ uint8_t s_bs[HLB_DBK_WARPS][32];
    if (!jobs[blockIdx.x].d.enabled) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (int)(sizeof(DbkJob) / 4); i += blockDim.x) ((uint32_t*)&j)[i] = ((const uint32_t*)&jobs[blockIdx.x].d)[i];
    for (int i = threadIdx.x; i < jobs[blockIdx.x].d.mbh; i += blockDim.x) s_prog[i] = 0;
    __syncthreads();
    volatile int* prog = s_prog;
    for (int y = warp; y < j.mbh; y += HLB_DBK_WARPS) {
        for (int x = 0; x < j.mbw; ++x) {
            if (y) {
                const int need = x + 2 < j.mbw ? x + 2 : j.mbw;
                while (prog[y - 1] < need) __nanosleep(64);
                __threadfence_block();
            }
            s_bs[warp][lane] = j.bs[((size_t)y * j.mbw + x) * 32 + lane];
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                dbk_edge(j, x, y, k >> 2, k & 3, lane, s_bs[warp]);
                __syncwarp();
            }
            __threadfence_block();
            if (lane == 0) prog[y] = x + 1;
        }
    }
}

__global__ void k_slice_init(SliceJob* jobs, int njobs, int* sched_buf, int total)
{
    Sched* s = (Sched*)sched_buf;
    int* queue = sched_buf + HLB_SCHED_WORDS;
    int* deps = queue + total;
    int* done = deps + total;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { s->head = 0; s->tail = njobs; s->total = total; s->abort = 0; for (int k = 0; k < 12; ++k) s->dbg[k] = 0; }
    if (t < njobs) {
        queue[t] = jobs[t].base;  // macroblock 0 of every picture is ready
        // Single_ctr chain entering the picture: last value any macroblock of the previous picture left (residual.c:882)
        const SliceJob& j = jobs[t];
        int v = *j.chain;
        for (int a = j.nmb - 1; a >= 0; --a)
            if (j.f.st[a].last_sctr != 255) { v = j.f.st[a].last_sctr; break; }
        jobs[t].prev_frame_sctr = v;
        *j.chain = v;
    }
    for (int i = t; i < total; i += gridDim.x * blockDim.x) {
        if (i >= njobs) queue[i] = -1;
        done[i] = 0;
        // which job / macroblock is item i
        int jb = 0;
        while (jb + 1 < njobs && jobs[jb + 1].base <= i) ++jb;
        const SliceJob& j = jobs[jb];
        const int mb = i - j.base, x = mb % j.f.mbw, y = mb / j.f.mbw;
        deps[i] = j.f.is_p ? ((x > 0) + (y > 0)) : (mb > 0);
    }
}

// ---- scheduler steps shared by both kernels (executed by ONE thread) ----
// pops the next ready macroblock: returns the item (or -1 when the batch is drained / aborted) and the job that owns it
__device__ __forceinline__ int sched_pop(Sched* s, const int* queue, int total, const SliceJob* jobs, int njobs, int* job_out, unsigned max_sleep_ns)
{
    const int idx = atomicAdd(&s->head, 1);
    int item = -1;
    if (idx < total && !ld_volatile(&s->abort)) {
        int spins = 0;
        unsigned ns = 64;   // exponential back-off: idle CTAs must not steal issue slots / L2 bandwidth from working ones
        while ((item = ld_volatile(queue + idx)) < 0) {
            if ((spins & 15) == 15 && ld_volatile(&s->abort)) break;
            if (++spins > HLB_SPIN_LIMIT / 8) { watchdog_fire(s, WD_QUEUE, idx, ld_volatile(&s->tail)); break; }
            __nanosleep(ns);
            if (ns < max_sleep_ns) ns <<= 1;
        }
        __threadfence();
    }
    if (item >= 0) {
        int lo = 0, hi = njobs - 1;   // jobs[].base is ascending
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].base <= item) lo = mid; else hi = mid - 1; }
        *job_out = lo;
    }
    return item;
}
// publishes a finished macroblock and releases its successors (see the header comment)
__device__ __forceinline__ void sched_finish(Sched* s, int* queue, int* deps, int* done, const FrameCtx& sf, int nmb, int item, int mb)
{
    __threadfence();
    atomicExch(done + item, 1);
    const int mbw = sf.mbw, mbh = sf.mbh, x = mb % mbw, y = mb / mbw;
    int succ[3], ns = 0;
    if (sf.is_p) {
        if (x + 1 < mbw) succ[ns++] = item + 1;
        if (y + 1 < mbh) {
            if (x >= 1) succ[ns++] = item + mbw - 1;          // (x-1, y+1): its top-right is this macroblock
            if (x == mbw - 1) succ[ns++] = item + mbw;        // last column waits for its top neighbour
        }
    } else if (mb + 1 < nmb) succ[ns++] = item + 1;
    for (int k = 0; k < ns; ++k)
        if (atomicSub(deps + succ[k], 1) == 1) {
            const int slot = atomicAdd(&s->tail, 1);
            atomicExch(queue + slot, succ[k]);
        }
}

__device__ __forceinline__ void tile_barrier_init(MbWork& w)
{
    // the mbarrier the TMA tile loads of this CTA complete on (one arrival = the issuing lane's expect_tx), made visible to the async proxy
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&w.tile_mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    w.tile_phase = 0;
}

__global__ void __launch_bounds__(HLB_CTA_THREADS, HLB_SLICE_MIN_CTAS) k_slice_encode(const SliceJob* __restrict__ jobs, int njobs, int* sched_buf)
{
    __shared__ MbWork w;
    __shared__ FrameCtx sf;
    __shared__ int s_item, s_job;
    Sched* s = (Sched*)sched_buf;
    const int total = s->total;
    int* queue = sched_buf + HLB_SCHED_WORDS;
    int* deps = queue + total;
    int* done = deps + total;
    const int tid = threadIdx.x;
    if (tid == 0) tile_barrier_init(w);
    __syncthreads();
    for (;;) {
        if (tid == 0) { int jb = 0; s_item = sched_pop(s, queue, total, jobs, njobs, &jb, HLB_SPIN_MAX_NS_CTA); s_job = jb; }
        __syncthreads();
        const int item = s_item;
        if (item < 0) break;
        const SliceJob* job = jobs + s_job;
        // picture context -> shared memory (read by every lane, many times)
        {
            const int* src = (const int*)&job->f;
            int* dst = (int*)&sf;
            for (int i = tid; i < (int)(sizeof(FrameCtx) / sizeof(int)); i += HLB_CTA_THREADS) dst[i] = src[i];
        }
        __syncthreads();
        const int mb = item - job->base;
        unsigned t_start = 0;
        if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); t_start = (unsigned)t; }
        if (tid < 32) {   // master warp: every lane executes the same control flow on the same data
            GpuExec x;
            x.w = &w; x.f = &sf; x.job = job; x.done = done; x.sched = s;
            mb_encode(x, w, sf, mb);
            if (w.stuck) watchdog_fire(s, WD_SEARCH, item, w.stuck);
            ((volatile int*)&w.cmd)[0] = CMD_EXIT;
            cta_bar();
        } else {
            const int wt = tid - 32;
            for (;;) {
                cta_bar();
                const int cmd = ((volatile int*)&w.cmd)[0];
                if (cmd == CMD_EXIT) break;
                const int nl = w.arg0_lanes;
                const int np = cmd_phases(cmd);
                for (int p = 0; p < np; ++p) {
                    for (int lane = wt; lane < nl; lane += HLB_WORKERS) cmd_phase(w, sf, cmd, p, lane);
                    cta_bar();
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            sf.rec[mb].t_start_ns = t_start; sf.rec[mb].t_end_ns = (unsigned)t;
#ifdef HLB_PROFILE_STEPS
            sf.rec[mb].mad = (int)w.prof_run_cycles; sf.rec[mb].i16_dc_level[0] = (int16_t)w.prof_runs; sf.rec[mb].me_interp_ops = w.prof_me_cycles;
            for (int i = 0; i < 16; ++i) { ((unsigned*)&sf.rec[mb].i16_ac_level[0][0])[i] = w.prof_lap[i]; ((unsigned*)&sf.rec[mb].i16_ac_level[0][0])[16 + i] = w.prof_cnt[i]; }
#endif
            sched_finish(s, queue, deps, done, sf, job->nmb, item, mb);
        }
        __syncthreads();
    }
}

// ---- throughput variant: ONE WARP per macroblock ----
// The same warp runs the serial control flow (all lanes redundantly) and, inside run(), the command phases with the command's logical lanes
// strided over its 32 lanes.  No CTA barrier exists, so no warp ever waits for another: with enough independent streams in the batch every
// resident warp is runnable all the time (the CTA variant above keeps two thirds of its warps parked at a barrier, r01c profile).
// One warp = one CTA, so the per-macroblock scratch is the CTA's static shared memory.
#ifndef HLB_WARP_MIN_CTAS
#define HLB_WARP_MIN_CTAS 16
#endif
struct WarpExec {
    MbWork* w;
    const FrameCtx* f;
    const SliceJob* job;
    const int* done;
    Sched* sched;
    __device__ __noinline__ void run(int cmd, int nlanes)
    {
        w->arg0_lanes = nlanes;
        __syncwarp();
        const int np = cmd_phases(cmd), l = (int)(threadIdx.x & 31);
#pragma unroll 1
        for (int p = 0; p < np; ++p) {
#pragma unroll 1
            for (int lane = l; lane < nlanes; lane += 32) cmd_phase(*w, *f, cmd, p, lane);
            __syncwarp();
        }
    }
    __device__ __forceinline__ void trials(int nlanes)   // the trial encodes of a search step, straight on this warp's lanes (no command dispatch)
    {
        __syncwarp();
#pragma unroll 1
        for (int lane = (int)(threadIdx.x & 31); lane < nlanes; lane += 32) me_phase_trial(*w, *f, lane);
        __syncwarp();
    }
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int nlanes() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // cross-lane reductions of the (master) warp; every lane receives the result
    __device__ __forceinline__ int reduce_min(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_add(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    // best-MV argmin: the smallest cost wins, the smaller candidate index among equal costs (warp-shuffle butterfly)
    __device__ __forceinline__ void reduce_argmin(double& c, int& i) const
    {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, c, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (oc < c || (oc == c && oi < i)) { c = oc; i = oi; }
        }
    }
    __device__ __noinline__ int prev_sctr(int mb)
    {
        for (int a = mb - 1; a >= 0; --a) {
            int spins = 0;
            while (ld_volatile(done + job->base + a) == 0) {
                if (ld_volatile(&sched->abort)) return 0;
                if (++spins > HLB_SPIN_LIMIT) { watchdog_fire(sched, WD_PREV, job->base + mb, a); return 0; }
                __nanosleep(200);
            }
            __threadfence();
            const int v = *(volatile const uint8_t*)&f->st[a].last_sctr;
            if (v != 255) return v;
        }
        return job->prev_frame_sctr;
    }
};

__global__ void __launch_bounds__(32, HLB_WARP_MIN_CTAS) k_slice_encode_warp(const SliceJob* __restrict__ jobs, int njobs, int* sched_buf)
{
    __shared__ MbWork w;
    __shared__ FrameCtx sf;
    Sched* s = (Sched*)sched_buf;
    const int total = s->total;
    int* queue = sched_buf + HLB_SCHED_WORDS;
    int* deps = queue + total;
    int* done = deps + total;
    const int tid = threadIdx.x;
    if (tid == 0) tile_barrier_init(w);
    __syncwarp();
    for (;;) {
        int item = -1, jb = 0;
        if (tid == 0) item = sched_pop(s, queue, total, jobs, njobs, &jb, HLB_SPIN_MAX_NS_WARP);
        item = __shfl_sync(0xffffffffu, item, 0); jb = __shfl_sync(0xffffffffu, jb, 0);
        if (item < 0) break;
        const SliceJob* job = jobs + jb;
        {
            const int* src = (const int*)&job->f;
            int* dst = (int*)&sf;
            for (int i = tid; i < (int)(sizeof(FrameCtx) / sizeof(int)); i += 32) dst[i] = src[i];
        }
        __syncwarp();
        const int mb = item - job->base;
        unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        WarpExec x;
        x.w = &w; x.f = &sf; x.job = job; x.done = done; x.sched = s;
        mb_encode(x, w, sf, mb);
        __syncwarp();
        if (tid == 0) {
            if (w.stuck) watchdog_fire(s, WD_SEARCH, item, w.stuck);
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            sf.rec[mb].t_start_ns = (unsigned)t0; sf.rec[mb].t_end_ns = (unsigned)t;
#ifdef HLB_PROFILE_STEPS
            for (int i = 0; i < 16; ++i) { ((unsigned*)&sf.rec[mb].i16_ac_level[0][0])[i] = w.prof_lap[i]; ((unsigned*)&sf.rec[mb].i16_ac_level[0][0])[16 + i] = w.prof_cnt[i]; }
#endif
            sched_finish(s, queue, deps, done, sf, job->nmb, item, mb);
        }
        __syncwarp();
    }
}

static size_t state_array_bytes(int nmb) { return ((size_t)nmb * sizeof(MbState) + 255) & ~(size_t)255; }
size_t mbstate_bytes(int nmb) { return state_array_bytes(nmb) + 256; }   // + the Single_ctr chain word
int slice_reset_state(hlb200_ctx* c)
{
    HLB_CUDA(cudaMemsetAsync(c->d_mbstate, 0, mbstate_bytes(c->nmb), c->stream));
    c->frame_count = 0;
    return HLB200_OK;
}

// HLB200_NO_TMA=1: the reference tiles are filled by plain clamped loads instead of TMA + border fix-up (same results; A/B and debugging aid)
static bool slice_no_tma()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("HLB200_NO_TMA"); v = (e && *e && *e != '0') ? 1 : 0; }
    return v != 0;
}
static int build_job(hlb200_ctx* c, const hlb200_slice_params_t* p, SliceJob* j, int base)
{
    if (!c || !p || p->qp < 12 || p->qp > 51 || p->cur_slot < 0 || p->cur_slot >= c->nslots || (p->slice_type != 0 && p->slice_type != 1)) return HLB200_ERR_INVALID_PARAMETER;
    memset(j, 0, sizeof(*j));
    FrameCtx& f = j->f;
    f.W = c->width; f.H = c->height; f.mbw = c->mbw; f.mbh = c->mbh;
    f.qp = p->qp; f.qpc = host_chroma_qp(p->qp, p->chroma_qp_index_offset);
    f.is_p = p->slice_type == 1;
    f.me_range = p->me_range < 1 ? 1 : (p->me_range > 64 ? 64 : p->me_range);   // rdo.c:847
    f.num_refs = f.is_p ? p->num_refs : 0;
    f.early_term = p->me_early_term_flag != 0;
    // the reference's edge map reads one sample around a 16x16 area shifted INTO the picture (rdo.c:894-895): with a 16-sample dimension it leaves the plane
    if (f.early_term && f.is_p && (c->width < 32 || c->height < 32)) { snprintf(g_err, sizeof(g_err), "me_early_term_flag needs a picture of at least 32x32 (rdo.c:894)"); return HLB200_ERR_NOT_IMPLEMENTED; }
    if (f.is_p && (f.num_refs < 1 || f.num_refs > c->max_refs || f.num_refs > HLB_ACTIVE_REFS)) return HLB200_ERR_INVALID_PARAMETER;
    f.lambda = 0.852 * (double)(1 << ((p->qp - 12) / 3));                       // slice.c:1766 (integer division in the exponent)
    frame_ctx_derive(f);
    for (int k = 0; k < 3; ++k) { f.src[k] = c->d_src_cur[k]; f.cur[k] = c->d_slot[p->cur_slot][k]; }
    for (int u = 0; u < f.num_refs; ++u) {
        const int s = p->ref_slot[u];
        if (s < 0 || s >= c->nslots || s == p->cur_slot) return HLB200_ERR_INVALID_PARAMETER;
        for (int k = 0; k < 3; ++k) f.ref[u][k] = c->d_slot[s][k];
        f.ref_tmap[u] = (c->d_tmaps && !slice_no_tma()) ? (const char*)c->d_tmaps + 128 * (size_t)s : nullptr;   // TMA descriptor of the slot's luma plane (hlb_api.cu)
    }
    f.st = (MbState*)c->d_mbstate;
    f.rec = c->d_records;
    j->nmb = c->nmb; j->base = base; j->prev_frame_sctr = 0;
    j->chain = (int*)((char*)c->d_mbstate + state_array_bytes(c->nmb));
    if (p->deblock_flag) {
        if (!c->d_dbk_bs) HLB_CUDA(cudaMalloc(&c->d_dbk_bs, (size_t)c->nmb * 32));
        DbkJob& d = j->d;
        for (int k = 0; k < 3; ++k) d.plane[k] = f.cur[k];
        d.rec = c->d_records; d.bs = (uint8_t*)c->d_dbk_bs;
        d.W = f.W; d.H = f.H; d.mbw = f.mbw; d.mbh = f.mbh; d.enabled = 1;
        dbk_job_thresholds(d, f.qp, f.qpc);
    }
    return HLB200_OK;
}

// Process-wide read-mostly caches (the only globals of the library besides the kernel-variant override below): resident CTAs per device and variant.
#define HLB_MAX_DEVICES 64
static int g_slice_grid[HLB_MAX_DEVICES][2];
static cudaEvent_t g_batch_done[HLB_MAX_DEVICES];   // completion of the last batch launch of each device (launch ordering only)
// resident CTAs of the whole CURRENT device for variant v (0 = CTA per macroblock, 1 = warp per macroblock)
static int slice_grid(int v)
{
    int dev = 0, sms = 0, per = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev >= 0 && dev < HLB_MAX_DEVICES && g_slice_grid[dev][v]) return g_slice_grid[dev][v];
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_slice_encode_warp, 32, 0) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_slice_encode, HLB_CTA_THREADS, 0);
    if (e != cudaSuccess || per < 1) {
        cudaGetLastError();
        return 148;
    }
    if (dev >= 0 && dev < HLB_MAX_DEVICES) g_slice_grid[dev][v] = sms * per;
    return sms * per;
}
// Which variant serves a batch.  The warp variant wins on throughput when the batch offers more ready macroblocks than the CTA variant has
// CTAs (its macroblocks take longer individually); the CTA variant wins on latency for small batches.  HLB200_SLICE_KERNEL=cta|warp overrides.
static int g_slice_force = -2;   // -2 = not initialised, -1 = automatic, 0 / 1 = forced
static int g_slice_last = 0;
static int slice_force()
{
    if (g_slice_force == -2) {
        const char* e = getenv("HLB200_SLICE_KERNEL");
        g_slice_force = !e ? -1 : (!strcmp(e, "warp") ? 1 : (!strcmp(e, "cta") ? 0 : -1));
    }
    return g_slice_force;
}
static int slice_variant(int n_pictures, int mean_wavefront)
{
    const int f = slice_force();
    if (f >= 0) return f;
    return (long long)n_pictures * mean_wavefront >= 4LL * slice_grid(0) ? 1 : 0;   // measured crossover at 1080p: ~100 pictures (r01d A/B: 64 -> CTA 1.41 M vs 1.21 M, 128 -> warp 2.13 M vs 2.05 M)
}


// ------------------------------------------------------------------------------------------------------------------
// Device-side CAVLC serialisation of slice_data(): the macroblock loop of hl_codec_264_nal_slice_data_encode as far as it WRITES
// (slice.c:1786-1894 -> _hl_codec_264_mb_write_no_pcm mb.c:543 incl. the mb_skip_run bookkeeping :588-618 -> hl_codec_264_residual_write residual.c:903),
// from the decision records and the per-macroblock state the slice kernel left.  Three kernels per batch of pictures:
//   k_bits_len    one thread per macroblock: bit length of its mb_skip_run + macroblock_layer()            (same code as the writer, counting sink)
//   k_bits_scan   one CTA per picture: exclusive prefix sum of the lengths, total, zeroing of the used part of the output
//   k_bits_write  one thread per macroblock: the bits, OR-ed into the picture's word buffer at the macroblock's offset
// Output = slice_data() WITHOUT rbsp_trailing_bits, MSB first in big-endian 32-bit words (the host appends them to the slice header with its own bit writer).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bits_len(const BitsJob* __restrict__ jobs)
{
    const BitsJob& j = jobs[blockIdx.y];
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= j.nmb) return;
    BitCount c;
    c.n = 0;
    bits_put_mb(c, j, mb);
    j.len[mb] = c.n;
}
__global__ void __launch_bounds__(256) k_bits_scan(const BitsJob* __restrict__ jobs)
{
    const BitsJob& j = jobs[blockIdx.x];
    __shared__ uint32_t part[256];
    const int t = threadIdx.x, per = (j.nmb + 255) / 256, a = t * per, b = min(j.nmb, a + per);
    uint32_t sum = 0;
    for (int i = a; i < b; ++i) sum += j.len[i];
    part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {   // inclusive scan of the 256 partial sums
        const uint32_t v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t off = t ? part[t - 1] : 0;
    for (int i = a; i < b; ++i) { const uint32_t l = j.len[i]; j.len[i] = off; off += l; }
    const uint32_t total = part[255];
    const uint32_t words = (total + 31) / 32 + 2;
    if (t == 0) { j.hdr[0] = total; j.hdr[1] = words > (uint32_t)j.cap_words ? 1u : 0u; }
    const uint32_t nz = words > (uint32_t)j.cap_words ? (uint32_t)j.cap_words : words;
    for (uint32_t i = t; i < nz; i += 256) j.out[i] = 0;
}
__global__ void __launch_bounds__(128) k_bits_write(const BitsJob* __restrict__ jobs)
{
    const BitsJob& j = jobs[blockIdx.y];
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= j.nmb || j.hdr[1]) return;
    BitWriter w;
    w.buf = j.out; w.pos = j.len[mb]; w.acc = 0; w.nacc = 0;
    bits_put_mb(w, j, mb);
    w.finish();
}

}  // namespace hlb
using namespace hlb;

extern "C" {

// Watchdog state of the launch that last covered `c` (its scheduler words live in the batch owner): synchronises the owner's stream once per launch.
static int slice_check_abort(hlb200_ctx* c)
{
    hlb200_ctx* o = c->batch_owner ? c->batch_owner : c;
    if (!o->last_sched) return HLB200_OK;
    if (o->abort_state < 0) {
        int words[4] = {0, 0, 0, 0};
        HLB_CUDA(cudaStreamSynchronize(o->stream));
        HLB_CUDA(cudaMemcpy(words, o->last_sched, sizeof(words), cudaMemcpyDeviceToHost));
        o->abort_state = words[3];
    }
    if (o->abort_state) {
        snprintf(g_err, sizeof(g_err), "slice kernel watchdog fired (code %d): the launch was drained early, its records are incomplete (hlb200_slice_status has the details)", o->abort_state);
        return HLB200_ERR_INVALID_STATE;
    }
    return HLB200_OK;
}

int hlb200_slice_encode_batch_async(hlb200_ctx_t** ctxs, const hlb200_slice_params_t* params, int n)
{
    if (!ctxs || !params || n < 1 || n > 4096) return HLB200_ERR_INVALID_PARAMETER;
    hlb200_ctx* c0 = ctxs[0];
    if (!c0) return HLB200_ERR_INVALID_PARAMETER;
    int total = 0;
    for (int i = 0; i < n; ++i) { if (!ctxs[i]) return HLB200_ERR_INVALID_PARAMETER; total += ctxs[i]->nmb; }
    // mean width of the 2:1 wavefront of a picture = macroblocks / (mbw + 2 (mbh - 1)) dependency steps
    const int variant = g_slice_last = slice_variant(n, (c0->nmb + c0->mbw + 2 * c0->mbh - 3) / (c0->mbw + 2 * c0->mbh - 2));
    int grid = slice_grid(variant);
    // A macroblock of a P picture that needs the Single_ctr chain value of its raster predecessor waits for it while holding its CTA (prev_sctr): at most one CTA
    // per picture can be parked that way (the parked macroblock blocks the first column below it), and the macroblock it waits for is always claimable by a
    // free CTA.  With fewer pictures than resident CTAs a free CTA always exists, so the launch cannot deadlock; larger batches are split.
    const int max_pictures = grid > 2 ? grid / 2 : 1;
    if (n > max_pictures) {
        for (int i = 0; i < n; i += max_pictures) {
            const int rc = hlb200_slice_encode_batch_async(ctxs + i, params + i, n - i < max_pictures ? n - i : max_pictures);
            if (rc) return rc;
        }
        return HLB200_OK;
    }
    // scheduler + job storage lives in the first context of the batch
    const size_t need = sizeof(SliceJob) * (size_t)n + sizeof(int) * (HLB_SCHED_WORDS + 3 * (size_t)total) + 512;
    if (c0->sched_bytes < need) {
        if (c0->d_sched) HLB_CUDA(cudaFree(c0->d_sched));
        c0->d_sched = nullptr; c0->sched_bytes = 0;
        HLB_CUDA(cudaMalloc((void**)&c0->d_sched, need));
        c0->sched_bytes = need;
    }
    if (!c0->h_jobs || c0->h_jobs_cap < n) {
        if (c0->h_jobs) { HLB_CUDA(cudaStreamSynchronize(c0->stream)); HLB_CUDA(cudaFreeHost(c0->h_jobs)); c0->h_jobs = nullptr; }
        HLB_CUDA(cudaMallocHost(&c0->h_jobs, sizeof(SliceJob) * (size_t)n));
        c0->h_jobs_cap = n;
    }
    if (!c0->ev_jobs) { HLB_CUDA(cudaEventCreateWithFlags(&c0->ev_jobs, cudaEventDisableTiming)); HLB_CUDA(cudaEventCreateWithFlags(&c0->ev_done, cudaEventDisableTiming)); }
    else HLB_CUDA(cudaEventSynchronize(c0->ev_jobs));   // the pinned job array must not be rewritten while the previous launch's copy of it is still in flight
    SliceJob* hj = (SliceJob*)c0->h_jobs;
    int base = 0;
    for (int i = 0; i < n; ++i) {
        const int rc = build_job(ctxs[i], params + i, hj + i, base);
        if (rc) return rc;
        base += ctxs[i]->nmb;
    }
    SliceJob* dj = (SliceJob*)c0->d_sched;
    int* sched = (int*)((char*)c0->d_sched + ((sizeof(SliceJob) * (size_t)n + 255) & ~(size_t)255));
    cudaStream_t st = c0->stream;
    // ORDERING RULE (include/hlb200.h): the batch kernel runs on the stream of ctxs[0].  Work queued earlier on the other contexts' streams (source uploads) is made
    // a dependency of the launch, and everything queued later on them (record / slot downloads, the next upload) waits for the launch: events, no host sync.
    for (int i = 1; i < n; ++i)
        if (ctxs[i]->stream != st) {
            if (!ctxs[i]->ev_done) HLB_CUDA(cudaEventCreateWithFlags(&ctxs[i]->ev_done, cudaEventDisableTiming));
            HLB_CUDA(cudaEventRecord(ctxs[i]->ev_done, ctxs[i]->stream));
            HLB_CUDA(cudaStreamWaitEvent(st, ctxs[i]->ev_done, 0));
        }
    // Batch launches of a device run back to back, whatever streams their contexts own: the slice kernel is persistent and sized to the whole device, two of them
    // side by side only take each other's SMs (measured: 256 streams as 2 concurrent batches 3.3 M MB/s, as one 4.0 M), while queued behind each other they let the
    // host work of one batch overlap the kernel of the next.
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < HLB_MAX_DEVICES) {
            if (!g_batch_done[dev]) HLB_CUDA(cudaEventCreateWithFlags(&g_batch_done[dev], cudaEventDisableTiming));
            else HLB_CUDA(cudaStreamWaitEvent(st, g_batch_done[dev], 0));
        }
    }
    HLB_CUDA(cudaMemcpyAsync(dj, hj, sizeof(SliceJob) * (size_t)n, cudaMemcpyHostToDevice, st));
    HLB_CUDA(cudaEventRecord(c0->ev_jobs, st));
    c0->last_sched = sched;
    k_slice_init<<<(total + 255) / 256, 256, 0, st>>>(dj, n, sched, total);
    HLB_CUDA(cudaGetLastError());
    if (grid > total) grid = total;
    if (variant) k_slice_encode_warp<<<grid, 32, 0, st>>>(dj, n, sched);
    else k_slice_encode<<<grid, HLB_CTA_THREADS, 0, st>>>(dj, n, sched);
    HLB_CUDA(cudaGetLastError());
    {   // loop filter of the pictures that ask for it, before anything can read them as references
        int any = 0, maxnmb = 0, maxmbh = 0;
        for (int i = 0; i < n; ++i)
            if (hj[i].d.enabled) { any = 1; maxnmb = hj[i].nmb > maxnmb ? hj[i].nmb : maxnmb; maxmbh = hj[i].d.mbh > maxmbh ? hj[i].d.mbh : maxmbh; }
        if (any) {
            k_dbk_bs<<<dim3((maxnmb * 32 + 255) / 256, n), 256, 0, st>>>(dj);
            k_dbk<<<n, HLB_DBK_WARPS * 32, sizeof(int) * (size_t)maxmbh, st>>>(dj);
            HLB_CUDA(cudaGetLastError());
        }
    }
    HLB_CUDA(cudaEventRecord(c0->ev_done, st));
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < HLB_MAX_DEVICES && g_batch_done[dev]) HLB_CUDA(cudaEventRecord(g_batch_done[dev], st)); }
    c0->abort_state = -1;
    for (int i = 0; i < n; ++i) {
        ctxs[i]->frame_count++;
        ctxs[i]->batch_owner = c0;
        if (ctxs[i]->stream != st) HLB_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, c0->ev_done, 0));
    }
    return HLB200_OK;
}


// ---- slice_data() bits of the pictures the contexts encoded last (any mix of sizes); the launch goes to the stream of ctxs[0], ordered like the slice launch ----
int hlb200_slice_bits_batch_async(hlb200_ctx_t** ctxs, const int32_t* slice_types, int n)
{
    if (!ctxs || !slice_types || n < 1 || n > 4096) return HLB200_ERR_INVALID_PARAMETER;
    hlb200_ctx* c0 = ctxs[0];
    if (!c0) return HLB200_ERR_INVALID_PARAMETER;
    int maxnmb = 0;
    for (int i = 0; i < n; ++i) {
        hlb200_ctx* c = ctxs[i];
        if (!c || (slice_types[i] != 0 && slice_types[i] != 1)) return HLB200_ERR_INVALID_PARAMETER;
        if (!c->d_bits) {
            c->bits_cap_words = c->nmb * HLB200_BITS_WORDS_PER_MB + 64;
            HLB_CUDA(cudaMalloc(&c->d_bits, sizeof(uint32_t) * ((size_t)c->bits_cap_words + (size_t)c->nmb + 4)));
        }
        maxnmb = c->nmb > maxnmb ? c->nmb : maxnmb;
    }
    if (c0->bits_jobs_cap < n) {
        if (c0->d_bits_jobs) { HLB_CUDA(cudaStreamSynchronize(c0->stream)); HLB_CUDA(cudaFree(c0->d_bits_jobs)); HLB_CUDA(cudaFreeHost(c0->h_bits_jobs)); c0->d_bits_jobs = nullptr; c0->h_bits_jobs = nullptr; }
        HLB_CUDA(cudaMalloc(&c0->d_bits_jobs, sizeof(BitsJob) * (size_t)n));
        HLB_CUDA(cudaMallocHost(&c0->h_bits_jobs, sizeof(BitsJob) * (size_t)n));
        c0->bits_jobs_cap = n;
    }
    if (!c0->ev_bits) HLB_CUDA(cudaEventCreateWithFlags(&c0->ev_bits, cudaEventDisableTiming));
    else HLB_CUDA(cudaEventSynchronize(c0->ev_bits));   // the pinned job array of the previous call has been consumed
    BitsJob* hj = (BitsJob*)c0->h_bits_jobs;
    for (int i = 0; i < n; ++i) {
        hlb200_ctx* c = ctxs[i];
        BitsJob& j = hj[i];
        j.rec = c->d_records; j.st = (const MbState*)c->d_mbstate;
        j.out = (uint32_t*)c->d_bits; j.len = j.out + c->bits_cap_words; j.hdr = j.len + c->nmb;
        j.nmb = c->nmb; j.mbw = c->mbw; j.is_p = slice_types[i]; j.cap_words = c->bits_cap_words;
    }
    cudaStream_t st = c0->stream;
    for (int i = 1; i < n; ++i)
        if (ctxs[i]->stream != st) {
            if (!ctxs[i]->ev_done) HLB_CUDA(cudaEventCreateWithFlags(&ctxs[i]->ev_done, cudaEventDisableTiming));
            HLB_CUDA(cudaEventRecord(ctxs[i]->ev_done, ctxs[i]->stream));
            HLB_CUDA(cudaStreamWaitEvent(st, ctxs[i]->ev_done, 0));
        }
    HLB_CUDA(cudaMemcpyAsync(c0->d_bits_jobs, hj, sizeof(BitsJob) * (size_t)n, cudaMemcpyHostToDevice, st));
    HLB_CUDA(cudaEventRecord(c0->ev_bits, st));
    const BitsJob* dj = (const BitsJob*)c0->d_bits_jobs;
    k_bits_len<<<dim3((maxnmb + 127) / 128, n), 128, 0, st>>>(dj);
    k_bits_scan<<<n, 256, 0, st>>>(dj);
    k_bits_write<<<dim3((maxnmb + 127) / 128, n), 128, 0, st>>>(dj);
    HLB_CUDA(cudaGetLastError());
    if (!c0->ev_done) HLB_CUDA(cudaEventCreateWithFlags(&c0->ev_done, cudaEventDisableTiming));
    HLB_CUDA(cudaEventRecord(c0->ev_done, st));
    // The next batch launch of the device waits for these three short kernels too: the persistent slice kernel fills every SM (registers and shared memory at
    // their limits), so serialisation kernels that become runnable together with it would only get their CTAs placed when it drains -- and the downloads
    // behind them would wait a whole launch (measured through hl_codec_encode with two groups of 256 streams in flight).
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < HLB_MAX_DEVICES && g_batch_done[dev]) HLB_CUDA(cudaEventRecord(g_batch_done[dev], st)); }
    for (int i = 1; i < n; ++i)
        if (ctxs[i]->stream != st) HLB_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, c0->ev_done, 0));
    return HLB200_OK;
}

// Downloads the bits of the context's last picture: *nbits_out = length of slice_data() in bits, out_words[0 .. (nbits + 31) / 32) = the bits, MSB first, each
// word in HOST byte order (word k holds bits 32k .. 32k+31 with bit 32k in its most significant position).
int hlb200_slice_bits_download(hlb200_ctx_t* c, uint32_t* out_words, size_t cap_words, uint32_t* nbits_out)
{
    if (!c || !out_words || !nbits_out || !c->d_bits) return HLB200_ERR_INVALID_PARAMETER;
    int rc = slice_check_abort(c);
    if (rc) return rc;
    uint32_t hdr[2] = {0, 0};
    const uint32_t* d_hdr = (const uint32_t*)c->d_bits + c->bits_cap_words + c->nmb;
    HLB_CUDA(cudaMemcpyAsync(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost, c->stream));
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    if (hdr[1]) { snprintf(g_err, sizeof(g_err), "slice data of %u bits does not fit the device bit buffer (%d words)", hdr[0], c->bits_cap_words); return HLB200_ERR_OUTOFMEMORY; }
    const size_t words = ((size_t)hdr[0] + 31) / 32;
    if (words > cap_words) return HLB200_ERR_OUTOFMEMORY;
    HLB_CUDA(cudaMemcpyAsync(out_words, c->d_bits, sizeof(uint32_t) * words, cudaMemcpyDeviceToHost, c->stream));
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    *nbits_out = hdr[0];
    return HLB200_OK;
}

int hlb200_slice_encode_async(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params) { return hlb200_slice_encode_batch_async(&ctx, params, 1); }

// reads the watchdog words of the last launch whose scheduler lives in `c` (after the stream has drained)
int hlb200_slice_status(hlb200_ctx_t* c, int* out16)
{
    if (!c || !out16) return HLB200_ERR_INVALID_PARAMETER;
    for (int k = 0; k < 16; ++k) out16[k] = 0;
    if (c->batch_owner) c = c->batch_owner;   // the scheduler words live in the first context of the batch
    if (!c->d_sched || !c->last_sched) return HLB200_OK;
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    HLB_CUDA(cudaMemcpy(out16, c->last_sched, sizeof(int) * 16, cudaMemcpyDeviceToHost));
    return out16[3] ? HLB200_ERR_INVALID_STATE : HLB200_OK;
}

int hlb200_records_download(hlb200_ctx_t* c, hlb200_mb_record_t* out_records)
{
    if (!c || !out_records) return HLB200_ERR_INVALID_PARAMETER;
    HLB_CUDA(cudaMemcpyAsync(out_records, c->d_records, sizeof(hlb200_mb_record_t) * (size_t)c->nmb, cudaMemcpyDeviceToHost, c->stream));
    HLB_CUDA(cudaStreamSynchronize(c->stream));
    return slice_check_abort(c);   // an aborted launch leaves macroblocks unencoded: never hand such records to a writer
}

int hlb200_slice_encode(hlb200_ctx_t* ctx, const hlb200_slice_params_t* params, hlb200_mb_record_t* out_records)
{
    int rc = hlb200_slice_encode_async(ctx, params);
    if (rc) return rc;
    return hlb200_records_download(ctx, out_records);
}

int hlb200_slice_grid_size(void) { return slice_grid(slice_force() >= 0 ? slice_force() : g_slice_last); }
int hlb200_slice_last_variant(void) { return g_slice_last; }
int hlb200_slice_set_variant(int variant)
{
    const int prev = slice_force();
    g_slice_force = variant == 0 || variant == 1 ? variant : -1;
    return prev;
}

}  // extern "C"
