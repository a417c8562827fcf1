// emu_main.cpp -- CPU emulation harness of hlb_mbcore.cuh (DEBUGGING AID, never shipped, never a fallback): compiles the same
// per-macroblock control flow as the CUDA kernel as plain C++ (lanes become loops, macroblocks run in raster order) so that the
// decision logic can be diffed against traces of the reference encoder without a GPU.
//   emu --size W H --frames N --qp Q --me-range R --refs K --in frames.yuv --out prefix
// writes prefix.recon (planes per frame), prefix.rec (hlb200_mb_record_t per MB per frame), prefix.st (MbState per MB per frame)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define HLB_EMU_DEBUG 1
int g_emu_dbg = 0;
#include "../../hartallo_b200/csrc/hlb_mbcore.cuh"
#include "../../hartallo_b200/csrc/hlb_deblock.cuh"

using namespace hlb;

struct CpuExec {
    MbWork* w;
    const FrameCtx* f;
    int prev_frame_sctr;
    void run(int cmd, int nlanes)
    {
        w->cmd = cmd; w->arg0_lanes = nlanes;
        const int np = cmd_phases(cmd);
        for (int p = 0; p < np; ++p)
            for (int lane = 0; lane < nlanes; ++lane) cmd_phase(*w, *f, cmd, p, lane);
    }
    void trials(int nlanes) { for (int lane = 0; lane < nlanes; ++lane) me_phase_trial(*w, *f, lane); }
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

int main(int argc, char** argv)
{
    int W = 352, H = 288, frames = 2, qp = 31, me_range = 16, refs = 1, active_refs = 1, memo_stats = 0, early_term = 0, deblock = 0;
    const char *in = nullptr, *out = "emu";
    int dbg_frame = -1, dbg_mb = -1;
    if (getenv("EMU_DEBUG_MB")) sscanf(getenv("EMU_DEBUG_MB"), "%d:%d", &dbg_frame, &dbg_mb);
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--size")) { W = atoi(argv[++i]); H = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--frames")) frames = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--qp")) qp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--me-range")) me_range = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--refs")) refs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--active-refs")) active_refs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--memo-stats")) memo_stats = 1;
        else if (!strcmp(argv[i], "--early-term")) early_term = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--deblock")) deblock = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--in")) in = argv[++i];
        else if (!strcmp(argv[i], "--out")) out = argv[++i];
    }
    if (!in) { fprintf(stderr, "--in required\n"); return 2; }
    const int nmb = (W / 16) * (H / 16);
    const size_t fb = (size_t)W * H * 3 / 2, ys = (size_t)W * H, cs = ys / 4;
    std::vector<uint8_t> src(fb);
    std::vector<std::vector<uint8_t>> slots(refs + 1, std::vector<uint8_t>(fb, 0));
    std::vector<MbState> st(nmb);
    memset(st.data(), 0, sizeof(MbState) * nmb);
    std::vector<hlb200_mb_record_t> rec(nmb);
    FILE* fi = fopen(in, "rb");
    if (!fi) { perror(in); return 2; }
    char path[512];
    snprintf(path, sizeof(path), "%s.recon", out); FILE* fr = fopen(path, "wb");
    snprintf(path, sizeof(path), "%s.rec", out); FILE* fc = fopen(path, "wb");
    snprintf(path, sizeof(path), "%s.st", out); FILE* fs = fopen(path, "wb");
    MbWork* w = new MbWork();
    memset(w, 0, sizeof(*w));
    int chain = 0;
    std::vector<int> order;  // slots holding previous reconstructions, most recent first
    for (int n = 0; n < frames; ++n) {
        if (fread(src.data(), 1, fb, fi) != fb) break;
        FrameCtx f;
        memset(&f, 0, sizeof(f));
        f.W = W; f.H = H; f.mbw = W / 16; f.mbh = H / 16; f.qp = qp;
        { int q = qp < 0 ? 0 : (qp > 51 ? 51 : qp); static const unsigned char t[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39}; f.qpc = q < 30 ? q : t[q - 30]; }
        f.early_term = early_term;
        f.is_p = n > 0; f.me_range = me_range < 1 ? 1 : (me_range > 64 ? 64 : me_range);
        f.lambda = 0.852 * (double)(1 << ((qp - 12) / 3));
        frame_ctx_derive(f);
        int cur = 0;
        for (int s = 0; s <= refs; ++s) { bool used = false; for (int o : order) used |= (o == s); if (!used) { cur = s; break; } }
        // the reference searches num_ref_idx_l0_active_minus1 + 1 = 1 list entry whatever max_ref_frame is (slice.c:289-291, rdo.c:845);
        // --active-refs raises that only to exercise the multi-reference loop
        f.num_refs = (int)order.size() < active_refs ? (int)order.size() : active_refs;
        if (f.num_refs > refs) f.num_refs = refs;
        if (f.num_refs < 1) f.num_refs = 1;
        for (int u = 0; u < f.num_refs && u < (int)order.size(); ++u) { f.ref[u][0] = slots[order[u]].data(); f.ref[u][1] = f.ref[u][0] + ys; f.ref[u][2] = f.ref[u][1] + cs; }
        f.src[0] = src.data(); f.src[1] = f.src[0] + ys; f.src[2] = f.src[1] + cs;
        f.cur[0] = slots[cur].data(); f.cur[1] = f.cur[0] + ys; f.cur[2] = f.cur[1] + cs;
        f.st = st.data(); f.rec = rec.data();
        CpuExec x{w, &f, chain};
        for (int mb = 0; mb < nmb; ++mb) {
            g_emu_dbg = (n == dbg_frame && mb == dbg_mb);
            if (g_emu_dbg) fprintf(stderr, "frame %d mb %d\n", n, mb);
            mb_encode(x, *w, f, mb);
        }
        chain = x.prev_sctr(nmb);
        if (deblock) {   // loop filter over the finished picture (slice.c:1897): what later pictures predict from
            hlb::DbkJob d;
            memset(&d, 0, sizeof(d));
            for (int k = 0; k < 3; ++k) d.plane[k] = f.cur[k];
            d.rec = rec.data(); d.W = W; d.H = H; d.mbw = f.mbw; d.mbh = f.mbh; d.enabled = 1;
            hlb::dbk_job_thresholds(d, f.qp, f.qpc);
            hlb::dbk_picture_serial(d);
        }
        fwrite(slots[cur].data(), 1, fb, fr);
        fwrite(rec.data(), sizeof(hlb200_mb_record_t), nmb, fc);
        fwrite(st.data(), sizeof(MbState), nmb, fs);
        order.insert(order.begin(), cur);
        if ((int)order.size() > refs) order.pop_back();
    }
    fclose(fi); fclose(fr); fclose(fc); fclose(fs);
    if (memo_stats) {
        const hlb::MemoStats& m = hlb::g_memo_stats;
        printf("trial memo on the reference trajectory: %ld trials, %ld repeats (%.1f%%); %ld search steps, %ld of repeats only (%.1f%%)\n", m.trials, m.hits,
               100.0 * m.hits / (m.trials ? m.trials : 1), m.steps, m.full_hit_steps, 100.0 * m.full_hit_steps / (m.steps ? m.steps : 1));
    }
    printf("emu: %d frames, sizeof(MbWork)=%zu sizeof(MbState)=%zu sizeof(rec)=%zu\n", frames, sizeof(MbWork), sizeof(MbState), sizeof(hlb200_mb_record_t));
    return 0;
}
