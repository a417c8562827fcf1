// hlb_mbcore.cuh -- per-macroblock decide + reconstruct of a P (or I) picture: the device-side counterpart of the loop body of
// hl_codec_264_nal_slice_data_encode (source/h264/hl_codec_264_slice.c:1786-1894), i.e.
//   hl_codec_264_rdo_mb_guess_best_inter_pred_avc   source/h264/hl_codec_264_rdo.c:678
//   hl_codec_264_me_ds_mb_find_best_cost            source/h264/hl_codec_264_me_ds.c:104   (+ compute_cost_mode :527)
//   mvp / PSkip mv / neighbour partitions           source/h264/hl_codec_264_utils.c:709-963, source/h264/hl_codec_264_mb.c:426
//   nC of the CAVLC rate term                       source/h264/hl_codec_264_residual.c:624-755
//   reconstruction, chroma, CBP                     rdo.c:2140-2780, source/h264/hl_codec_264_transf.c:161
//   intra decision (I pictures and inside P)        rdo.c:99-300, 1526-2137  (hlb_mbintra.cuh)
//
// Execution model ("master / lanes"): the serial control flow of one macroblock (search trajectory, mode loop, commits) is
// written once as ordinary sequential code; every data-parallel piece (<= 9 candidates x 16 4x4 blocks of trial encodes, the
// 24 blocks of a reconstruction, intra mode x block trials) is a *command* made of phases, each phase a function of (work, lane).
// On the GPU one CTA owns one macroblock: thread 0 runs the control flow, posts a command in shared memory and all threads
// run its phases with a CTA barrier between phases.  The same source compiles as plain C++ (lanes become loops) for the CPU
// emulation harness under tools/emu, which exists to debug the control flow against traces of the reference -- it is never
// part of the shipped library.
#pragma once
#include <float.h>
#include <limits.h>
#include <stddef.h>

#include "../../include/hlb200.h"
#include "hlb_prims.cuh"
#include "hlb_fast.cuh"
#include "hlb_intra.cuh"

#ifdef HLB_EMU_DEBUG
#include <stdio.h>
extern int g_emu_dbg;   // set by the emulation harness for the macroblock under investigation
#define HLB_DBG(...) do { if (g_emu_dbg) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define HLB_DBG(...) do { } while (0)
#endif

namespace hlb {

enum { MBK_PSKIP = HLB200_MB_P_SKIP, MBK_INTER = HLB200_MB_P_INTER, MBK_I16 = HLB200_MB_I16x16, MBK_I4 = HLB200_MB_I4x4 };

// State of one macroblock address that outlives the macroblock (read by later MBs of the picture and, for the fields the
// reference never resets, by the same address in the next picture -- SURVEY Appendix C).
struct MbState {
    uint8_t kind;          // MBK_*
    uint8_t part_mode;     // 0 16x16, 1 16x8, 2 8x16, 3 8x8 (final geometry; 0 for PSkip)
    uint8_t sub_mode[4];   // 0 8x8, 1 8x4, 2 4x8, 3 4x4
    uint8_t cbp_luma;      // CodedBlockPatternLuma (8x8 bits)
    uint8_t cbp_chroma;    // CodedBlockPatternChroma
    uint8_t tc_luma[16];   // TotalCoeffsLuma[] exactly as the reference leaves it (trial leftovers included)
    uint8_t tc_cac[2][4];  // TotalCoeffsChromaACCbCr
    int8_t ref_idx[4];     // RefIdxL0
    uint8_t i4_mode[16];   // Intra4x4PredMode
    uint8_t last_sctr;     // pc_esd->rdo.Single_ctr after this macroblock (a chain through raster order, residual.c:882)
    uint8_t pad[3];
    alignas(4) int16_t mv[4][4][2];       // MvL0
    alignas(4) int16_t chroma_ac[2][4][16];  // ChromaACLevel (persistent: only rewritten for blocks with a non-zero residual, rdo.c:2577)
    int16_t chroma_dc[2][4];      // ChromaDCLevel
};

#define HLB_NB_WORDS 30   /* MbState up to and including mv[][][] */
static_assert(offsetof(MbState, chroma_ac) == 4 * HLB_NB_WORDS, "MbState head layout");

#define HLB_ACTIVE_REFS 4   /* list entries one launch can search (the picture context travels to shared memory: kept small) */
struct FrameCtx {
    int W, H, mbw, mbh;
    int qp, qpc;
    int is_p, me_range, num_refs;
    int early_term;                      // me_early_term_flag: homogeneous-block detection narrows the partition modes searched (rdo.c:889-935)
    double lambda;                       // lambda_mode = 0.852 * (1 << ((QP-12)/3)), slice.c:1766
    QuantK qk;                           // quantiser constants of the luma trial encodes (hlb_fast.cuh), derived from qp by frame_ctx_derive()
    QuantK qkc;                          // chroma AC: QPC with the intra rounding offset whatever the macroblock is (rdo.c:2588)
    const uint8_t* src[3];
    uint8_t* cur[3];                     // reconstruction of the current picture (frame-store planes, pitch = W / W/2)
    const uint8_t* ref[HLB_ACTIVE_REFS][3];  // active list entries: the reference's slice headers always carry num_ref_idx_l0_active_minus1 = 0 (slice.c:289-291)
    const void* ref_tmap[HLB_ACTIVE_REFS];   // device: CUtensorMap (in global memory) of each reference luma plane for the TMA tile loads; null = plain loads
    MbState* st;
    hlb200_mb_record_t* rec;
};

// fields derived from the others (host side, once per picture)
inline void frame_ctx_derive(FrameCtx& f) { quantk_make(f.qk, f.qp); quantk_make(f.qkc, f.qpc, true); }

// ---- partition geometry of the 7 search modes (rdo.c:711-809): 0 16x16, 1 16x8, 2 8x16, 3 8x8, 4 8x4, 5 4x8, 6 4x4 ----
HLB_HD int mode_nparts(int m) { return m == 0 ? 1 : (m < 3 ? 2 : 4); }
HLB_HD int mode_nsub(int m) { return m <= 3 ? 1 : (m == 6 ? 4 : 2); }
HLB_HD int mode_part_w(int m) { return (m == 0 || m == 1) ? 16 : 8; }
HLB_HD int mode_part_h(int m) { return (m == 0 || m == 2) ? 16 : 8; }
HLB_HD int mode_sub_w(int m) { return m <= 3 ? mode_part_w(m) : ((m == 4) ? 8 : 4); }
HLB_HD int mode_sub_h(int m) { return m <= 3 ? mode_part_h(m) : ((m == 5) ? 8 : 4); }
HLB_HD void mode_rect(int m, int part, int sub, int& ox, int& oy, int& w, int& h)
{
    const int pw = mode_part_w(m), ph = mode_part_h(m);
    w = mode_sub_w(m); h = mode_sub_h(m);
    if (pw == 16) { ox = 0; oy = part * ph; }
    else { ox = (part & 1) * 8; oy = ph == 16 ? 0 : (part >> 1) * 8; }
    if (m > 3) {
        if (w == 8) oy += sub * h;
        else { ox += (sub & 1) * 4; oy += h == 8 ? 0 : (sub >> 1) * 4; }
    }
}
// (mbPartIdx, subMbPartIdx) covering luma position (x,y) for a geometry given as part_mode + sub_mode[] (6.4.12.4, mb.h:313)
HLB_HD void part_at(int part_mode, const uint8_t* sub_mode, int x, int y, int& part, int& sub)
{
    sub = 0;
    switch (part_mode) {
    case 0: part = 0; break;
    case 1: part = y >> 3; break;
    case 2: part = x >> 3; break;
    default: {
        part = ((y >> 3) << 1) | (x >> 3);
        const int lx = x & 7, ly = y & 7;
        switch (sub_mode[part]) {
        case 0: sub = 0; break;
        case 1: sub = ly >> 2; break;
        case 2: sub = lx >> 2; break;
        default: sub = ((ly >> 2) << 1) | (lx >> 2); break;
        }
    }
    }
}

#define HLB_MAXC 9
#define HLB_MB_LANES 160   /* lanes a command needs at most = worker threads of the GPU CTA */
#define HLB_MEMO_SLOTS 32   /* per 4x4 block; a macroblock meets ~25 distinct vectors per block (G2 CIF), probing never evicts */
/* Shared-memory reference tile.  A TMA tile load wants its innermost start coordinate on a 16-byte boundary (measured: any other x faults, profiles/r02b), so
 * the tile is 64 samples wide, starts at a multiple of 16 and guarantees a window of HLB_TILE_SPAN_W = 48 anywhere inside it; 40 rows = partition (16) + 6-tap
 * halo (5) + +-9 rows of search freedom (a search step spans at most +-2 integer samples; wider steps are evaluated candidate by candidate). */
#define HLB_TILE_W 64
#define HLB_TILE_H 40
#define HLB_TILE_SPAN_W 48
#define HLB_TILE_BYTES (HLB_TILE_W * HLB_TILE_H)
enum { CMD_NONE = 0, CMD_EXIT, CMD_LOAD, CMD_TILE, CMD_TILE_TMA, CMD_TILE_FIX, CMD_ME_EVAL, CMD_PRED_INTER, CMD_RECON_LUMA, CMD_CHROMA, CMD_STORE, CMD_I16_EVAL, CMD_I16_RATE, CMD_I16_RECON, CMD_I4_EVAL, CMD_I4_COMMIT, CMD_PRED_CHROMA_INTRA };

// Lap timer of the profiling build: attributes the cycles since the previous HLB_LAP of this macroblock to section `slot`.
// Sections: 0 begin/load, 1 mvp+pattern (search control), 2 me_eval prelude, 3 tile, 4 trial run, 5 scan+token, 6 cost, 7 compare,
//           8 mode bookkeeping, 9 pskip chroma check, 10 intra, 11 final recon, 12 commit
#if defined(HLB_PROFILE_STEPS) && defined(__CUDA_ARCH__)
#define HLB_LAP(w, slot) do { const long long _t = clock64(); (w).prof_lap[slot] += (unsigned)(_t - (w).prof_last); (w).prof_cnt[slot]++; (w).prof_last = _t; } while (0)
#else
#define HLB_LAP(w, slot) do { } while (0)
#endif

// All lanes of the (master) warp call this converged; every loop is strided over them (lane, nl) -- on the CPU harness lane = 0, nl = 1.
#if defined(__CUDA_ARCH__)
#define HLB_LANE_SYNC() __syncwarp()
#else
#define HLB_LANE_SYNC() do { } while (0)
#endif

// Scratch of the macroblock being encoded (shared memory on the GPU)
struct MbWork {
    // reference tile: tile[j * HLB_TILE_W + i] = ref_y[clampY(tile_y0 + j)][clampX(tile_x0 + i)]; TMA destination (128-byte aligned, first member: no padding);
    // the tail pads the word loads of the last row
    alignas(128) uint8_t tile[HLB_TILE_BYTES + 16];
    alignas(8) unsigned long long tile_mbar;          // mbarrier the TMA tile load completes on (device)
    // command mailbox
    int cmd, arg0, arg1, arg0_lanes;
    // identity / neighbourhood
    int mb, mbx, mby;
    int availA, availB, availC, availD;
    // source samples
    alignas(4) uint8_t src_y[256];   // moved with 32-bit accesses
    alignas(4) uint8_t src_c[2][64];
    // evolving state of the current macroblock
    uint8_t tc[16];        // TotalCoeffsLuma
    uint8_t tc_cac[2][4];
    uint8_t cbp_gate;      // CodedBlockPatternLuma as left by the previous picture (gate of in-MB neighbours, utils.h:10-20)
    int8_t extA[16], extB[16];   // nA / nB contributed by the neighbouring macroblocks (-1 = not available), per luma4x4BlkIdx on the MB edge
    // leading HLB_NB_WORDS words of the persistent state of [0] this address (as the previous picture left it), [1] A, [2] B, [3] C, [4] D:
    // everything the motion / nC derivations read, fetched once with one coalesced copy instead of dependent L2 round trips per derivation
    uint32_t nbw[5][HLB_NB_WORDS];
    unsigned prof_run_cycles, prof_runs, prof_me_cycles;   // HLB_PROFILE_STEPS builds only
#ifdef HLB_PROFILE_STEPS
    unsigned prof_lap[16], prof_cnt[16];                    // lap timer: cycles / visits per section (HLB_LAP)
    long long prof_last;
#endif
    unsigned stat_trials, stat_interp, stat_cands, stat_intra;   // work counters of the trajectory (roofline accounting)
    int stuck;             // set when a search loop exceeded its iteration cap (cannot happen for a finite window; watchdog aid)
    int last_sctr;         // rdo.Single_ctr chain; -1 = not yet written by this macroblock
    int need_prev_sctr;    // set when the chain value of the raster predecessor was consumed
    alignas(4) int16_t chroma_ac[2][4][16];   // copied with 32-bit accesses
    int16_t chroma_dc[2][4];
    // ---- ME ----
    int mode, ref;
    const uint8_t* ref_y;
    alignas(4) int16_t mv_cur[4][4][2];   // (*MvL0) of the current macroblock during the search
    int8_t ref_cur[4];         // RefIdxL0 of the current macroblock (stale during the search)
    int16_t mvp[4][4][2];
    int16_t best_mv[4][4][2];
    double best_cost[4][4];
    int best_dist[4][4], best_sctr[4][4], best_cbp[4][4];
    int probably_pskip;
    // reference tile origin (tile_x0 is a multiple of 16); samples follow the per-sample clamp of interpol.c:108-131
    int tile_x0, tile_y0, tile_ref, tile_valid;
    int tile_phase;                                   // parity of the next completion of tile_mbar
    // one evaluation step
    int c_begin, c_end;        // candidates evaluated by the current CMD_ME_EVAL
    int part_ox, part_oy, part_w, part_h;
    int part_mask;                     // luma4x4BlkIdx bits of the blocks inside the current partition (set_part)
    int counts_only;           // the step's trials need TotalCoeff / TrailingOnes / Single_ctr only (me_step)
    uint32_t cand_mv[HLB_MAXC];        // candidate vectors of the step: (uint16)mvx | (uint16)mvy << 16
    uint8_t cand_pat[HLB_MAXC];        // their pattern indices
    uint32_t r_val[HLB_MAXC][16];      // trial result words by (candidate, luma4x4BlkIdx): dist:12 | bits_rest:10 | TotalCoeff:5 | TrailingOnes:2 | lone Single_ctr:2
    uint8_t eff[HLB_MAXC][16];
    uint8_t blk_coded[16];             // final reconstruction: block has non-zero levels
    uint32_t tail_mv[28];              // candidate list of me_pskip_tail
    int tail_n;
    int32_t c_dist[HLB_MAXC], c_rbc[HLB_MAXC], c_sctr[HLB_MAXC], c_cbp[HLB_MAXC];
    double c_cost[HLB_MAXC];
    int step_last;             // ((c+1) << 12 | k << 8 | Single_ctr) of the last non-zero trial block of the step, -1 if none
    int bw_log2, nblk_log2;    // log2 of the partition width / size in 4x4 blocks
    // ---- reconstruction ----
    int fin_mode, fin_sub[4];          // committed geometry (part_mode, sub_mode[])
    int16_t fin_mv[4][4][2];
    int8_t fin_ref[4];
    alignas(4) uint8_t pred_y[256];
    alignas(4) uint8_t pred_c[2][64];
    alignas(4) uint8_t rec_y[256];
    alignas(4) uint8_t rec_c[2][64];
    alignas(4) int16_t luma_level[16][16];
    int luma_skip_residual;            // Single_ctr_luma < 6 (rdo.c:2419)
    int cbp_luma4x4;
    int cbp_ac[2], cbp_dc[2];
    int mb_is_intra;
    int32_t c_dccoef[2][4];
    alignas(4) uint8_t c_acnz[2][4], c_sc[2][4], c_tc[2][4], c_resnz[2][4];   // c_acnz rows are read as words
    // ---- intra ----
    int i16_mode, i16_cbp4x4, i4_cbp4x4, intra_chroma_mode;
    int16_t i16_dc[16];
    alignas(4) int16_t i16_ac[16][16];
    uint8_t i4_mode[16], prev_i4[16], rem_i4[16];
    int32_t nbr_y[41], nbr_c[2][17];   // reconstructed samples around the macroblock (intra_fetch_borders)
    int32_t p33[33];
    int32_t p17[2][17];
    int32_t p13[13];
    // The trial memo of the inter search and the scratch of the intra trials are never live together (the intra decision of a P macroblock
    // runs after its inter search, an I macroblock has no search): they share storage.
    union {
        alignas(8) unsigned long long memo[16][HLB_MEMO_SLOTS];   // trial memo of the current macroblock (see me_phase_trial)
        struct {
            // scratch of the intra trials
    int16_t t_ac[4][16][16];
    int32_t t_dcw[4][16];
    int16_t t_dc[4][16];
    uint8_t t_nz[4][16], t_tc[4][16], t_t1[4][16], t_sc[4][16];
    uint16_t t_bits[4][16];
    int32_t t_dist[4][16];
    int t_mode_ok[4], t_cbp[4], t_rate[4], t_sctr[4], t_dcbits[4];
    uint8_t t_pred[4][256];
    int i4_blk;
    int32_t q_dist[9];
    uint16_t q_bits[9];
    uint8_t q_nz[9], q_tc[9], q_t1[9], q_sc[9], q_ok[9], q_res0[9];
    alignas(4) int16_t q_lv[9][16];   // 32-bit accesses
    alignas(4) uint8_t q_pred[9][16];
        };
    };
};

// ------------------------------------------------------------------------------------------------------------------
// Neighbour derivation for motion data (6.4.10.7 + 8.4.1.3.2; mb.c:426-541, utils.c:854-963)
// ------------------------------------------------------------------------------------------------------------------
struct NbMotion { int avail; int ref; int mvx, mvy; };

HLB_FN NbMotion nb_motion_at(const MbWork& w, const FrameCtx& f, int xN, int yN, int cur_part, int cur_sub)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    NbMotion r;
    r.avail = 0; r.ref = -1; r.mvx = 0; r.mvy = 0;
    if (xN >= 0 && xN <= 15 && yN >= 0 && yN <= 15) {  // inside the current macroblock: geometry of the mode being searched
        uint8_t sm[4];
        const int s = w.mode <= 3 ? 0 : w.mode - 3;
        sm[0] = sm[1] = sm[2] = sm[3] = (uint8_t)s;
        int p, q;
        part_at(w.mode < 3 ? w.mode : 3, sm, xN, yN, p, q);
        if (p > cur_part || (p == cur_part && q > cur_sub)) return r;  // not yet searched
        r.avail = 1; r.ref = w.ref_cur[p]; r.mvx = w.mv_cur[p][q][0]; r.mvy = w.mv_cur[p][q][1];
        return r;
    }
    int k;
    if (yN < 0 && xN >= 0 && xN <= 15) { if (!w.availB) return r; k = 2; }
    else if (yN < 0 && xN > 15) { if (!w.availC) return r; k = 3; }
    else if (yN < 0 && xN < 0) { if (!w.availD) return r; k = 4; }
    else if (xN < 0 && yN >= 0 && yN <= 15) { if (!w.availA) return r; k = 1; }
    else return r;
    const MbState& s = *(const MbState*)w.nbw[k];   // only the head (HLB_NB_WORDS words) is valid
    r.avail = 1;
    if (s.kind == MBK_I16 || s.kind == MBK_I4) return r;  // intra: ref -1, mv 0
    int p, q;
    part_at(s.part_mode, s.sub_mode, (xN + 16) & 15, (yN + 16) & 15, p, q);
    r.ref = s.ref_idx[p]; r.mvx = s.mv[p][q][0]; r.mvy = s.mv[p][q][1];
    return r;
}

HLB_HD int median3(int a, int b, int c)
{
    const int mn = a < b ? (a < c ? a : c) : (b < c ? b : c), mx = a > b ? (a > c ? a : c) : (b > c ? b : c);
    return a + b + c - mn - mx;
}

// 8.4.1.3 (utils.c:751-798).  (ox,oy,pw) = origin and predPartWidth of the partition under the current search mode
HLB_FN void derive_mvp(const MbWork& w, const FrameCtx& f, int part, int sub, int ref, int& mx, int& my)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    int ox, oy, pw, ph;
    mode_rect(w.mode, part, sub, ox, oy, pw, ph);
    NbMotion A = nb_motion_at(w, f, ox - 1, oy, part, sub);
    NbMotion B = nb_motion_at(w, f, ox, oy - 1, part, sub);
    NbMotion C = nb_motion_at(w, f, ox + pw, oy - 1, part, sub);
    if (!C.avail) C = nb_motion_at(w, f, ox - 1, oy - 1, part, sub);
    const int mw = mode_part_w(w.mode), mh = mode_part_h(w.mode);
    if (mw == 16 && mh == 8 && part == 0 && B.ref == ref) { mx = B.mvx; my = B.mvy; return; }
    if (mw == 16 && mh == 8 && part == 1 && A.ref == ref) { mx = A.mvx; my = A.mvy; return; }
    if (mw == 8 && mh == 16 && part == 0 && A.ref == ref) { mx = A.mvx; my = A.mvy; return; }
    if (mw == 8 && mh == 16 && part == 1 && C.ref == ref) { mx = C.mvx; my = C.mvy; return; }
    if (!B.avail && !C.avail && A.avail) { B = A; C = A; }
    if (A.ref == ref && B.ref != ref && C.ref != ref) { mx = A.mvx; my = A.mvy; }
    else if (B.ref == ref && C.ref != ref && A.ref != ref) { mx = B.mvx; my = B.mvy; }
    else if (C.ref == ref && B.ref != ref && A.ref != ref) { mx = C.mvx; my = C.mvy; }
    else { mx = median3(A.mvx, B.mvx, C.mvx); my = median3(A.mvy, B.mvy, C.mvy); }
}

// 8.4.1.1 (utils.c:709-748); only ever called with the 16x16 geometry
// 8.4.1.1: the P_Skip vector is zero when A or B is unavailable or is a zero vector on reference 0, else the 16x16 predictor on reference 0
HLB_HD bool pskip_mv_is_zero(const MbWork& w, const FrameCtx& f)
{
    const NbMotion A = nb_motion_at(w, f, -1, 0, 0, 0), B = nb_motion_at(w, f, 0, -1, 0, 0);
    return !A.avail || !B.avail || (A.ref == 0 && A.mvx == 0 && A.mvy == 0) || (B.ref == 0 && B.mvx == 0 && B.mvy == 0);
}
HLB_FN void derive_pskip_mv(const MbWork& w, const FrameCtx& f, int& mx, int& my)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (pskip_mv_is_zero(w, f)) { mx = my = 0; return; }
    derive_mvp(w, f, 0, 0, 0, mx, my);
}

// ------------------------------------------------------------------------------------------------------------------
// nC inputs coming from the neighbouring macroblocks (residual.c:698-741)
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int nb_count(const MbState& s, int blk)
{
    if (s.kind == MBK_PSKIP) return 0;
    if (((s.cbp_luma >> (blk >> 2)) & 1) == 0) return 0;
    return s.tc_luma[blk];
}
HLB_HD int nc_from(int nA, int nB)
{
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    if (nA >= 0) return nA;
    if (nB >= 0) return nB;
    return 0;
}
// nC of luma block `blk` of the current macroblock given per-block counts `cnt` (TotalCoeffsLuma view to use for in-MB neighbours)
HLB_HD int luma_nc(const MbWork& w, const uint8_t* cnt, int blk)
{
    const int x = blk_x(blk), y = blk_y(blk);
    int nA, nB;
    if (x > 0) { const int a = blk_idx_from_xy(x - 4, y); nA = ((w.cbp_gate >> (a >> 2)) & 1) ? cnt[a] : 0; }
    else nA = w.extA[blk];
    if (y > 0) { const int b = blk_idx_from_xy(x, y - 4); nB = ((w.cbp_gate >> (b >> 2)) & 1) ? cnt[b] : 0; }
    else nB = w.extB[blk];
    return nc_from(nA, nB);
}

// ------------------------------------------------------------------------------------------------------------------
// Sample access
// ------------------------------------------------------------------------------------------------------------------
// 9x9 window (rows/cols -2..+6 around a 4x4 block whose pixel (0,0) sits at (X,Y)), per-sample clamp (interpol.c:108-131)
HLB_HD void ref_window9(const uint8_t* plane, int W, int H, int X, int Y, uint8_t t[81])
{
    if (X >= 2 && Y >= 2 && X + 7 <= W && Y + 7 <= H) {
        const uint8_t* p = plane + (Y - 2) * W + (X - 2);
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
            for (int c = 0; c < 9; ++c) t[r * 9 + c] = HLB_LDG(p + r * W + c);
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int y = clip3(0, H - 1, Y - 2 + r);
#pragma unroll
            for (int c = 0; c < 9; ++c) t[r * 9 + c] = HLB_LDG(plane + y * W + clip3(0, W - 1, X - 2 + c));
        }
    }
}

// luma prediction of the 4x4 block at (bx,by) (MB coordinates) belonging to a partition with origin (ox,oy) and motion (mvx,mvy).
// `win` = 9x9 staging area of the calling lane (shared memory on the GPU): rows/cols -2..+6 around the block, fetched with the
// reference's per-sample clamp (interpol.c:108-131) by a row loop that is not unrolled (runs once or twice per macroblock).
HLB_HD void pred_luma_4x4(const FrameCtx& f, const uint8_t* ref_y, int mbx, int mby, int ox, int oy, int bx, int by, int mvx, int mvy, uint8_t* win, uint8_t out[16])
{
    const int X = clip3(-17, f.W + 17, mbx * 16 + ox + (mvx >> 2)) + (bx - ox);   // origin clip of pred_inter.c:395 applies to the partition
    const int Y = clip3(-17, f.H + 17, mby * 16 + oy + (mvy >> 2)) + (by - oy);
    const int W = f.W, Hm1 = f.H - 1, Wm1 = f.W - 1;
    int xc[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) xc[c] = clip3(0, Wm1, X - 2 + c);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const uint8_t* row = ref_y + clip3(0, Hm1, Y - 2 + r) * W;
#pragma unroll
        for (int c = 0; c < 9; ++c) win[r * 9 + c] = HLB_LDG(row + xc[c]);
    }
    interp_luma_4x4(win + 20, 9, mvx & 3, mvy & 3, out);
}

// ------------------------------------------------------------------------------------------------------------------
// CMD_ME_EVAL: the trial encodes of one search step (me_ds.c:527-688 for every candidate of the step)
// ------------------------------------------------------------------------------------------------------------------
HLB_HD int mv_x(uint32_t mv) { return (int)(int16_t)(uint16_t)(mv & 0xffffu); }
HLB_HD int mv_y(uint32_t mv) { return (int)(int16_t)(uint16_t)(mv >> 16); }
HLB_HD uint32_t mv_pack(int x, int y) { return (uint32_t)(uint16_t)x | ((uint32_t)(uint16_t)y << 16); }
// origin (after the partition-origin clip of pred_inter.c:395-396) of a candidate's partition in picture coordinates
HLB_HD void cand_origin(const MbWork& w, const FrameCtx& f, uint32_t mv, int& X, int& Y)
{
    X = clip3(-17, f.W + 17, w.mbx * 16 + w.part_ox + (mv_x(mv) >> 2));
    Y = clip3(-17, f.H + 17, w.mby * 16 + w.part_oy + (mv_y(mv) >> 2));
}
// The reference tile: tile[j * HLB_TILE_W + i] = ref_y[clampY(tile_y0 + j)][clampX(tile_x0 + i)].
// Device: ONE TMA 2D tile load (CMD_TILE_TMA: cp.async.bulk.tensor, completion on an mbarrier; samples outside the picture arrive as zeros), then, only for tiles
//         that cross the picture edge, the border fix-up (CMD_TILE_FIX): the reference clamps every sample coordinate (interpol.c:108-131), so out-of-picture
//         columns and then rows of the tile are replicas of the nearest in-picture column / row.
// A tile without a single in-picture column or row (vectors far outside the picture) and the CPU harness take the plain clamped loop (CMD_TILE).
HLB_HD bool tile_uses_tma(const MbWork& w, const FrameCtx& f)
{
#if defined(__CUDA_ARCH__)
    return f.ref_tmap[w.ref] != nullptr && w.tile_x0 > -HLB_TILE_W && w.tile_x0 < f.W && w.tile_y0 > -HLB_TILE_H && w.tile_y0 < f.H;
#else
    (void)w; (void)f;
    return false;
#endif
}
// CMD_TILE_TMA (one lane): issue the load, wait for its bytes on the mbarrier
HLB_FN void phase_tile_tma(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
#if defined(__CUDA_ARCH__)
    if (lane != 0) return;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&w.tile_mbar), dst = (uint32_t)__cvta_generic_to_shared(w.tile);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of the tile are ordered before the async-proxy write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(HLB_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(f.ref_tmap[w.ref]), "r"(w.tile_x0), "r"(w.tile_y0), "r"(bar) : "memory");
    const uint32_t parity = (uint32_t)w.tile_phase & 1u;
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 16)) { w.stuck = 2; break; }   // a tile load that never completes is reported through the watchdog (hlb200_slice_status), not waited for
    }
    w.tile_phase ^= 1;
#else
    (void)w; (void)f; (void)lane;
#endif
}
// CMD_TILE_FIX: phase 0 = out-of-picture columns of the in-picture rows, phase 1 = out-of-picture rows (word-wise, from the now column-complete nearest row)
HLB_HD bool tile_crosses_edge(const MbWork& w, const FrameCtx& f) { return w.tile_x0 < 0 || w.tile_y0 < 0 || w.tile_x0 + HLB_TILE_W > f.W || w.tile_y0 + HLB_TILE_H > f.H; }
HLB_FN void phase_tile_fix(MbWork& w, const FrameCtx& f, int phase, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int nl = w.arg0_lanes, x0 = w.tile_x0, y0 = w.tile_y0;
    const int cx0 = x0 < 0 ? -x0 : 0, cx1 = x0 + HLB_TILE_W > f.W ? f.W - x0 : HLB_TILE_W;           // in-picture columns [cx0, cx1)
    const int ry0 = y0 < 0 ? -y0 : 0, ry1 = y0 + HLB_TILE_H > f.H ? f.H - y0 : HLB_TILE_H;           // in-picture rows    [ry0, ry1)
    if (phase == 0 && (cx0 > 0 || cx1 < HLB_TILE_W)) {
        const int nbad = cx0 + (HLB_TILE_W - cx1);
#pragma unroll 1
        for (int i = lane; i < (ry1 - ry0) * nbad; i += nl) {
            const int r = ry0 + i / nbad, k = i % nbad, c = k < cx0 ? k : cx1 + (k - cx0);
            w.tile[r * HLB_TILE_W + c] = w.tile[r * HLB_TILE_W + (k < cx0 ? cx0 : cx1 - 1)];
        }
    }
    if (phase == 1 && (ry0 > 0 || ry1 < HLB_TILE_H)) {
        const int nbad = ry0 + (HLB_TILE_H - ry1);
#pragma unroll 1
        for (int i = lane; i < nbad * (HLB_TILE_W / 4); i += nl) {
            const int k = i / (HLB_TILE_W / 4), q = i % (HLB_TILE_W / 4), r = k < ry0 ? k : ry1 + (k - ry0);
            ((uint32_t*)w.tile)[r * (HLB_TILE_W / 4) + q] = ((const uint32_t*)w.tile)[(k < ry0 ? ry0 : ry1 - 1) * (HLB_TILE_W / 4) + q];
        }
    }
}
// CMD_TILE: the plain clamped loads (tiles without an in-picture sample, launches without tensor maps, the CPU harness)
HLB_FN void phase_tile_load(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int nl = w.arg0_lanes;
    const int W = f.W, Hm1 = f.H - 1, Wm1 = f.W - 1, x0 = w.tile_x0, y0 = w.tile_y0;
    const uint8_t* plane = w.ref_y;
    // eight independent loads in flight per lane
#pragma unroll 1
    for (int i0 = lane; i0 < HLB_TILE_BYTES; i0 += 8 * nl) {
        uint8_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nl;
            const int ty = i / HLB_TILE_W, tx = i - ty * HLB_TILE_W;
            const int y = clip3(0, Hm1, y0 + ty), x = clip3(0, Wm1, x0 + tx);
            v[u] = i < HLB_TILE_BYTES ? HLB_LDG(plane + y * W + x) : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nl;
            if (i < HLB_TILE_BYTES) w.tile[i] = v[u];
        }
    }
}
// ---- trial memo ----
// A trial encode is a pure function of (4x4 block, motion vector): prediction, levels, CAVLC counts, reconstruction error.  Only the
// coeff_token term depends on encoder history, and that is resolved afterwards from (TotalCoeff, TrailingOnes) by the scan / cost phases.
// The seven partition modes of a macroblock walk almost the same diamonds around almost the same vectors, so most trials of modes 1..6 repeat
// a (block, vector) pair an earlier search already encoded: on the reference trajectory 36 % (G1 1080p, mostly PSkip) to 87 % (G2 CIF) of all
// trials, and 63 % / 95 % of the search STEPS consist of repeats only (measured with the CPU harness).  Each block keeps a small open-addressed
// table (linear probing, no eviction) in shared memory, reset per macroblock; a hit skips interpolation, transform, quantisation, CAVLC
// counting and reconstruction.  Entry = key (mvx | mvy << 16) | value << 32, value = dist:12 | bits_rest:10 | TotalCoeff:5 | TrailingOnes:2 |
// lone-coefficient Single_ctr:2; written with one 64-bit store so concurrent lanes see it whole or not at all.
#define HLB_MEMO_EMPTY_KEY 0x80008000u   /* mv = (-32768, -32768) cannot occur */
#define HLB_MEMO_PROBES 8
#if defined(__CUDA_ARCH__)
#define HLB_MEMO_CAS(p, old, val) (atomicCAS((p), (old), (val)) == (old))
#define HLB_ATOMIC_MAX(p, v) atomicMax((p), (v))
#else
#define HLB_MEMO_CAS(p, old, val) (*(p) == (old) ? (*(p) = (val), true) : false)
#define HLB_ATOMIC_MAX(p, v) (*(p) = *(p) > (v) ? *(p) : (v))
#endif
#if !defined(__CUDACC__)
// CPU harness only (tools/emu --memo-stats): how many trials / search steps of the reference trajectory are repeats
struct MemoStats { long trials, hits, steps, full_hit_steps, step_trials, step_hits; };
static MemoStats g_memo_stats = {0, 0, 0, 0, 0, 0};
#define HLB_MEMO_COUNT(hit) do { g_memo_stats.trials++; g_memo_stats.step_trials++; if (hit) { g_memo_stats.hits++; g_memo_stats.step_hits++; } } while (0)
#define HLB_MEMO_STEP_BEGIN() do { g_memo_stats.step_trials = g_memo_stats.step_hits = 0; } while (0)
#define HLB_MEMO_STEP_END() do { g_memo_stats.steps++; if (g_memo_stats.step_trials == g_memo_stats.step_hits) g_memo_stats.full_hit_steps++; } while (0)
#else
#define HLB_MEMO_COUNT(hit) do { } while (0)
#define HLB_MEMO_STEP_BEGIN() do { } while (0)
#define HLB_MEMO_STEP_END() do { } while (0)
#endif
HLB_HD void memo_reset(MbWork& w, int lane, int nl)
{
    unsigned long long* t = &w.memo[0][0];
#pragma unroll 1
    for (int i = lane; i < 16 * HLB_MEMO_SLOTS; i += nl) t[i] = (unsigned long long)HLB_MEMO_EMPTY_KEY;
}
// CMD_ME_EVAL: lane = (candidate - c_begin) * nblk + k, k = raster index of the 4x4 block inside the partition.  Result word -> w.r_val[c][blk].
// The arithmetic is hlb_fast.cuh's: prediction straight out of the shared-memory tile as packed rows, dot-product transform, packed SAD.
HLB_FN void me_phase_trial(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int c = w.c_begin + (lane >> w.nblk_log2), k = lane & ((1 << w.nblk_log2) - 1);
    if (c >= w.c_end) return;
    const int bx = w.part_ox + ((k & ((1 << w.bw_log2) - 1)) << 2), by = w.part_oy + ((k >> w.bw_log2) << 2);
    const int blk = blk_idx_from_xy(bx, by);
    const uint32_t key = w.cand_mv[c];
    const int mvx = mv_x(key), mvy = mv_y(key);
    const bool counts_only = w.counts_only != 0;
    unsigned long long* tab = w.memo[blk];
    uint32_t val = 0;
    bool hit = false;
    int ins = -1;
    if (!counts_only) {   // a search that can no longer improve needs no distortion / bit count: nothing to remember, nothing worth looking up
        const int h0 = (mvx * 5 + mvy * 23) & (HLB_MEMO_SLOTS - 1);
#pragma unroll 1
        for (int p = 0; p < HLB_MEMO_PROBES; ++p) {
            const int j = (h0 + p) & (HLB_MEMO_SLOTS - 1);
            const unsigned long long e = *(volatile unsigned long long*)&tab[j];
            if ((uint32_t)e == key) { hit = true; val = (uint32_t)(e >> 32); break; }
            if ((uint32_t)e == HLB_MEMO_EMPTY_KEY) { ins = j; break; }
        }
        HLB_MEMO_COUNT(hit);
    }
    if (!hit) {
        int X, Y;
        cand_origin(w, f, key, X, Y);
        const Rows4 p = fast_pred_luma((const uint32_t*)w.tile, HLB_TILE_W / 4, X + (bx - w.part_ox) - w.tile_x0, Y + (by - w.part_oy) - w.tile_y0, mvx & 3, mvy & 3);
        Rows4 s;
#pragma unroll
        for (int r = 0; r < 4; ++r) s.r[r] = ((const uint32_t*)w.src_y)[((by + r) * 16 + bx) >> 2];   // bx is a multiple of 4
        val = fast_trial(s, p, f.qk, counts_only);
        if (ins >= 0) (void)HLB_MEMO_CAS(&tab[ins], (unsigned long long)HLB_MEMO_EMPTY_KEY, ((unsigned long long)val << 32) | key);   // lost races / full rows: not cached
    }
    w.r_val[c][blk] = val;
}
HLB_HD bool blk_in_part(const MbWork& w, int blk) { return (w.part_mask >> blk) & 1; }   // set_part: kPartMask
// block k (raster index inside the current partition = evaluation order) -> luma4x4BlkIdx
HLB_HD int part_blk(const MbWork& w, int k) { return blk_idx_from_xy(w.part_ox + ((k & ((1 << w.bw_log2) - 1)) << 2), w.part_oy + ((k >> w.bw_log2) << 2)); }
// per block OF THE PARTITION: TotalCoeffsLuma[blk] as every candidate of the step sees / leaves it, in evaluation order (residual.c:796-806: only non-zero
// blocks are stored), and the last non-zero trial block of the step in evaluation order (its Single_ctr stays in pc_esd->rdo.Single_ctr, residual.c:882).
// Blocks outside the partition keep w.tc[] throughout the step.
HLB_HD void me_scan_block(MbWork& w, int k, int n, bool keep_eff)
{
    const int blk = part_blk(w, k);
    int e = w.tc[blk], last = -1;
#ifdef HLB_SCAN_UNROLLED
    uint32_t v[HLB_MAXC];
#pragma unroll
    for (int c = 0; c < HLB_MAXC; ++c) v[c] = c < n ? w.r_val[c][blk] : 0u;   // independent loads, then the serial "last non-zero wins" in registers
#pragma unroll
    for (int c = 0; c < HLB_MAXC; ++c) {
        const int tcv = (int)((v[c] >> 22) & 31u);
        if (tcv) {
            e = tcv;
            last = ((c + 1) << 12) | ((tcv == 1 && ((v[c] >> 27) & 3u) == 1) ? (int)((v[c] >> 29) & 3u) : 9);
        }
        if (keep_eff && c < n) w.eff[c][blk] = (uint8_t)e;
    }
#else
    // rolled over the step's candidates (n <= 9, often 4-5 after pruning): the kernel is instruction-fetch bound, the nine unrolled copies cost more than the
    // shared-memory latency they hide
    uint32_t lastv = 0;
    int lastc = -1;
    const uint32_t* rv = &w.r_val[0][blk];
    uint8_t* ef = &w.eff[0][blk];
#pragma unroll 1
    for (int c = 0; c < n; ++c, rv += 16, ef += 16) {
        const uint32_t v = *rv;
        const int tcv = (int)((v >> 22) & 31u);
        if (tcv) { e = tcv; lastc = c; lastv = v; }
        if (keep_eff) *ef = (uint8_t)e;
    }
    if (lastc >= 0) last = ((lastc + 1) << 12) | ((e == 1 && ((lastv >> 27) & 3u) == 1) ? (int)((lastv >> 29) & 3u) : 9);
#endif
    w.tc[blk] = (uint8_t)e;
    if (last >= 0) HLB_ATOMIC_MAX(&w.step_last, last | (k << 8));
}
// TotalCoeffsLuma of neighbour block `a` as candidate c sees it: its own evolving value inside the partition, the macroblock's current one outside
HLB_HD int me_cnt(const MbWork& w, int c, int a) { return blk_in_part(w, a) ? w.eff[c][a] : w.tc[a]; }
// contribution of one (candidate, block) pair to the candidate's sums: lo = dist | bits << 16 (coeff_token resolved with the history-exact nC), hi = Single_ctr | CBP bit << 8
HLB_HD void me_cost_term(const MbWork& w, int c, int k, uint32_t& lo, uint32_t& hi)
{
    const int blk = part_blk(w, k);
    const uint32_t v = w.r_val[c][blk];
    lo = v & 4095u; hi = 0;
    const int tcv = (int)((v >> 22) & 31u);
    if (tcv) {
        const int t1 = (int)((v >> 27) & 3u), x = blk_x(blk), y = blk_y(blk);
        int nA, nB;
        if (x > 0) { const int a = blk_idx_from_xy(x - 4, y); nA = ((w.cbp_gate >> (a >> 2)) & 1) ? me_cnt(w, c, a) : 0; }
        else nA = w.extA[blk];
        if (y > 0) { const int b = blk_idx_from_xy(x, y - 4); nB = ((w.cbp_gate >> (b >> 2)) & 1) ? me_cnt(w, c, b) : 0; }
        else nB = w.extB[blk];
        lo |= (uint32_t)((int)((v >> 12) & 1023u) + coeff_token_len(nc_from(nA, nB), tcv, t1)) << 16;
        hi = (uint32_t)((tcv == 1 && t1 == 1) ? (int)((v >> 29) & 3u) : 9) | (1u << (8 + blk));
    }
}
// per candidate: distortion, bits, Single_ctr, CBP bits, RD cost (me_ds.c:287,297,345) from the sums of its blocks' terms
HLB_HD void me_cost_finish(MbWork& w, const FrameCtx& f, int c, uint32_t lo, uint32_t hi, int px, int py)
{
    const int dist = (int)(lo & 0xffffu), rbc = (int)(lo >> 16);
    const uint32_t mv = w.cand_mv[c];
    w.c_dist[c] = dist; w.c_rbc[c] = rbc; w.c_sctr[c] = (int)(hi & 255u); w.c_cbp[c] = (int)(hi >> 8);
    w.c_cost[c] = dist + ((rbc + se_len(mv_x(mv) - px) + se_len(mv_y(mv) - py)) * f.lambda);
}
// all candidates of a step: one lane per (candidate, block), the blocks of a candidate summed over its (aligned) group of lanes by a shuffle butterfly
template <class X>
HLB_HD void me_cost_all(X& x, MbWork& w, const FrameCtx& f, int n, int px, int py)
{
    const int total = n << w.nblk_log2, kmask = (1 << w.nblk_log2) - 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
    for (int i0 = 0; i0 < total; i0 += 32) {   // the master warp: 32 lanes, every lane takes part in the shuffles
        const int i = i0 + x.lane(), c = i >> w.nblk_log2, k = i & kmask;
        uint32_t lo = 0, hi = 0;
        if (i < total) me_cost_term(w, c, k, lo, hi);
#pragma unroll 1
        for (int o = (1 << w.nblk_log2) >> 1; o; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
        if (i < total && k == 0) me_cost_finish(w, f, c, lo, hi, px, py);
    }
#else
    for (int c = 0; c < n; ++c) {
        uint32_t lo = 0, hi = 0;
        for (int k = 0; k <= kmask; ++k) { uint32_t a, b; me_cost_term(w, c, k, a, b); lo += a; hi += b; }
        me_cost_finish(w, f, c, lo, hi, px, py);
    }
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// Search of one mode (me_ds.c:104-477)
// ------------------------------------------------------------------------------------------------------------------
HLB_TABLE static const int8_t kDsp[3][9][2] = {   // [shift: 0 quarter, 1 half, 2 integer][pattern index]
    {{-1, 1}, {0, 1}, {1, 1}, {-1, 0}, {0, 0}, {1, 0}, {-1, -1}, {0, -1}, {1, -1}},
    {{0, 1}, {-1, 0}, {0, -1}, {1, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}},
    {{0, 2}, {-1, 1}, {1, 1}, {-2, 0}, {0, 0}, {2, 0}, {-1, -1}, {1, -1}, {0, -2}}};
// points skipped in the next iteration after the best point `idx` (me_ds.c:384-465), as bit masks over pattern indices
HLB_TABLE static const uint16_t kPrune[3][9] = {{0x1B0, 0x1F8, 0x03F, 0x1B6, 0x000, 0x0DB, 0x036, 0x03F, 0x01B},
                                                {0x014, 0x018, 0x011, 0x012, 0x000, 0, 0, 0, 0},
                                                {0x1D0, 0x130, 0x1DA, 0x0B4, 0x000, 0x05A, 0x0B7, 0x05F, 0x017}};
HLB_TABLE static const uint8_t kHeaderBits[7] = {3, 5, 5, 11, 19, 19, 27};

// luma4x4BlkIdx bits of the 4x4 blocks inside partition (mode, part, sub) -- mode_rect() tabulated
HLB_TABLE static const uint16_t kPartMask[7][4][4] = {
    {{0xFFFF, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}},
    {{0x00FF, 0x0000, 0x0000, 0x0000}, {0xFF00, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}},
    {{0x0F0F, 0x0000, 0x0000, 0x0000}, {0xF0F0, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}, {0x0000, 0x0000, 0x0000, 0x0000}},
    {{0x000F, 0x0000, 0x0000, 0x0000}, {0x00F0, 0x0000, 0x0000, 0x0000}, {0x0F00, 0x0000, 0x0000, 0x0000}, {0xF000, 0x0000, 0x0000, 0x0000}},
    {{0x0003, 0x000C, 0x0000, 0x0000}, {0x0030, 0x00C0, 0x0000, 0x0000}, {0x0300, 0x0C00, 0x0000, 0x0000}, {0x3000, 0xC000, 0x0000, 0x0000}},
    {{0x0005, 0x000A, 0x0000, 0x0000}, {0x0050, 0x00A0, 0x0000, 0x0000}, {0x0500, 0x0A00, 0x0000, 0x0000}, {0x5000, 0xA000, 0x0000, 0x0000}},
    {{0x0001, 0x0002, 0x0004, 0x0008}, {0x0010, 0x0020, 0x0040, 0x0080}, {0x0100, 0x0200, 0x0400, 0x0800}, {0x1000, 0x2000, 0x4000, 0x8000}}};
HLB_HD void set_part(MbWork& w, int mode, int p, int s)
{
    mode_rect(mode, p, s, w.part_ox, w.part_oy, w.part_w, w.part_h);
    w.part_mask = kPartMask[mode][p][s];
    w.bw_log2 = w.part_w == 16 ? 2 : (w.part_w == 8 ? 1 : 0);
    w.nblk_log2 = w.bw_log2 + (w.part_h == 16 ? 2 : (w.part_h == 8 ? 1 : 0));
}

// Trial encodes of candidates [c0, c1) (they fit the reference tile together).  (bx0..by1) = bounding box of the samples they read.
template <class X>
HLB_FN void me_eval_range(X& x, MbWork& w, const FrameCtx& f, int c0, int c1, int bx0, int by0, int bx1, int by1)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (!(w.tile_valid && w.tile_ref == w.ref && bx0 >= w.tile_x0 && by0 >= w.tile_y0 && bx1 <= w.tile_x0 + HLB_TILE_W && by1 <= w.tile_y0 + HLB_TILE_H)) {
        x.sync();   // every lane has evaluated the condition before the tile origin changes
        // centred as far as the 16-sample alignment of the origin allows: origin <= bx0 - m and origin + 64 >= bx1 for every origin in (bx0 - m - 16, bx0 - m]
        const int span = bx1 - bx0, m = (HLB_TILE_W - span) >> 1 < HLB_TILE_W - 15 - span ? (HLB_TILE_W - span) >> 1 : HLB_TILE_W - 15 - span;
        w.tile_x0 = (bx0 - m) & ~15; w.tile_y0 = by0 - ((HLB_TILE_H - (by1 - by0)) >> 1);
        w.tile_ref = w.ref; w.tile_valid = 1;
        HLB_LAP(w, 2);
        x.sync();
        if (tile_uses_tma(w, f)) {
            x.run(CMD_TILE_TMA, 1);
            if (tile_crosses_edge(w, f)) x.run(CMD_TILE_FIX, 64);
        } else x.run(CMD_TILE, HLB_MB_LANES);
        HLB_LAP(w, 3);
    }
    w.c_begin = c0; w.c_end = c1;
    HLB_LAP(w, 2);
    HLB_MEMO_STEP_BEGIN();
    x.trials((c1 - c0) << w.nblk_log2);
    HLB_MEMO_STEP_END();
    HLB_LAP(w, 4);
}
// One search step: evaluates the n candidates w.cand_mv[0..n) of the current partition in order; leaves TotalCoeffsLuma / the Single_ctr chain
// as the reference's sequential evaluation would and, unless counts_only, per-candidate dist / rbc / sctr / cbp and the RD cost
// (me_ds.c:287,297,345: dist + (rbc + mvd bits) * lambda) in w.c_*.  counts_only = the search can no longer improve (its best cost is 0: costs are
// never negative and a candidate only wins with a strictly smaller one), so nothing but the rate state the trials leave behind matters.
template <class X>
HLB_FN void me_step(X& x, MbWork& w, const FrameCtx& f, int n, int px, int py, bool counts_only, bool count_work = true)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    HLB_LAP(w, 1);
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, iops = 0;
#pragma unroll 1
    for (int c = x.lane(); c < n; c += x.nlanes()) {
        const uint32_t mv = w.cand_mv[c];
        int OX, OY;
        cand_origin(w, f, mv, OX, OY);   // the origin clip is monotonic, so the box of the clipped origins is the clipped box
        mnx = OX < mnx ? OX : mnx; mxx = OX > mxx ? OX : mxx; mny = OY < mny ? OY : mny; mxy = OY > mxy ? OY : mxy;
        // interpolation ops per 4x4 block by fractional class
        const int xf = mv_x(mv) & 3, yf = mv_y(mv) & 3;
        iops += (xf == 0 && yf == 0) ? 0 : ((xf == 0 || yf == 0) ? (((xf | yf) == 2) ? 176 : 208) : (((xf & 1) && (yf & 1)) ? 352 : 880));
    }
    mnx = x.reduce_min(mnx); mny = x.reduce_min(mny); mxx = x.reduce_max(mxx); mxy = x.reduce_max(mxy); iops = x.reduce_add(iops);
    if (x.lane() == 0) {
        w.counts_only = counts_only ? 1 : 0; w.step_last = -1;
        if (count_work) { w.stat_trials += (unsigned)(n << w.nblk_log2); w.stat_cands += (unsigned)n; w.stat_interp += (unsigned)(iops << w.nblk_log2); }
    }
    if (mxx - mnx + w.part_w + 5 <= HLB_TILE_SPAN_W && mxy - mny + w.part_h + 5 <= HLB_TILE_H) me_eval_range(x, w, f, 0, n, mnx - 2, mny - 2, mxx + w.part_w + 3, mxy + w.part_h + 3);
    else {
#pragma unroll 1
        for (int c = 0; c < n; ++c) {   // windows too far apart for one tile: one by one, same results
            int PX, PY;
            cand_origin(w, f, w.cand_mv[c], PX, PY);
            me_eval_range(x, w, f, c, c + 1, PX - 2, PY - 2, PX + w.part_w + 3, PY + w.part_h + 3);
        }
    }
#pragma unroll 1
    for (int k = x.lane(); k < (1 << w.nblk_log2); k += x.nlanes()) me_scan_block(w, k, n, !counts_only);
    x.sync();
    if (w.step_last >= 0 && x.lane() == 0) w.last_sctr = w.step_last & 255;
    HLB_LAP(w, 5);
    if (!counts_only) {
        me_cost_all(x, w, f, n, px, py);
        x.sync();
    }
    HLB_LAP(w, 6);
}

HLB_HD void set_best(MbWork& w, int p, int s, double cost, int c)
{
    w.best_cost[p][s] = cost; w.best_sctr[p][s] = w.c_sctr[c]; w.best_dist[p][s] = w.c_dist[c]; w.best_cbp[p][s] = w.c_cbp[c];
    w.best_mv[p][s][0] = (int16_t)mv_x(w.cand_mv[c]); w.best_mv[p][s][1] = (int16_t)mv_y(w.cand_mv[c]);
}

// Start of the search of one mode (me_ds.c:104-228): resets the per-partition bests and, for the 16x16 mode on reference 0, runs the PSkip probe
// (me_ds.c:229-261).  Returns 1 when the probe passed: the partition's best is then the predictor at cost 0, which no candidate can beat.
template <class X>
HLB_FN int me_search_begin(X& x, MbWork& w, const FrameCtx& f, int mode)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (x.lane() == 0) w.mode = mode;
#pragma unroll 1
    for (int i = x.lane(); i < 16; i += x.nlanes()) { w.best_sctr[i >> 2][i & 3] = 9; w.best_dist[i >> 2][i & 3] = INT_MAX; w.best_cost[i >> 2][i & 3] = DBL_MAX; }
    int probably_pskip = 0;
    x.sync();
    if (mode == 0 && w.ref == 0) {  // PSkip probe (me_ds.c:229-261)
        int sx, sy, px, py;
        derive_mvp(w, f, 0, 0, 0, px, py);   // w.ref == 0: the predictor the P_Skip derivation would compute itself
        if (pskip_mv_is_zero(w, f)) sx = sy = 0; else { sx = px; sy = py; }
        if (px == sx && py == sy) {
            x.sync();
            if (x.lane() == 0) { set_part(w, mode, 0, 0); w.cand_mv[0] = mv_pack(px, py); }
            x.sync();
            me_step(x, w, f, 1, px, py, false);
            if (w.c_rbc[0] == 0 || w.c_sctr[0] < 6) {
                probably_pskip = 1;
                x.sync();
                if (x.lane() == 0) set_best(w, 0, 0, 0.0, 0);
            }
        }
    }
    x.sync();
    if (x.lane() == 0) w.probably_pskip = probably_pskip;
    x.sync();
    return probably_pskip;
}

// The partition searches of the mode (me_ds.c:262-477)
template <class X>
HLB_FN void me_search_parts(X& x, MbWork& w, const FrameCtx& f, int mode)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int nparts = mode_nparts(mode), nsub = mode_nsub(mode);
    int n;
#pragma unroll 1
    for (int p = 0; p < nparts; ++p) {
#pragma unroll 1
        for (int s = 0; s < nsub; ++s) {
            int shift = 2, flags = 0xFFFFFF, px, py;
            derive_mvp(w, f, p, s, w.ref, px, py);
            // best candidate so far, identical in every lane: cost, vector (kept in registers; mirrored in w.best_* by lane 0)
            double bcost = w.best_cost[p][s];
            int bmx = w.best_mv[p][s][0], bmy = w.best_mv[p][s][1];
            x.sync();
            n = (px != 0 || py != 0) ? 2 : 1;
            if (x.lane() == 0) {
                w.mvp[p][s][0] = (int16_t)px; w.mvp[p][s][1] = (int16_t)py;
                set_part(w, mode, p, s);
                // cost at the predictor, then at (0,0) (me_ds.c:283-299)
                w.cand_mv[0] = mv_pack(px, py); w.cand_mv[1] = 0;
            }
            x.sync();
            me_step(x, w, f, n, px, py, bcost == 0.0);
            if (bcost != 0.0) {
#pragma unroll 1
                for (int c = 0; c < n; ++c) {
                    HLB_DBG("    cand m%d p%d s%d mv (%d,%d) dist %d bits %d sctr %d cost %.4f [init]\n", mode, p, s, mv_x(w.cand_mv[c]), mv_y(w.cand_mv[c]), w.c_dist[c], w.c_rbc[c], w.c_sctr[c], w.c_cost[c]);
                    const double cc = w.c_cost[c];
                    if (cc < bcost) { bcost = cc; bmx = mv_x(w.cand_mv[c]); bmy = mv_y(w.cand_mv[c]); if (x.lane() == 0) set_best(w, p, s, cc, c); }
                }
            }
            int cx = bmx >> 2, cy = bmy >> 2;
            int wl = cx - f.me_range, wr = cx + f.me_range, wt = cy - f.me_range, wb = cy + f.me_range;
#pragma unroll 1
            for (int iter = 0;; ++iter) {
                if (iter > 4096) { w.stuck = 1; break; }
                int best_idx = -1;
                const int count = shift == 1 ? 5 : 9;
                // pattern points of this iteration, in table order: kept when not pruned and inside the window (me_ds.c:317-328); one lane per point
                x.sync();   // the previous step's candidates have been read by every lane
                int okmask = 0;
#if defined(__CUDA_ARCH__)
                {
                    const int i = x.lane() < 9 ? x.lane() : 8, mx = cx + kDsp[shift][i][0], my = cy + kDsp[shift][i][1];
                    const bool ok = x.lane() < count && ((flags >> i) & 1) && mx >= wl && mx <= wr && my >= wt && my <= wb;
                    okmask = (int)__ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int at = hlb_popc((uint32_t)okmask & ((1u << i) - 1u));
                        w.cand_mv[at] = mv_pack(mx * (1 << shift), my * (1 << shift));
                        w.cand_pat[at] = (uint8_t)i;
                    }
                }
#else
                for (int i = 0; i < count; ++i) {
                    const int mx = cx + kDsp[shift][i][0], my = cy + kDsp[shift][i][1];
                    if (((flags >> i) & 1) && mx >= wl && mx <= wr && my >= wt && my <= wb) {
                        const int at = hlb_popc((uint32_t)okmask);
                        okmask |= 1 << i;
                        w.cand_mv[at] = mv_pack(mx * (1 << shift), my * (1 << shift));
                        w.cand_pat[at] = (uint8_t)i;
                    }
                }
#endif
                n = hlb_popc((uint32_t)okmask);
                x.sync();
                if (n > 0) {
                    const bool dead = bcost == 0.0;
                    me_step(x, w, f, n, px, py, dead);
                    if (!dead) {
                        // the first strictly smaller cost in candidate order wins (me_ds.c:345) = the smallest cost, the earliest candidate among equals
                        double mc = DBL_MAX;
                        int mi = INT_MAX;
#pragma unroll 1
                        for (int c = x.lane(); c < n; c += x.nlanes()) {
                            HLB_DBG("    cand m%d p%d s%d mv (%d,%d) dist %d bits %d sctr %d cost %.4f\n", mode, p, s, mv_x(w.cand_mv[c]), mv_y(w.cand_mv[c]), w.c_dist[c], w.c_rbc[c], w.c_sctr[c], w.c_cost[c]);
                            const double cc = w.c_cost[c];
                            if (cc < mc) { mc = cc; mi = c; }
                        }
                        x.reduce_argmin(mc, mi);
                        if (mc < bcost) {
                            bcost = mc; bmx = mv_x(w.cand_mv[mi]); bmy = mv_y(w.cand_mv[mi]);
                            best_idx = w.cand_pat[mi];
                            if (x.lane() == 0) set_best(w, p, s, mc, mi);
                        }
                    }
                }
                flags = 0xFFFFFF;
                if (shift == 2 && best_idx == -1) {
                    shift = 1;
                    cx = bmx >> 2; cy = bmy >> 2;  // integer-pel value used as a half-pel centre
                    wl = cx - f.me_range; wr = cx + f.me_range; wt = cy - f.me_range; wb = cy + f.me_range;
                } else if (best_idx == -1) {
                    if (shift == 1) {
                        shift = 0;
                        cx = bmx; cy = bmy;
                        wl = cx - f.me_range; wr = cx + f.me_range; wt = cy - f.me_range; wb = cy + f.me_range;
                    } else break;
                } else {
                    cx = bmx >> shift; cy = bmy >> shift;
                    flags &= ~(int)kPrune[shift][best_idx];
                }
            }
            x.sync();
            if (x.lane() == 0) { w.mv_cur[p][s][0] = (int16_t)bmx; w.mv_cur[p][s][1] = (int16_t)bmy; }
            x.sync();
            HLB_DBG("   part m%d p%d s%d: mv (%d,%d) mvp (%d,%d) cost %.4f dist %d sctr %d\n", mode, p, s, bmx, bmy, px, py, w.best_cost[p][s], w.best_dist[p][s], w.best_sctr[p][s]);
        }
    }
}

template <class X>
HLB_FN void me_find_best_cost(X& x, MbWork& w, const FrameCtx& f, int mode)
{
    me_search_begin(x, w, f, mode);
    me_search_parts(x, w, f, mode);
}

// ---- the 16x16 search after a passed PSkip probe, when the macroblock is already known to end as P_Skip ----
// After the probe the partition's best cost is 0, so the rest of the search (predictor + zero vector, one integer, one half-pel and one quarter-pel pattern
// around the predictor, me_ds.c:283-477) evaluates a FIXED list of <= 26 candidates and changes nothing but (a) TotalCoeffsLuma[] of the macroblock and (b) the
// Single_ctr chain (the value the LAST non-zero trial block leaves, residual.c:882).  (a) is dead state for a P_Skip macroblock: neighbours count 0 coefficients
// for a skipped macroblock whatever it holds (residual.c:712,733), and the next picture gates every in-macroblock count by this macroblock's CodedBlockPatternLuma,
// which a skip sets to 0 (utils.h:10-20; a later coded macroblock at this address rewrites the counts of every 8x8 it codes).  So only (b) is computed: the
// candidates are walked from the last one backwards until a trial block with non-zero levels is found.  The work counters still count the whole list.
template <class X>
HLB_FN void me_pskip_tail(X& x, MbWork& w, const FrameCtx& f)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int px = w.best_mv[0][0][0], py = w.best_mv[0][0][1], R = f.me_range;
    x.sync();
    if (x.lane() == 0) {
        w.mvp[0][0][0] = (int16_t)px; w.mvp[0][0][1] = (int16_t)py;
        w.mv_cur[0][0][0] = (int16_t)px; w.mv_cur[0][0][1] = (int16_t)py;
        set_part(w, 0, 0, 0);
    }
    // The candidates the full search would visit, in its order: the predictor, (0,0) when that is another vector, then the integer, half and quarter
    // patterns -- centres: integer and half-pel stage around (mvp >> 2), quarter-pel stage around mvp; window = centre +- me_range.  One slot per
    // lane (25 slots), compacted in slot order.
    int n = 0, iops = 0;
    {
#if defined(__CUDA_ARCH__)
        const int j0 = x.lane(), j1 = j0 + 1;
#else
        const int j0 = 0, j1 = 25;
#endif
        uint32_t okmask = 0, mymv = 0;
        bool myok = false;
        for (int j = j0; j < j1 && j < 25; ++j) {
            bool ok;
            uint32_t mv;
            if (j == 0) { ok = true; mv = mv_pack(px, py); }
            else if (j == 1) { ok = px != 0 || py != 0; mv = 0; }
            else {
                const int k = j - 2, shift = k < 9 ? 2 : (k < 14 ? 1 : 0), i = k < 9 ? k : (k < 14 ? k - 9 : k - 14);
                const int cx = shift ? px >> 2 : px, cy = shift ? py >> 2 : py, dx = kDsp[shift][i][0], dy = kDsp[shift][i][1];
                ok = dx >= -R && dx <= R && dy >= -R && dy <= R;
                mv = mv_pack((cx + dx) * (1 << shift), (cy + dy) * (1 << shift));
            }
            if (ok) {
                const int xf = mv_x(mv) & 3, yf = mv_y(mv) & 3;
                iops += (xf == 0 && yf == 0) ? 0 : ((xf == 0 || yf == 0) ? (((xf | yf) == 2) ? 176 : 208) : (((xf & 1) && (yf & 1)) ? 352 : 880));
#if !defined(__CUDA_ARCH__)
                w.tail_mv[n] = mv;
#endif
                ++n;
            }
            myok = ok; mymv = mv;
        }
#if defined(__CUDA_ARCH__)
        okmask = __ballot_sync(0xffffffffu, myok && x.lane() < 25);
        if (myok && x.lane() < 25) w.tail_mv[hlb_popc(okmask & ((1u << x.lane()) - 1u))] = mymv;
        n = hlb_popc(okmask);
        if (x.lane() >= 25) iops = 0;
#else
        (void)okmask; (void)mymv; (void)myok;
#endif
    }
    iops = x.reduce_add(iops);
    if (x.lane() == 0) {
        w.tail_n = n;
        w.stat_trials += (unsigned)(n << 4); w.stat_cands += (unsigned)n; w.stat_interp += (unsigned)(iops << 4);
    }
    x.sync();
#pragma unroll 1
    for (int c = w.tail_n - 1; c >= 0; --c) {
        if (x.lane() == 0) w.cand_mv[0] = w.tail_mv[c];
        x.sync();
        me_step(x, w, f, 1, px, py, true, false);   // sets w.last_sctr when a trial block of the candidate has non-zero levels
        const int found = w.step_last;
        x.sync();
        if (found >= 0) break;
    }
    x.sync();
}

// ------------------------------------------------------------------------------------------------------------------
// Inter prediction of the whole macroblock with the committed geometry (rdo.c:2331-2416) -> w.pred_y / w.pred_c
// lanes 0..15 luma blocks (raster), 16..143 chroma samples (Cb 0..63, Cr 0..63)
// ------------------------------------------------------------------------------------------------------------------
HLB_HD void fin_rect(const MbWork& w, int x, int y, int& part, int& sub, int& ox, int& oy)
{
    uint8_t sm[4] = {(uint8_t)w.fin_sub[0], (uint8_t)w.fin_sub[1], (uint8_t)w.fin_sub[2], (uint8_t)w.fin_sub[3]};
    part_at(w.fin_mode, sm, x, y, part, sub);
    int m = w.fin_mode < 3 ? w.fin_mode : 3 + w.fin_sub[part], pw, ph;
    mode_rect(m, part, sub, ox, oy, pw, ph);
}
HLB_FN void phase_pred_inter(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    // lanes 0..31: chroma (one round of a warp, its two dependent reference fetches per lane in flight together with everyone else's); lanes 32..47: luma
    if (lane >= 32 && lane < 48) {
        lane -= 32;
        const int bx = (lane & 3) * 4, by = (lane >> 2) * 4;
        int p, s, ox, oy;
        fin_rect(w, bx, by, p, s, ox, oy);
        // straight out of the search's reference tile when it covers the block's window (it does for every vector the search ended on or near), with the packed
        // formulation of the trials; otherwise from the picture with the reference's clamp
        const int mvx = w.fin_mv[p][s][0], mvy = w.fin_mv[p][s][1];
        const int tx = clip3(-17, f.W + 17, w.mbx * 16 + ox + (mvx >> 2)) + (bx - ox) - w.tile_x0, ty = clip3(-17, f.H + 17, w.mby * 16 + oy + (mvy >> 2)) + (by - oy) - w.tile_y0;
        Rows4 o;
        if (w.tile_valid && w.tile_ref == w.fin_ref[p] && tx >= 2 && ty >= 2 && tx + 7 <= HLB_TILE_W && ty + 7 <= HLB_TILE_H)
            o = fast_pred_luma((const uint32_t*)w.tile, HLB_TILE_W / 4, tx, ty, mvx & 3, mvy & 3);
        else {
            uint8_t pv[16], win[84];
            pred_luma_4x4(f, f.ref[w.fin_ref[p]][0], w.mbx, w.mby, ox, oy, bx, by, mvx, mvy, win, pv);
#pragma unroll
            for (int r = 0; r < 4; ++r) o.r[r] = (uint32_t)pv[r * 4] | ((uint32_t)pv[r * 4 + 1] << 8) | ((uint32_t)pv[r * 4 + 2] << 16) | ((uint32_t)pv[r * 4 + 3] << 24);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) ((uint32_t*)w.pred_y)[((by + r) * 16 + bx) >> 2] = o.r[r];
    } else if (lane < 32) {   // four chroma samples per lane (plane, row, half row): two sample pairs, each with the vector of its own (sub-)partition
        const int g = lane, c = g >> 4, py = (g >> 1) & 7, px = (g & 1) * 4;
        const int Wc = f.W >> 1, Hc = f.H >> 1;
        uint8_t sm[4] = {(uint8_t)w.fin_sub[0], (uint8_t)w.fin_sub[1], (uint8_t)w.fin_sub[2], (uint8_t)w.fin_sub[3]};
        uint32_t out = 0;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {   // rolled: this runs once per macroblock, its footprint in the instruction cache counts more than its instructions
            int p, s;
            part_at(w.fin_mode, sm, (px + 2 * h) * 2, py * 2, p, s);
            const int mvx = w.fin_mv[p][s][0], mvy = w.fin_mv[p][s][1];
            out |= fast_chroma_two(f.ref[w.fin_ref[p]][1 + c], Wc, Hc, w.mbx * 8 + px + 2 * h + (mvx >> 3), w.mby * 8 + py + (mvy >> 3), mvx & 7, mvy & 7) << (16 * h);
        }
        ((uint32_t*)w.pred_c[c])[(py * 8 + px) >> 2] = out;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Luma residual coding + reconstruction of an inter macroblock (rdo.c:2418-2478): lanes 0..15 = luma4x4BlkIdx
// ------------------------------------------------------------------------------------------------------------------
HLB_FN void phase_recon_luma(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane >= 16) return;
    const int bx = blk_x(lane), by = blk_y(lane);
    Rows4 sv, pv;
    bool nz = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        sv.r[r] = ((const uint32_t*)w.src_y)[((by + r) * 16 + bx) >> 2]; pv.r[r] = ((const uint32_t*)w.pred_y)[((by + r) * 16 + bx) >> 2];
        nz |= sv.r[r] != pv.r[r];
    }
    Rows4 rec = pv;
    bool coded = false;
    if (!w.luma_skip_residual) {
        uint32_t* lvw = (uint32_t*)w.luma_level[lane];
        if (nz) {
            int m[16];
            fast_fwd_transform(sv, pv, m);
            fast_quant(m, f.qk);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) any |= (uint32_t)m[i];
            coded = any != 0;
            if (coded) {   // levels in zig-zag order (0 1 4 8 5 2 3 6 9 12 13 10 7 11 14 15), two per word
                lvw[0] = pack16(m[0], m[1]); lvw[1] = pack16(m[4], m[8]); lvw[2] = pack16(m[5], m[2]); lvw[3] = pack16(m[3], m[6]);
                lvw[4] = pack16(m[9], m[12]); lvw[5] = pack16(m[13], m[10]); lvw[6] = pack16(m[7], m[11]); lvw[7] = pack16(m[14], m[15]);
                fast_dequant_inverse(m, f.qk, false);
                rec = fast_recon_clip(pv, m);
            }
        }
        if (!coded) {
#pragma unroll
            for (int i = 0; i < 8; ++i) lvw[i] = 0;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) ((uint32_t*)w.rec_y)[((by + r) * 16 + bx) >> 2] = rec.r[r];
    w.blk_coded[lane] = coded ? 1 : 0;   // gathered into cbp_luma4x4 by the master
}

// ------------------------------------------------------------------------------------------------------------------
// Chroma residual coding + reconstruction (rdo.c:2502-2700, transf.c:161-296) from w.pred_c.
// phase 0: lanes 0..7 = (plane, block): residual, transform, AC quantisation (intra offset, always), CAVLC info
// phase 1: lane 0: the serial bookkeeping (single-coefficient elimination, DC Hadamard + quantisation)
// phase 2: lanes 0..7: reconstruction
// ------------------------------------------------------------------------------------------------------------------
HLB_FN void phase_chroma_tq(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane >= 8) return;
    const int c = lane >> 2, b = lane & 3, o4 = ((b >> 1) * 32 + (b & 1) * 4) >> 2;   // word offset of the block in the 8x8 plane
    Rows4 sv, pv;
    bool nz = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        sv.r[r] = ((const uint32_t*)w.src_c[c])[o4 + 2 * r]; pv.r[r] = ((const uint32_t*)w.pred_c[c])[o4 + 2 * r];
        nz |= sv.r[r] != pv.r[r];
    }
    w.c_resnz[c][b] = nz ? 1 : 0;
    w.c_dccoef[c][b] = 0; w.c_acnz[c][b] = 0; w.c_sc[c][b] = 9; w.c_tc[c][b] = 0;
    if (!nz) return;
    uint32_t* acw = (uint32_t*)w.chroma_ac[c][b];   // ChromaACLevel[0..14] = zig-zag positions 1..15; element 15 is never written (kept from earlier pictures)
    int m[16];
    fast_fwd_transform(sv, pv, m);
    w.c_dccoef[c][b] = m[0];
    fast_quant(m, f.qkc);
    // zig-zag 0 1 4 8 5 2 3 6 9 12 13 10 7 11 14 15: AC list = positions 1..15
    int l16[16] = {m[1], m[4], m[8], m[5], m[2], m[3], m[6], m[9], m[12], m[13], m[10], m[7], m[11], m[14], m[15], 0};
#pragma unroll
    for (int i = 0; i < 7; ++i) acw[i] = pack16(l16[2 * i], l16[2 * i + 1]);
    w.chroma_ac[c][b][14] = (int16_t)l16[14];
    const uint32_t mask = level_mask16(l16);
    w.c_acnz[c][b] = mask ? 1 : 0;
    if (mask) {
        const CavlcInfo ci = cavlc_block_info16(l16, mask);
        w.c_sc[c][b] = ci.single_ctr; w.c_tc[c][b] = ci.total_coeff;
    }
}
HLB_FN void phase_chroma_serial(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane != 0) return;
    // one lane, once per macroblock: loops stay rolled and the common case (no AC level anywhere) skips the elimination walk -- the footprint in the instruction
    // cache and the executed instructions matter more than the trip counts.  c_acnz / c_dccoef are only ever non-zero for blocks with a residual.
    int cbp_ac[2] = {0, 0};
    if ((*(const uint32_t*)w.c_acnz[0] | *(const uint32_t*)w.c_acnz[1]) != 0) {
        int single[2] = {0, 0}, totc[2] = {0, 0};
#pragma unroll 1
        for (int b = 0; b < 4; ++b)
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {   // block-major, plane-minor: the order the Single_ctr chain is written in (rdo.c:2599-2625)
                if (w.c_acnz[c][b]) cbp_ac[c] |= 1 << b;
                if (single[c] < 7 && ((cbp_ac[c] >> b) & 1)) {
                    single[c] += w.c_sc[c][b]; totc[c] += w.c_tc[c][b];
                    w.tc_cac[c][b] = w.c_tc[c][b];
                    w.last_sctr = w.c_sc[c][b];
                }
            }
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (single[c] < 7 && totc[c] == 1) cbp_ac[c] = 0;
    }
    w.cbp_ac[0] = cbp_ac[0]; w.cbp_ac[1] = cbp_ac[1];
    // 2x2 DC: Hadamard, quantisation with the macroblock's own rounding offset (quant.c:150-189: qbits + 1, 2f, MF(0,0)), de-quantisation (transf.c:612:
    // ((f * LevelScale(QPC % 6, 0, 0)) << (QPC / 6)) >> 5, with QuantK::dq_mul[0] = LevelScale << (QPC / 6 - 4) once QPC >= 24)
    const int qb1 = f.qkc.qbits + 1, f2 = ((1 << f.qkc.qbits) / (w.mb_is_intra ? 3 : 6)) << 1, mf = f.qkc.mf[0], q6 = f.qkc.qbits - 15;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        int h[4] = {w.c_dccoef[c][0], w.c_dccoef[c][1], w.c_dccoef[c][2], w.c_dccoef[c][3]};
        int d[4] = {0, 0, 0, 0}, mask = 0;
        if ((h[0] | h[1] | h[2] | h[3]) != 0) {
            hadamard2x2(h);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int z = (iabs(h[k]) * mf + f2) >> qb1;
                h[k] = h[k] >= 0 ? z : -z;
                w.chroma_dc[c][k] = (int16_t)h[k]; mask |= (h[k] != 0) << k;
            }
            if (mask) {
                hadamard2x2(h);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const int v = h[k] * f.qkc.dq_mul[0]; d[k] = q6 >= 4 ? v >> 1 : (v << q6) >> 5; }
            }
        }
        w.cbp_dc[c] = mask;
        // de-quantised DC per block, parked in c_dccoef for the reconstruction lanes
#pragma unroll
        for (int k = 0; k < 4; ++k) w.c_dccoef[c][k] = d[k];
    }
}
HLB_FN void phase_chroma_recon(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane >= 8) return;
    const int c = lane >> 2, b = lane & 3, o4 = ((b >> 1) * 32 + (b & 1) * 4) >> 2;
    const int dc = w.c_dccoef[c][b];
    const bool use = (w.cbp_dc[c] || w.cbp_ac[c]) && (dc != 0 || ((w.cbp_ac[c] >> b) & 1));
    Rows4 pv;
#pragma unroll
    for (int r = 0; r < 4; ++r) pv.r[r] = ((const uint32_t*)w.pred_c[c])[o4 + 2 * r];
    Rows4 rec = pv;
    if (use) {
        const int16_t* e = w.chroma_ac[c][b];   // AC levels in zig-zag order back to raster positions
        int m[16] = {dc, e[0], e[4], e[5], e[1], e[3], e[6], e[11], e[2], e[7], e[10], e[12], e[8], e[9], e[13], e[14]};
        fast_dequant_inverse(m, f.qkc, /*keep_dc*/ true);
        rec = fast_recon_clip(pv, m);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) ((uint32_t*)w.rec_c[c])[o4 + 2 * r] = rec.r[r];
}

// ------------------------------------------------------------------------------------------------------------------
// load / store of the macroblock's samples: 96 lanes, one aligned 32-bit word each (64 luma words = 16 rows x 4, then 2 x 16 chroma words);
// macroblock columns are 16-sample aligned and the plane pitch is a multiple of 16 (8 for chroma), the plane bases 4-byte aligned
// ------------------------------------------------------------------------------------------------------------------
HLB_FN void phase_load(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane < 64) {
        const int r = lane >> 2, q = lane & 3;
        ((uint32_t*)w.src_y)[lane] = HLB_LDG((const uint32_t*)(f.src[0] + (w.mby * 16 + r) * f.W + w.mbx * 16) + q);
    } else if (lane < 96) {
        const int i = lane - 64, c = i >> 4, r = (i >> 1) & 7, q = i & 1, Wc = f.W >> 1;
        ((uint32_t*)w.src_c[c])[i & 15] = HLB_LDG((const uint32_t*)(f.src[1 + c] + (w.mby * 8 + r) * Wc + w.mbx * 8) + q);
    }
}
// arg0 bit 0: luma, bit 1: chroma, bit 2: the luma reconstruction IS the prediction (P_Skip: no separate copy into rec_y)
HLB_FN void phase_store(MbWork& w, const FrameCtx& f, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    if (lane < 64 && (w.arg0 & 1)) {
        const int r = lane >> 2, q = lane & 3;
        ((uint32_t*)(f.cur[0] + (w.mby * 16 + r) * f.W + w.mbx * 16))[q] = (w.arg0 & 4) ? ((const uint32_t*)w.pred_y)[lane] : ((const uint32_t*)w.rec_y)[lane];
    } else if (lane >= 64 && lane < 96 && (w.arg0 & 2)) {
        const int i = lane - 64, c = i >> 4, r = (i >> 1) & 7, q = i & 1, Wc = f.W >> 1;
        ((uint32_t*)(f.cur[1 + c] + (w.mby * 8 + r) * Wc + w.mbx * 8))[q] = ((const uint32_t*)w.rec_c[c])[i & 15];
    }
}

}  // namespace hlb

#include "hlb_mbintra.cuh"

namespace hlb {

// number of phases and the phase dispatcher (the only place that knows which function a command runs)
HLB_HD int cmd_phases(int cmd)
{
    switch (cmd) {
    case CMD_ME_EVAL: return 1;
    case CMD_TILE_FIX: return 2;
    case CMD_CHROMA: return 3;
    case CMD_I16_EVAL: return 2;
    case CMD_I4_EVAL: return 2;
    default: return 1;
    }
}
HLB_FN void cmd_phase(MbWork& w, const FrameCtx& f, int cmd, int phase, int lane)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    switch (cmd) {
    case CMD_LOAD: phase_load(w, f, lane); break;
    case CMD_TILE: phase_tile_load(w, f, lane); break;
    case CMD_TILE_TMA: phase_tile_tma(w, f, lane); break;
    case CMD_TILE_FIX: phase_tile_fix(w, f, phase, lane); break;
    case CMD_ME_EVAL: me_phase_trial(w, f, lane); break;
    case CMD_PRED_INTER: phase_pred_inter(w, f, lane); break;
    case CMD_RECON_LUMA: phase_recon_luma(w, f, lane); break;
    case CMD_CHROMA:
        if (phase == 0) phase_chroma_tq(w, f, lane);
        else if (phase == 1) phase_chroma_serial(w, f, lane);
        else phase_chroma_recon(w, f, lane);
        break;
    case CMD_STORE: phase_store(w, f, lane); break;
    case CMD_I16_EVAL: i16_phase(w, f, phase, lane); break;
    case CMD_I16_RATE: i16_phase(w, f, 2, lane); break;
    case CMD_I16_RECON: i16_recon_phase(w, f, lane); break;
    case CMD_I4_EVAL: i4_phase(w, f, phase, lane); break;
    case CMD_I4_COMMIT: i4_commit_phase(w, f, lane); break;
    case CMD_PRED_CHROMA_INTRA: intra_chroma_pred_phase(w, f, lane); break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The macroblock: rdo.c:678-1270 (P) / rdo.c:99 (I)
// ------------------------------------------------------------------------------------------------------------------
template <class X>
HLB_HD void chroma_code(X& x, MbWork& w) { x.run(CMD_CHROMA, 8); }

HLB_FN void mb_begin(MbWork& w, const FrameCtx& f, int mb, int lane, int nl)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    const int mbx = mb % f.mbw, mby = mb / f.mbw;
    const int availA = mbx > 0, availB = mby > 0, availC = mby > 0 && mbx < f.mbw - 1, availD = mbx > 0 && mby > 0;
    w.mb = mb; w.mbx = mbx; w.mby = mby;
    w.availA = availA; w.availB = availB; w.availC = availC; w.availD = availD;
    const MbState& s = f.st[mb];
    // heads of the own and the four neighbour states: one coalesced copy
#pragma unroll 1
    for (int i = lane; i < 5 * HLB_NB_WORDS; i += nl) {
        const int k = i / HLB_NB_WORDS, j = i - k * HLB_NB_WORDS;
        const int addr = k == 0 ? mb : (k == 1 ? mb - 1 : (k == 2 ? mb - f.mbw : (k == 3 ? mb - f.mbw + 1 : mb - f.mbw - 1)));
        const bool ok = k == 0 || (k == 1 ? availA : (k == 2 ? availB : (k == 3 ? availC : availD)));
        w.nbw[k][j] = ok ? HLB_LDCG((const uint32_t*)&f.st[addr] + j) : 0u;   // neighbours' state was written by other CTAs of this launch
    }
#pragma unroll 1
    for (int i = lane; i < 64; i += nl) ((uint32_t*)&w.chroma_ac[0][0][0])[i] = HLB_LDCG((const uint32_t*)&s.chroma_ac[0][0][0] + i);
#pragma unroll 1
    for (int i = lane; i < 4; i += nl) ((uint32_t*)&w.chroma_dc[0][0])[i] = HLB_LDCG((const uint32_t*)&s.chroma_dc[0][0] + i);
    HLB_LANE_SYNC();
    const MbState& o = *(const MbState*)w.nbw[0];
    const MbState& sA = *(const MbState*)w.nbw[1];
    const MbState& sB = *(const MbState*)w.nbw[2];
#pragma unroll 1
    for (int i = lane; i < 16; i += nl) {
        w.tc[i] = o.tc_luma[i]; w.i4_mode[i] = o.i4_mode[i];
        ((uint32_t*)&w.mv_cur[0][0][0])[i] = ((const uint32_t*)&o.mv[0][0][0])[i];
        const int bx = blk_x(i), by = blk_y(i);
        w.extA[i] = (int8_t)((bx == 0) ? (availA ? nb_count(sA, blk_idx_from_xy(12, by)) : -1) : 0);
        w.extB[i] = (int8_t)((by == 0) ? (availB ? nb_count(sB, blk_idx_from_xy(bx, 12)) : -1) : 0);
    }
#pragma unroll 1
    for (int i = lane; i < 8; i += nl) w.tc_cac[i >> 2][i & 3] = o.tc_cac[i >> 2][i & 3];
#pragma unroll 1
    for (int i = lane; i < 4; i += nl) w.ref_cur[i] = o.ref_idx[i];
    w.cbp_gate = o.cbp_luma;
    w.last_sctr = -1; w.need_prev_sctr = 0; w.stuck = 0; w.tile_valid = 0;
    w.stat_trials = w.stat_interp = w.stat_cands = w.stat_intra = 0;
    w.prof_run_cycles = w.prof_runs = w.prof_me_cycles = 0;
#if defined(HLB_PROFILE_STEPS) && defined(__CUDA_ARCH__)
    for (int i = 0; i < 16; ++i) { w.prof_lap[i] = 0; w.prof_cnt[i] = 0; }
    w.prof_last = clock64();
#endif
    w.mb_is_intra = 0;
}

// state the writer leaves behind (residual.c:903-1094) + publication of the macroblock's final state and record
HLB_HD int nnz16(const int16_t* lv, int n)
{
    int k = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) k += (lv[i] != 0);
    return k;
}

HLB_FN void mb_commit(MbWork& w, const FrameCtx& f, int kind, int cbp_luma, int cbp_chroma, int coded_block_pattern, int mb_type, const int16_t mvd[4][4][2], int mad, int lane, int nl)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    MbState& s = f.st[w.mb];
    hlb200_mb_record_t& r = f.rec[w.mb];
    // ---- record (what the host writer consumes) ----
    if (lane == 0) {
        r.mb_class = (uint8_t)kind; r.mb_type = (uint8_t)mb_type;
        r.part_mode = (uint8_t)w.fin_mode;
        r.i16_pred_mode = (uint8_t)w.i16_mode; r.intra_chroma_pred_mode = (uint8_t)w.intra_chroma_mode;
        r.coded_block_pattern = (uint8_t)coded_block_pattern; r.cbp_luma = (uint8_t)cbp_luma; r.cbp_chroma = (uint8_t)cbp_chroma;
        r.cbp_luma4x4 = (uint16_t)w.cbp_luma4x4;
        r.mb_qp_delta = 0; r.qp_y = (uint8_t)f.qp; r.qp_c[0] = r.qp_c[1] = (uint8_t)f.qpc;
        r.mad = mad;
        r.me_trials = w.stat_trials; r.me_interp_ops = w.stat_interp; r.me_candidates = (uint16_t)w.stat_cands; r.intra_trials = (uint16_t)w.stat_intra;
    }
#pragma unroll 1
    for (int i = lane; i < 4; i += nl) { r.sub_mode[i] = (uint8_t)w.fin_sub[i]; r.ref_idx[i] = w.fin_ref[i]; }
#pragma unroll 1
    for (int c = lane; c < 2; c += nl) { r.cbp_chroma_dc4x4[c] = (uint8_t)w.cbp_dc[c]; r.cbp_chroma_ac4x4[c] = (uint8_t)w.cbp_ac[c]; }
#pragma unroll 1
    for (int i = lane; i < 16; i += nl) { r.i4_pred_mode[i] = w.i4_mode[i]; r.prev_intra4x4_pred_mode_flag[i] = w.prev_i4[i]; r.rem_intra4x4_pred_mode[i] = w.rem_i4[i]; r.i16_dc_level[i] = w.i16_dc[i]; }
#pragma unroll 1
    for (int i = lane; i < 32; i += nl) {
        const int p = i >> 3, q = (i >> 1) & 3, k = i & 1;
        r.mv[p][q][k] = w.fin_mv[p][q][k]; r.mvd[p][q][k] = mvd ? mvd[p][q][k] : (int16_t)0;
    }
    // level arrays: two int16 per 32-bit store, lanes strided
    {
        uint32_t* d0 = (uint32_t*)&r.luma_level[0][0]; const uint32_t* s0 = (const uint32_t*)&w.luma_level[0][0];
        uint32_t* d1 = (uint32_t*)&r.i16_ac_level[0][0]; const uint32_t* s1 = (const uint32_t*)&w.i16_ac[0][0];
        uint32_t* d2 = (uint32_t*)&r.chroma_ac_level[0][0][0]; const uint32_t* s2 = (const uint32_t*)&w.chroma_ac[0][0][0];
        const bool intra16 = kind == MBK_I16;
#pragma unroll 1
        for (int i = lane; i < 128; i += nl) { d0[i] = intra16 ? 0u : s0[i]; d1[i] = intra16 ? s1[i] : 0u; }
#pragma unroll 1
        for (int i = lane; i < 64; i += nl) d2[i] = s2[i];
    }
#pragma unroll 1
    for (int i = lane; i < 8; i += nl) r.chroma_dc_level[i >> 2][i & 3] = w.chroma_dc[i >> 2][i & 3];
    // ---- TotalCoeffs as the writer finalises them ----
    HLB_LANE_SYNC();
    if (kind != MBK_PSKIP) {
#pragma unroll 1
        for (int b = lane; b < 16; b += nl) {
            int v = -1;
            if ((cbp_luma >> (b >> 2)) & 1) v = kind == MBK_I16 ? nnz16(w.i16_ac[b], 15) : nnz16(w.luma_level[b], 16);
            else if (kind == MBK_I16 && b == 0) v = nnz16(w.i16_dc, 16);
            if (v >= 0) w.tc[b] = (uint8_t)v;
        }
        if (cbp_chroma & 2)
#pragma unroll 1
            for (int i = lane; i < 8; i += nl) { const int c = i >> 2, b = i & 3; w.tc_cac[c][b] = (uint8_t)(((w.cbp_ac[c] >> b) & 1) ? nnz16(w.chroma_ac[c][b], 15) : 0); }
    }
    HLB_LANE_SYNC();
    // ---- persistent state ----
    if (lane == 0) {
        s.kind = (uint8_t)kind;
        s.part_mode = (uint8_t)w.fin_mode;
        s.cbp_luma = (uint8_t)cbp_luma; s.cbp_chroma = (uint8_t)cbp_chroma;
        s.last_sctr = (uint8_t)(w.last_sctr < 0 ? 255 : w.last_sctr);
    }
#pragma unroll 1
    for (int i = lane; i < 4; i += nl) { s.sub_mode[i] = (uint8_t)w.fin_sub[i]; s.ref_idx[i] = w.ref_cur[i]; }
#pragma unroll 1
    for (int i = lane; i < 16; i += nl) { s.tc_luma[i] = w.tc[i]; s.i4_mode[i] = w.i4_mode[i]; }
#pragma unroll 1
    for (int i = lane; i < 8; i += nl) { s.tc_cac[i >> 2][i & 3] = w.tc_cac[i >> 2][i & 3]; s.chroma_dc[i >> 2][i & 3] = w.chroma_dc[i >> 2][i & 3]; }
#pragma unroll 1
    for (int i = lane; i < 64; i += nl) ((uint32_t*)&s.chroma_ac[0][0][0])[i] = ((const uint32_t*)&w.chroma_ac[0][0][0])[i];
#pragma unroll 1
    for (int i = lane; i < 16; i += nl) ((uint32_t*)&s.mv[0][0][0])[i] = ((const uint32_t*)&w.mv_cur[0][0][0])[i];
}

HLB_HD int guess_cbp_luma(int cbp4x4, bool i16)
{
    if (i16) return cbp4x4 ? 15 : 0;
    int c = 0;
    for (int b8 = 0; b8 < 4; ++b8)
        if ((cbp4x4 >> (b8 * 4)) & 15) c |= 1 << b8;
    return c;
}
HLB_HD int guess_cbp_chroma(const MbWork& w)
{
    if ((w.cbp_dc[0] || w.cbp_dc[1]) && !w.cbp_ac[0] && !w.cbp_ac[1]) return 1;
    if (w.cbp_ac[0] || w.cbp_ac[1]) return 2;
    return 0;
}

// ---- JVT-O079 2.1.3.4 homogeneous block detection (rdo.c:889-935): which of the 7 search modes stay enabled for this macroblock ----
// The edge map of hl_math_homogeneousity8x8_u8 (hl_math.c:470-486): sum over an 8x8 block of |dx| + |dy| with the Sobel pair on the SOURCE picture.  The
// reference starts the 16x16 area one sample inside the picture when the macroblock lies on the left / top edge and one sample to the left / above when it lies
// on the right / bottom edge (rdo.c:894-895), so the 3x3 support never leaves the plane.  Bit m of the result = mode m of mode_rect() (reference Mode m + 1).
template <class X>
HLB_FN int me_mode_mask(X& x, const MbWork& w, const FrameCtx& f)
{
    const int xl = w.mbx * 16, yl = w.mby * 16;
    const int x0 = xl == 0 ? 1 : (xl == f.W - 16 ? f.W - 17 : xl), y0 = yl == 0 ? 1 : (yl == f.H - 16 ? f.H - 17 : yl);
    int hsum[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int p = x.lane(); p < 256; p += x.nlanes()) {
        const int b = p >> 6, j = (p >> 3) & 7, i = p & 7;
        const uint8_t* c = f.src[0] + (size_t)(y0 + (b >> 1) * 8 + j) * f.W + (x0 + (b & 1) * 8 + i);
        const uint8_t *up = c - f.W, *dn = c + f.W;
        const int ul = HLB_LDG(up - 1), um = HLB_LDG(up), ur = HLB_LDG(up + 1), ml = HLB_LDG(c - 1), mr = HLB_LDG(c + 1), dl = HLB_LDG(dn - 1), dm = HLB_LDG(dn), dr = HLB_LDG(dn + 1);
        const int dx = dl + 2 * dm + dr - ul - 2 * um - ur, dy = ur + 2 * mr + dr - ul - 2 * ml - dl;
        const int v = iabs(dx) + iabs(dy);
#pragma unroll
        for (int k = 0; k < 4; ++k) hsum[k] += b == k ? v : 0;
    }
    int h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = x.reduce_add(hsum[k]);
    HLB_DBG("  homogeneity %d %d %d %d\n", h[0], h[1], h[2], h[3]);
    const int TH16 = 20000, TH8 = 5000, TH84 = 7500;   // HL_CODEC_264_RDO_HOMOGENEOUSITY_TH16X16 / TH8X8 / TH8X4 (hl_codec_264_defs.h:61-63)
    if (h[0] < TH8 && h[1] < TH8 && h[2] < TH8 && h[3] < TH8) return 0x01;
    if (h[0] + h[1] + h[2] + h[3] < TH16) return (h[0] < TH8 && h[1] < TH8) ? 0x03 : 0x05;
    if (h[0] < TH84 && h[1] < TH84 && h[2] < TH84 && h[3] < TH84) return 0x3f;
    return 0x7f;
}

template <class X>
HLB_FN void mb_encode_p(X& x, MbWork& w, const FrameCtx& f)
{
    HLB_IN_SHARED(w); HLB_IN_SHARED(f);
    double best_cost = DBL_MAX;
    int best_dist = 0, best_mode = -1, best_sctr = 9, best_ref = 0, found = 0, b_pskip = 0, probably_pskip = 0, pskip_early = 0;
    int16_t best_mv[4][4][2], best_mvp[4][4][2];
    // the mode flags are reset and recomputed per reference picture (rdo.c:875, :889) from the source alone: once is the same
    const int mode_mask = f.early_term ? me_mode_mask(x, w, f) : 0x7f;
#pragma unroll 1
    for (int u = 0; u < f.num_refs; ++u) {
        if (!f.ref[u][0]) continue;
        w.ref = u; w.ref_y = f.ref[u][0];
        memo_reset(w, x.lane(), x.nlanes());   // remembered trials belong to one reference picture
        x.sync();
#pragma unroll 1
        for (int g = 0; g < 4 && !found; ++g) {
            const int m0 = g < 3 ? g : 3, m1 = g < 3 ? g : 6;
            pskip_early = 0;
#pragma unroll 1
            for (int mode = m0; mode <= m1; ++mode) {
                if (!((mode_mask >> mode) & 1)) continue;   // rdo.c:884
                HLB_LAP(w, 8);
                if (me_search_begin(x, w, f, mode)) {
                    // The PSkip probe passed (16x16, reference 0): the macroblock is a P_Skip iff its chroma also quantises to nothing (rdo.c:1125-1139, :2140).  The
                    // reference checks that after the rest of the 16x16 search; the two only meet in the Single_ctr chain, so the check runs first and the search
                    // shrinks to its only observable effect when the answer is yes (me_pskip_tail).
                    const int sctr_before = w.last_sctr;
                    w.fin_mode = 0; w.fin_sub[0] = w.fin_sub[1] = w.fin_sub[2] = w.fin_sub[3] = 0;
                    w.fin_ref[0] = 0; w.fin_mv[0][0][0] = w.best_mv[0][0][0]; w.fin_mv[0][0][1] = w.best_mv[0][0][1];
                    w.mb_is_intra = 0;
                    x.sync();
                    if (x.lane() == 0) w.last_sctr = -2;   // "not written by the chroma pass"
                    x.sync();
                    x.run(CMD_PRED_INTER, 48);
                    chroma_code(x, w);
                    const int chroma_zero = !w.cbp_ac[0] && !w.cbp_ac[1] && !w.cbp_dc[0] && !w.cbp_dc[1], sctr_chroma = w.last_sctr;
                    x.sync();
                    if (x.lane() == 0) w.last_sctr = sctr_before;
                    x.sync();
                    HLB_LAP(w, 9);
                    if (chroma_zero) { me_pskip_tail(x, w, f); pskip_early = 1; }
                    else { me_search_parts(x, w, f, mode); pskip_early = 2; }
                    if (sctr_chroma != -2) {   // the chroma pass comes last in the reference's order
                        x.sync();
                        if (x.lane() == 0) w.last_sctr = sctr_chroma;
                        x.sync();
                    }
                } else me_search_parts(x, w, f, mode);
                HLB_LAP(w, 7);
                double cost_sum = 0;
                int dist_sum = 0, sctr_sum = 0;
                probably_pskip = w.probably_pskip;
                const int np = mode_nparts(mode), ns = mode_nsub(mode);
#pragma unroll 1
                for (int p = 0; p < np; ++p)
#pragma unroll 1
                    for (int s = 0; s < ns; ++s) { cost_sum += w.best_cost[p][s]; dist_sum += w.best_dist[p][s]; sctr_sum += w.best_sctr[p][s]; }
                if (!probably_pskip && cost_sum != 0 && sctr_sum < 6 && mode == 0) {  // rdo.c:1040-1054
                    int sx, sy;
                    derive_pskip_mv(w, f, sx, sy);
                    probably_pskip = (sx == w.mvp[0][0][0] && sy == w.mvp[0][0][1]) && (w.mv_cur[0][0][0] == w.mvp[0][0][0] && w.mv_cur[0][0][1] == w.mvp[0][0][1]);
                }
                cost_sum += f.lambda * kHeaderBits[mode];
                HLB_DBG("  mode %d ref %d: cost_sum %.4f dist %d sctr %d pskip %d | p0: mv (%d,%d) mvp (%d,%d) cost %.4f\n", mode, u, cost_sum, dist_sum, sctr_sum, probably_pskip,
                        w.mv_cur[0][0][0], w.mv_cur[0][0][1], w.mvp[0][0][0], w.mvp[0][0][1], w.best_cost[0][0]);
                if (cost_sum < best_cost) {
                    best_cost = cost_sum; best_dist = dist_sum; best_sctr = sctr_sum; best_mode = mode; best_ref = u;
#pragma unroll 1
                    for (int p = 0; p < np; ++p)
#pragma unroll 1
                        for (int s = 0; s < ns; ++s)
#pragma unroll 1
                            for (int k = 0; k < 2; ++k) { best_mv[p][s][k] = w.mv_cur[p][s][k]; best_mvp[p][s][k] = w.mvp[p][s][k]; }
                }
            }
            b_pskip = probably_pskip;
            if (pskip_early) b_pskip = pskip_early == 1;   // the chroma check has already run (above)
            else if (b_pskip) {  // chroma must quantise to nothing (rdo.c:1125-1139, :2140)
                w.fin_mode = 0; w.fin_sub[0] = w.fin_sub[1] = w.fin_sub[2] = w.fin_sub[3] = 0;
                w.fin_ref[0] = 0; w.fin_mv[0][0][0] = best_mv[0][0][0]; w.fin_mv[0][0][1] = best_mv[0][0][1];
                w.mb_is_intra = 0;
                HLB_LAP(w, 8);
                x.run(CMD_PRED_INTER, 48);
                chroma_code(x, w);
                b_pskip = !w.cbp_ac[0] && !w.cbp_ac[1] && !w.cbp_dc[0] && !w.cbp_dc[1];
                HLB_LAP(w, 9);
            }
            found |= (best_cost == 0) || b_pskip;
        }
    }
    if (!b_pskip) {
        double intra_cost;
        HLB_LAP(w, 8);
        const int intra_kind = mb_encode_intra(x, w, f, intra_cost);  // reconstructs into the picture and commits when it wins (rdo.c:1161-1167)
        HLB_LAP(w, 10);
        HLB_DBG("  intra cost %.4f (kind %d) vs inter %.4f (mode %d)\n", intra_cost, intra_kind, best_cost, best_mode);
        if (intra_cost <= best_cost) { mb_commit_intra(w, f, intra_kind, x.lane(), x.nlanes()); return; }
    }
    // ---- commit the best inter layout (rdo.c:1170-1218) ----
    HLB_LAP(w, 8);
    w.mb_is_intra = 0;
    w.fin_mode = best_mode < 3 ? best_mode : 3;
    const int fs = best_mode <= 3 ? 0 : best_mode - 3;
    int16_t mvd[4][4][2];
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        w.fin_sub[p] = fs; w.ref_cur[p] = 0; w.fin_ref[p] = 0;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) mvd[p][s][0] = mvd[p][s][1] = 0;
    }
    const int np = mode_nparts(best_mode), ns = mode_nsub(best_mode);
#pragma unroll 1
    for (int p = 0; p < np; ++p) {
        w.ref_cur[p] = (int8_t)best_ref; w.fin_ref[p] = (int8_t)best_ref;
#pragma unroll 1
        for (int s = 0; s < ns; ++s)
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                w.fin_mv[p][s][k] = best_mv[p][s][k]; w.mv_cur[p][s][k] = best_mv[p][s][k];
                mvd[p][s][k] = (int16_t)(best_mv[p][s][k] - best_mvp[p][s][k]);
            }
    }
    int kind, cbp_luma = 0, cbp_chroma = 0, cbp = 0, mb_type;
    if (b_pskip) {
        // luma = prediction (_hl_codec_264_rdo_mb_reconstruct_luma_pskip); chroma already reconstructed by the zero check.  Nothing reads rec_y / luma_level /
        // blk_coded of a P_Skip macroblock afterwards (mb_commit writes no levels for it), so the samples go straight from the prediction to the picture.
        w.arg0 = 7; x.run(CMD_STORE, 96);
        w.cbp_luma4x4 = 0;
        kind = MBK_PSKIP; mb_type = 5;
    } else {
        x.run(CMD_PRED_INTER, 48);
        w.luma_skip_residual = best_sctr < 6;
        x.run(CMD_RECON_LUMA, 16);
        w.cbp_luma4x4 = 0;
        if (!w.luma_skip_residual)
#pragma unroll 1
            for (int b = 0; b < 16; ++b) w.cbp_luma4x4 |= w.blk_coded[b] << b;
        chroma_code(x, w);
        w.arg0 = 3; x.run(CMD_STORE, 96);
        cbp_luma = guess_cbp_luma(w.cbp_luma4x4, false);
        cbp_chroma = guess_cbp_chroma(w);
        cbp = (cbp_chroma << 4) | cbp_luma;
        if (cbp > 47) { cbp -= 16; cbp_chroma = cbp >> 4; }
        kind = MBK_INTER;
        mb_type = best_mode < 3 ? best_mode : 4;  // P_8x8ref0
        // late PSkip (rdo.c:1252-1262): the layout is 16x16 here, so the 16x16 derivation applies
        if (cbp == 0 && best_mode == 0 && mvd[0][0][0] == 0 && mvd[0][0][1] == 0) {
            int sx, sy;
            w.mode = 0;
            derive_pskip_mv(w, f, sx, sy);
            if (sx == best_mvp[0][0][0] && sy == best_mvp[0][0][1]) { kind = MBK_PSKIP; mb_type = 5; }
        }
    }
    HLB_LAP(w, 11);
    mb_commit(w, f, kind, cbp_luma, cbp_chroma, cbp, mb_type, mvd, best_dist, x.lane(), x.nlanes());
    HLB_LAP(w, 12);
}

template <class X>
HLB_HD void mb_encode(X& x, MbWork& w, const FrameCtx& f, int mb)
{
    mb_begin(w, f, mb, x.lane(), x.nlanes());
    x.sync();
    x.run(CMD_LOAD, 96);
    HLB_LAP(w, 0);
    if (f.is_p) mb_encode_p(x, w, f);
    else {
        double c;
        const int kind = mb_encode_intra(x, w, f, c);
        mb_commit_intra(w, f, kind, x.lane(), x.nlanes());
    }
}

}  // namespace hlb
