#!/usr/bin/env python
"""Scheduling analysis of one slice launch: per-macroblock latency by class, dependency wait, critical path (GPU box only)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hartallo_b200 import lib as hl, synth
W, H = 1920, 1088
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
encs = [hl.Encoder(W, H, qp=31, me_range=32) for _ in range(S)]
gens = [synth.G1(W, H, seed=12345 + 7919 * i) for i in range(S)]
for n in range(3):
    hl.encode_batch(encs, [g.next() for g in gens])
    for e in encs:
        hl.check(e.st.lib.hlb200_stream_sync(e.st.ctx), "sync")
mbw = W // 16
# whole-launch occupancy of the persistent CTAs: busy CTA-time / (span x grid), and MBs in flight over time
allrec = []
for e in encs:
    r_ = np.zeros(e.st.nmb, hl.MB_RECORD)
    hl.check(e.st.lib.hlb200_records_download(e.st.ctx, hl.ptr(r_)), "dl")
    allrec.append(r_)
A0 = np.concatenate([r_["t_start_ns"].astype(np.int64) for r_ in allrec]); A1 = np.concatenate([r_["t_end_ns"].astype(np.int64) for r_ in allrec])
b0 = A0.min(); A0 -= b0; A1 -= b0
A1 = np.where(A1 < A0, A1 + (1 << 32), A1)
span = A1.max()
grid = e.st.lib.hlb200_slice_grid_size()
print("launch: %d streams, span %.1f ms, grid %d CTAs, busy CTA-time %.1f s -> CTA utilisation %.1f%%, mean MB latency %.0f us" % (S, span / 1e6, grid, (A1 - A0).sum() / 1e9, 100.0 * (A1 - A0).sum() / (span * grid), (A1 - A0).mean() / 1e3))
ts = np.linspace(0, span, 21)[1:-1]
print("MBs in flight at 5%..95% of the span:", [int(((A0 <= t) & (A1 > t)).sum()) for t in ts])
fin = np.array([ (r_["t_end_ns"].astype(np.int64) - b0).max() for r_ in allrec]) / 1e6
print("per-stream finish time ms: min %.1f median %.1f max %.1f" % (fin.min(), np.median(fin), fin.max()))
for i, e in enumerate(encs[:1]):
    rec = np.zeros(e.st.nmb, hl.MB_RECORD)
    hl.check(e.st.lib.hlb200_records_download(e.st.ctx, hl.ptr(rec)), "dl")
    t0 = rec["t_start_ns"].astype(np.int64); t1 = rec["t_end_ns"].astype(np.int64)
    t1 = np.where(t1 < t0, t1 + (1 << 32), t1)
    base = t0.min(); t0 -= base; t1 -= base
    lat = (t1 - t0) / 1e3
    print("frame span %.1f ms" % ((t1.max()) / 1e6))
    for k, name in enumerate(["pskip", "inter", "i16", "i4"]):
        m = rec["mb_class"] == k
        if m.any():
            print("%-6s n=%5d latency us: mean %.0f median %.0f p90 %.0f max %.0f | cands mean %.0f trials mean %.0f" % (name, m.sum(), lat[m].mean(), np.median(lat[m]), np.percentile(lat[m], 90), lat[m].max(), rec["me_candidates"][m].mean(), rec["me_trials"][m].mean()))
    # dependency wait: start - max(end of left, end of top-right)
    x = np.arange(e.st.nmb) % mbw; y = np.arange(e.st.nmb) // mbw
    dep = np.zeros(e.st.nmb, np.int64)
    left = np.where(x > 0, t1[np.maximum(np.arange(e.st.nmb) - 1, 0)], 0)
    tr_idx = np.where(x < mbw - 1, np.arange(e.st.nmb) - mbw + 1, np.arange(e.st.nmb) - mbw)
    tr = np.where(y > 0, t1[np.maximum(tr_idx, 0)], 0)
    ready = np.maximum(left, tr)
    wait = (t0 - ready) / 1e3
    print("queue wait after ready, us: mean %.1f median %.1f p99 %.1f" % (wait[1:].mean(), np.median(wait[1:]), np.percentile(wait[1:], 99)))
    # critical path: walk back from the last finishing MB through the dependency that finished last
    a = int(np.argmax(t1)); path = []
    while True:
        path.append(a)
        cand = []
        if x[a] > 0: cand.append(a - 1)
        if y[a] > 0: cand.append(a - mbw + 1 if x[a] < mbw - 1 else a - mbw)
        if not cand: break
        a = max(cand, key=lambda c: t1[c])
    path = np.array(path)
    print("critical path: %d MBs, sum latency %.1f ms, classes %s, total cands %d" % (len(path), lat[path].sum() / 1e3, np.bincount(rec["mb_class"][path], minlength=4), rec["me_candidates"][path].sum()))
    print("us per candidate step (inter MBs): %.1f" % (lat[rec["mb_class"] == 1].sum() / max(1, rec["me_candidates"][rec["mb_class"] == 1].sum())))
    if os.environ.get("HLB200_LIB", "").endswith("_prof.so"):
        names = ["begin/load", "search ctl", "eval prelude", "tile", "trial run", "scan+token", "cost", "find tail", "mode bookkeeping", "pskip chroma", "intra", "final recon", "commit"]
        for k, name in enumerate(["pskip", "inter"]):
            m = rec["mb_class"] == k
            laps = rec["i16_ac_level"][m].reshape(m.sum(), -1).view(np.uint32)
            cyc = laps[:, :13].astype(np.float64).mean(axis=0); cnt = laps[:, 16:29].astype(np.float64).mean(axis=0)
            tot = ((t1[m] - t0[m]).astype(np.float64) * 1.965).mean()
            print("%s: %.0f cycles per MB; sections (cycles/MB, %% of MB, visits, cycles/visit):" % (name, tot))
            for i, nm in enumerate(names):
                print("   %-18s %9.0f %5.1f%% %7.1f %8.0f" % (nm, cyc[i], 100 * cyc[i] / tot, cnt[i], cyc[i] / max(cnt[i], 1e-9)))
        for k, name in enumerate(["pskip", "inter"]):
            m = rec["mb_class"] == k
            cyc = (t1[m] - t0[m]).astype(np.float64) * 1.965   # ns -> cycles at 1965 MHz
            print("%-6s runs/MB %.0f  cycles in run() %.0f%% of MB (ME_EVAL runs %.0f%%), cycles per run %.0f, master-only cycles per run %.0f" % (
                name, rec["i16_dc_level"][m, 0].astype(np.uint16).mean(), 100 * rec["mad"][m].astype(np.uint32).sum() / cyc.sum(), 100 * rec["me_interp_ops"][m].sum() / cyc.sum(),
                rec["mad"][m].astype(np.uint32).sum() / rec["i16_dc_level"][m, 0].astype(np.uint16).sum(), (cyc.sum() - rec["mad"][m].astype(np.uint32).sum()) / rec["i16_dc_level"][m, 0].astype(np.uint16).sum()))
