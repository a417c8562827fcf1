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
for i, e in enumerate(encs[:1]):
    rec = np.zeros(e.st.nmb, hl.MB_RECORD)
    hl.check(e.st.lib.hlb200_records_download(e.st.ctx, hl.ptr(rec)), "dl")
    t0 = rec["t_start_ns"].astype(np.int64); t1 = rec["t_end_ns"].astype(np.int64)
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
        for k, name in enumerate(["pskip", "inter"]):
            m = rec["mb_class"] == k
            cyc = (t1[m] - t0[m]).astype(np.float64) * 1.965   # ns -> cycles at 1965 MHz
            print("%-6s runs/MB %.0f  cycles in run() %.0f%% of MB (ME_EVAL runs %.0f%%), cycles per run %.0f, master-only cycles per run %.0f" % (
                name, rec["i16_dc_level"][m, 0].astype(np.uint16).mean(), 100 * rec["mad"][m].astype(np.uint32).sum() / cyc.sum(), 100 * rec["me_interp_ops"][m].sum() / cyc.sum(),
                rec["mad"][m].astype(np.uint32).sum() / rec["i16_dc_level"][m, 0].astype(np.uint16).sum(), (cyc.sum() - rec["mad"][m].astype(np.uint32).sum()) / rec["i16_dc_level"][m, 0].astype(np.uint16).sum()))
