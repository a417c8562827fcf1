#!/usr/bin/env python
"""Achieved HBM bandwidth of the stateless whole-picture kernels (luma / chroma interpolation, transform-quantisation-reconstruction) as a
function of the picture batch per launch (GPU box only).  Algorithmic bytes per macroblock: SURVEY.md 8(d) -- interpolation 768 B
(512 luma + 256 chroma), tq_recon 1,920 B, the fused SVC base-mode kernel (prediction + residual coding + reconstruction, no prediction planes)
1,920 B = 384 reference + 384 source + 384 reconstruction + 768 levels, against 2,688 B for the three separate kernels.  Peak = MEASURED_PEAKS.json hbm_gbs.  Prints one JSON line per batch size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hartallo_b200 import lib as hl, synth, workload  # noqa: E402

W, H, QP = 1920, 1088, 31
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    PEAK_KIND = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    PEAK, PEAK_KIND = 6549.0, "fallback (driver-measured copy bandwidth of this pool, BASELINE.md)"
lib = hl.load()
dev = torch.device("cuda:0")
ysz, csz = W * H, W * H // 4
fb = ysz + 2 * csz
nmb = (W // 16) * (H // 16)
g = synth.G1(W, H, seed=12345)
base = [g.next() for _ in range(4)]
m1 = workload.motion_field(W, H)
sp = torch.cuda.current_stream().cuda_stream
for n in [int(a) for a in (sys.argv[1:] or ["1", "8", "32", "128"])]:
    ref = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).to(dev)
    src = torch.from_numpy(np.stack([base[(i + 1) % 4] for i in range(n)])).to(dev)
    pred, rec = torch.zeros_like(ref), torch.zeros_like(ref)
    motion = torch.from_numpy(np.concatenate([m1] * n).view(np.uint8)).to(dev)
    coef = torch.zeros(n * nmb * hl.MB_COEFFS.itemsize, dtype=torch.uint8, device=dev)
    state = torch.zeros(n * nmb * hl.SVC_STATE.itemsize, dtype=torch.uint8, device=dev)
    b, p, s, r = ref.data_ptr(), pred.data_ptr(), src.data_ptr(), rec.data_ptr()
    ks = {
        "interp_luma": lambda: lib.hlb200_dev_interp_luma_batch(b, W, H, n, fb, motion.data_ptr(), p, sp),
        "interp_chroma": lambda: lib.hlb200_dev_interp_chroma_batch(b + ysz, b + ysz + csz, W, H, n, fb, motion.data_ptr(), p + ysz, p + ysz + csz, sp),
        "tq_recon": lambda: lib.hlb200_dev_tq_recon_batch(s, s + ysz, s + ysz + csz, p, p + ysz, p + ysz + csz, W, H, n, fb, QP, 0, coef.data_ptr(), r, r + ysz, r + ysz + csz, sp),
        "svc_inter_recon": lambda: lib.hlb200_dev_svc_inter_recon_batch(s, s + ysz, s + ysz + csz, b, b + ysz, b + ysz + csz, W, H, n, fb, QP, 0, motion.data_ptr(), state.data_ptr(),
                                                                        coef.data_ptr(), r, r + ysz, r + ysz + csz, sp),
        # Intra_Base resampling: the first quarter of each reference picture's bytes stands in for a (W/2 x H/2) reference-layer picture
        "svc_resample_intra": lambda: lib.hlb200_dev_svc_resample_intra_batch(b, b + ysz // 4, b + ysz // 4 + csz // 4, W // 2, H // 2, p, p + ysz, p + ysz + csz, W, H, n, fb, fb, 0, sp),
    }
    # algorithmic bytes per macroblock; resampling: 96 B of the reference layer read + 384 B of prediction written
    alg = {"interp_luma": 512, "interp_chroma": 256, "tq_recon": 1920, "svc_inter_recon": 1920, "svc_resample_intra": 480}
    # inter-layer motion derivation (dyadic: a (W/2 x H/2) reference layer): 84 B of reference-layer fields per FOUR macroblocks read, 80 B of motion + 2 B of flags written per macroblock
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import svc_util
    rngd = np.random.default_rng(5)
    dbase = torch.from_numpy(np.concatenate([svc_util.random_base_field(rngd, W // 2, H // 2, intra_frac=0.02)] * n).view(np.uint8).copy()).to(dev)
    dgeom = np.zeros(1, hl.SVC_GEOM)
    dgeom["ref_width"], dgeom["ref_height"], dgeom["scaled_width"], dgeom["scaled_height"], dgeom["level_idc"], dgeom["restricted"] = W // 2, H // 2, W, H, 40, 1
    dmotion = torch.zeros(n * nmb * hl.MB_MOTION.itemsize, dtype=torch.uint8, device=dev)
    dhad, dstatus = torch.zeros(n * nmb, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int32, device=dev)
    ks["svc_derive_motion"] = lambda: lib.hlb200_dev_svc_derive_motion_batch(dbase.data_ptr(), dgeom.ctypes.data, W, H, n, dhad.data_ptr(), dmotion.data_ptr(), dstatus.data_ptr(), sp)
    alg["svc_derive_motion"] = 103
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"pictures_per_launch": n, "peak_gbs": PEAK, "peak_kind": PEAK_KIND, "kernels": {}}
    only = os.environ.get("HBM_ONLY")   # restrict the table to one kernel (A/B runs of build variants selected with HLB200_LIB)
    for name, k in ks.items():
        if only and name != only:
            continue
        for _ in range(3):
            hl.check(k(), name)
        ts = []
        for _ in range(10):
            flush.fill_(1)                      # batches smaller than L2 (126 MB) would otherwise be served from it
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); hl.check(k(), name); e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        ms = float(np.median(ts))
        gbs = alg[name] * nmb * n / (ms * 1e-3) / 1e9
        out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_bytes": alg[name] * nmb * n, "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 4)}
    print(json.dumps(out))
    del ref, src, pred, rec, motion, coef, state
