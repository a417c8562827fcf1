#!/usr/bin/env python
"""Makes tests/golden/svc_derive.npz: the inputs (reference-layer macroblock fields + layer geometry, trace tag 11) and the results (partition layout, refIdxL0, mvL0 per
macroblock, trace tag 6) of the reference's inter-layer motion derivation for every enhancement-layer P picture of a few small multi-layer encodes, run live through
oracle/_ref/hl_ref_driver (the unmodified reference).  Build container only.  usage: python tests/golden/make_golden_svc_derive.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svc_util as S  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svc_derive.npz")
CONFIGS = S.CONFIGS + [
    ("g2_3layer_64", ["--size", "64", "64", "--layers", "3", "--frames", "5", "--gen", "g2", "--seed", "998", "--qp", "38"]),     # first P picture refused (no prediction source)
    ("g1_2layer_112", ["--size", "112", "32", "--layers", "2", "--frames", "3", "--gen", "g1", "--seed", "1899", "--qp", "46"]),
    ("g2_ess_3to2", ["--size", "128", "64", "--layers", "2", "--frames", "4", "--gen", "g2", "--seed", "11", "--qp", "30", "--scale", "3", "2"]),   # general case: layers scaled 3:2
    ("g1_ess_3layer", ["--size", "64", "128", "--layers", "3", "--frames", "3", "--gen", "g1", "--seed", "5", "--qp", "26", "--scale", "3", "2"]),
    ("g1_3layer_80", ["--size", "80", "64", "--layers", "3", "--frames", "3", "--gen", "g1", "--seed", "5", "--qp", "31"]),
]
KEYS = ("geom", "base", "kind", "part_mode", "sub_mode", "ref_idx", "mv", "nparts", "nsub", "stale_parts", "motion", "valid")
d, index = {}, []
for name, args in CONFIGS:
    if (int(args[1]) & 15) or (int(args[2]) & 15):
        continue
    tr = "/tmp/golden_svc_derive.trace"
    S.run_driver_svc(args + ["--no-levels"], tr)
    for i, p in enumerate(S.derive_pictures_from_trace(tr)):
        tag = "%s.%d" % (name, i)
        index.append(tag)
        d[tag + ".meta"] = np.array([p["w"], p["h"], p["frame"], p["dqid"], p["spatial_change"]], np.int32)
        for k in KEYS:
            d[tag + "." + k] = p[k].view(np.uint8) if p[k].dtype.names else p[k]
    os.remove(tr)
d["index"] = np.array(index)
np.savez_compressed(OUT, **d)
print("%s: %d pictures, %d bytes" % (OUT, len(index), os.path.getsize(OUT)))
