#!/usr/bin/env python
"""Differential fuzz of the device-side CAVLC serialisation through the whole drop-in on the CPU (oracle/_ref/hl_glue_check_full: the glue object of
hl_b200_encoder with the device sources compiled as C++) against the unmodified reference encoder: random single-layer configurations, bitstream MD5.
usage: fuzz_bits.py [n_cases] [first_seed]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "hl_ref_driver")
GLUE = os.path.join(ROOT, "oracle", "_ref", "hl_glue_check_full")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = done = 0
for case in range(first, first + n_cases):
    rng = np.random.default_rng(case)
    w, h = int(rng.integers(1, 12)) * 16, int(rng.integers(1, 10)) * 16
    args = ["--size", str(w), str(h), "--frames", str(int(rng.integers(2, 6))), "--qp", str(int(rng.integers(12, 52))), "--me-range", str(int(rng.choice([1, 2, 4, 8, 16, 32, 64]))),
            "--gen", str(rng.choice(["g1", "g2"])), "--seed", str(int(rng.integers(1, 10000))), "--refs", str(int(rng.choice([1, 1, 2, 4])))]
    r = subprocess.run([REF] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        print("case %d: %s -> reference encoder failed, skipped" % (case, " ".join(args)), flush=True)
        continue
    g = subprocess.run([GLUE] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rj = json.loads(r.stdout.strip().splitlines()[-1])
    gj = json.loads(g.stdout.strip().splitlines()[-1]) if g.returncode == 0 and g.stdout.strip() else {"md5": "failed: " + g.stderr[-200:], "bytes": -1}
    ok = (rj["md5"], rj["bytes"]) == (gj["md5"], gj["bytes"])
    print("case %d: %s -> %s (%d bytes)" % (case, " ".join(args), "OK" if ok else "MISMATCH " + str(gj["md5"])[:80], rj["bytes"]), flush=True)
    bad += not ok
    done += 1
print("%d cases compared, %d mismatches" % (done, bad))
sys.exit(1 if bad else 0)
