#!/usr/bin/env python
"""Differential fuzz of the inter-layer motion derivation (hartallo_b200/csrc/hlb_svc_derive.cuh compiled as C++, tools/emu/svc_emu.cpp) against the unmodified reference
encoder run live with 2 or 3 spatial layers: random base sizes, QPs, generators, seeds, frame counts.  Per enhancement-layer P picture it feeds the reference layer's
macroblock fields (trace tag 11) to the device source and compares partition layout, refIdxL0 and mvL0 of every macroblock with what the reference derived (tag 6), the
inherited-prediction markers with the glue's, and the status bits with the refusals the glue would raise.  Build container only (needs oracle/_ref).
usage: fuzz_derive.py [n_cases] [first_seed]"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reftrace as rt  # noqa: E402
import svc_util as S  # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "tools", "emu", "libsvc_emu.so"))
P = ctypes.c_void_p


def emu_derive(d, had):
    nmb = (d["w"] // 16) * (d["h"] // 16)
    mot, st = np.zeros(nmb, S.MB_MOTION), ctypes.c_int32(0)
    rc = lib.svc_emu_derive_motion(d["base"].ctypes.data_as(P), d["geom"].ctypes.data_as(P), d["w"], d["h"], had.ctypes.data_as(P), mot.ctypes.data_as(P), ctypes.byref(st))
    assert rc == 0, rc
    return mot, st.value


n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = n_mb = n_pic = 0
status_seen = {}
for case in range(first, first + n_cases):
    rng = np.random.default_rng(case)
    layers = int(rng.choice([2, 2, 3]))
    w, h = int(rng.integers(1, 9 if layers == 2 else 6)) * 16, int(rng.integers(1, 8 if layers == 2 else 5)) * 16
    frames, qp, gen, seed = int(rng.integers(2, 6)), int(rng.integers(16, 52)), str(rng.choice(["g1", "g2"])), int(rng.integers(1, 10000))
    args = ["--size", str(w), str(h), "--layers", str(layers), "--frames", str(frames), "--gen", gen, "--seed", str(seed), "--qp", str(qp)]
    if os.environ.get("FUZZ_SCALE"):   # e.g. "3 2": every layer 1.5 times the one below (extended spatial scalability: the general case of the derivation)
        n, d = (int(v) for v in os.environ["FUZZ_SCALE"].split())
        unit = 16 * d ** (layers - 1)
        w, h = int(rng.integers(1, 5)) * unit, int(rng.integers(1, 4)) * unit
        args = ["--size", str(w), str(h), "--layers", str(layers), "--frames", str(frames), "--gen", gen, "--seed", str(seed), "--qp", str(qp), "--scale", str(n), str(d)]
    tr = "/tmp/fuzz_derive_%d.trace" % case
    try:
        subprocess.run([rt.DRIVER] + args + ["--trace", tr, "--no-levels"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    except subprocess.CalledProcessError:
        print("case %d: %s -> reference encoder failed, skipped" % (case, " ".join(args)), flush=True)
        continue
    had, ok = {}, True
    try:
        for d in S.derive_pictures_from_trace(tr):
            hp = had.setdefault(d["dqid"], np.zeros((d["w"] // 16) * (d["h"] // 16), np.uint8))
            m, st = emu_derive(d, hp)
            n_mb += S.compare_derived(d, m, st, "device source on the CPU")
            n_pic += 1
            status_seen[st] = status_seen.get(st, 0) + 1
    except AssertionError as e:
        ok = False
        print(str(e)[:400])
    bad += not ok
    print("case %d: %s -> %s" % (case, " ".join(args), "ok" if ok else "MISMATCH"), flush=True)
    os.remove(tr)
print("fuzz_derive: %d cases, %d pictures, %d macroblocks compared, %d mismatching cases; pictures by status bits %s" % (n_cases, n_pic, n_mb, bad, status_seen))
sys.exit(1 if bad else 0)
