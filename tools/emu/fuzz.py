#!/usr/bin/env python
"""Differential fuzz of the per-macroblock control flow (CPU harness tools/emu/emu == the code of the slice kernel, lanes as loops)
against the unmodified reference encoder (oracle/_ref/hl_ref_driver): random picture sizes, QPs, search ranges, generators and seeds;
compares every reconstructed picture.  Build container only (needs oracle/_ref).  usage: fuzz.py [n_cases] [first_seed]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reftrace as rt  # noqa: E402
from hartallo_b200 import synth  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools", "emu"), "emu"], stdout=subprocess.DEVNULL)
bad = 0
for case in range(first, first + n_cases):
    rng = np.random.default_rng(case)
    w, h = int(rng.integers(1, 12)) * 16, int(rng.integers(1, 10)) * 16
    frames = int(rng.integers(2, 6))
    qp = int(rng.integers(12, 52))
    me_range = int(rng.choice([1, 2, 4, 8, 16, 24, 32, 64]))
    gen = str(rng.choice(["g1", "g2"] if not os.environ.get("FUZZ_GEN") else os.environ["FUZZ_GEN"].split(",")))
    seed = int(rng.integers(1, 10000))
    refs = int(rng.choice([1, 1, 1, 2, 4]))
    early = int(os.environ.get("FUZZ_EARLY_TERM", "0"))
    deblock = int(os.environ.get("FUZZ_DEBLOCK", "0"))
    pre = "/tmp/fuzz_%d" % case
    try:
        rt.run_driver(pre, w, h, frames, gen=gen, seed=seed, qp=qp, me_range=me_range, refs=refs, levels=False, state=False, early_term=early, deblock=deblock)
    except subprocess.CalledProcessError:
        # the reference itself gives up on some inputs (e.g. "Memory too short" in hl_codec_264_rbsp_avc_escape for noisy pictures at low QP)
        print("case %d: %dx%d %s seed %d qp %d -> reference encoder failed, skipped" % (case, w, h, gen, seed, qp), flush=True)
        continue
    ref = np.fromfile(pre + ".recon", np.uint8).reshape(frames, -1)
    g = synth.make(gen, w, h, seed)
    with open(pre + ".yuv", "wb") as f:
        for _ in range(frames):
            f.write(g.next().tobytes())
    subprocess.check_call([os.path.join(ROOT, "tools", "emu", "emu"), "--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp), "--me-range", str(me_range),
                           "--refs", str(refs), "--early-term", str(early), "--deblock", str(deblock), "--in", pre + ".yuv", "--out", pre + "_emu"], stdout=subprocess.DEVNULL)
    emu = np.fromfile(pre + "_emu.recon", np.uint8).reshape(frames, -1)
    ok = [bool(np.array_equal(emu[i], ref[i])) for i in range(frames)]
    print("case %d: %dx%d %s seed %d frames %d qp %d range %d refs %d -> %s" % (case, w, h, gen, seed, frames, qp, me_range, refs, "OK" if all(ok) else "MISMATCH %s" % ok), flush=True)
    bad += not all(ok)
    for ext in (".recon", ".yuv", ".trace", ".264", "_emu.recon", "_emu.rec", "_emu.st"):
        try:
            os.remove(pre + ext)
        except OSError:
            pass
print("%d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)
