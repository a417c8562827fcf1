#!/usr/bin/env python
"""Differential fuzz ON THE GPU: random configurations (size, QP, search range, generator, seed, max_ref_frame, early termination, deblocking) encoded by the
unmodified reference (oracle/_ref/hl_ref_driver, run live on the box's CPU) and by the library with both slice-kernel variants; every reconstructed picture must
be identical.  The CPU twin (tools/emu/fuzz.py) runs the same per-macroblock source with lanes as loops; this one exercises the 32-lane reductions / ballots.
usage: python tools/gpu_fuzz.py [cases] [first]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reftrace as rt  # noqa: E402
from hartallo_b200 import lib as hl, synth  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = ran = 0
for case in range(first, first + n_cases):
    rng = np.random.default_rng(1000 + case)
    w, h = int(rng.integers(2, 14)) * 16, int(rng.integers(2, 10)) * 16
    frames = int(rng.integers(2, 5))
    qp = int(rng.integers(12, 52))
    me_range = int(rng.choice([1, 4, 8, 16, 32, 64]))
    gen = str(rng.choice(["g1", "g2", "g3"]))
    seed = int(rng.integers(1, 10000))
    refs = int(rng.choice([1, 1, 2, 4]))
    early, deblock = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    pre = "/tmp/gfuzz_%d" % case
    try:
        rt.run_driver(pre, w, h, frames, gen=gen, seed=seed, qp=qp, me_range=me_range, refs=refs, levels=False, state=False, early_term=early, deblock=deblock)
    except Exception:
        print("case %d: reference encoder failed, skipped" % case, flush=True)
        continue
    ref = np.fromfile(pre + ".recon", np.uint8).reshape(frames, -1)
    g = synth.make(gen, w, h, seed)
    fr = [g.next() for _ in range(frames)]
    ok = True
    for variant in (0, 1):
        prev = hl.load().hlb200_slice_set_variant(variant)
        enc = hl.Encoder(w, h, qp=qp, me_range=me_range, refs=refs, early_term=early, deblock=deblock)
        for n in range(frames):
            _, recon = enc.encode(fr[n], want_recon=True)
            if not np.array_equal(recon, ref[n]):
                ok = False
                print("  MISMATCH case %d variant %d frame %d" % (case, variant, n), flush=True)
                break
        enc.close()
        hl.load().hlb200_slice_set_variant(prev)
    ran += 1
    bad += not ok
    print("case %d: %dx%d %s seed %d frames %d qp %d range %d refs %d early %d deblock %d -> %s" % (case, w, h, gen, seed, frames, qp, me_range, refs, early, deblock, "OK" if ok else "MISMATCH"), flush=True)
    for ext in (".recon", ".trace", ".264"):
        try:
            os.remove(pre + ext)
        except OSError:
            pass
print("%d cases run, %d mismatches" % (ran, bad))
sys.exit(1 if bad else 0)
