#!/usr/bin/env python
"""Runs hlb200_dev_tma_probe in one process per case (a faulting kernel poisons its CUDA context): which way of handing the tensor map to the TMA unit works on this box."""
import ctypes as C
import subprocess
import sys

if len(sys.argv) > 1:
    import numpy as np
    import torch
    sys.path.insert(0, ".")
    from hartallo_b200 import lib as hl
    mode, w, h, x0, y0 = (int(a) for a in sys.argv[1:6])
    lib = hl.load()
    hl.check(lib.hlb200_init(0), "init")
    plane = torch.from_numpy(np.random.default_rng(1).integers(1, 256, w * h).astype(np.uint8)).cuda()
    bad = C.c_int(-1)
    rc = lib.hlb200_dev_tma_probe(plane.data_ptr(), w, h, x0, y0, mode, C.byref(bad))
    print("mode %d %dx%d at (%d,%d): rc %d mismatches %d %s" % (mode, w, h, x0, y0, rc, bad.value, lib.hlb200_last_error().decode() if rc else ""))
    sys.exit(0)
for mode in (2, 1, 0):
    for (w, h, x0, y0) in ((176, 144, 16, 16), (176, 144, -20, -7), (176, 144, 150, 120), (64, 48, 3, 0), (1920, 1088, 901, 517), (48, 16, -5, -5)):
        r = subprocess.run([sys.executable, sys.argv[0], str(mode), str(w), str(h), str(x0), str(y0)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output (rc %d)" % r.returncode, flush=True)
