"""recon MD5 per frame of a golden configuration through the library (debug aid): python tools/dbg_golden.py <name> [frames]"""
import sys, hashlib
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from hartallo_b200 import lib as hl, synth
name = sys.argv[1]
g = np.load("tests/golden/encoder_%s.npz" % name)
w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
frames = min(frames, int(sys.argv[2])) if len(sys.argv) > 2 else frames
gen = synth.make(str(g["gen"]), w, h, seed)
enc = hl.Encoder(w, h, qp=qp, me_range=me_range, refs=int(g["refs"]) if "refs" in g.files else 1)
for n in range(frames):
    rec, recon = enc.encode(gen.next(), want_recon=True)
    ok = hashlib.md5(recon.tobytes()).hexdigest() == str(g["recon_md5"][n])
    kb = np.flatnonzero(rec["mb_class"] != g["kind"][n])
    mbw = w // 16
    # first macroblock whose reconstruction differs is not known without the reference planes: report class / mad mismatches instead
    mad_bad = np.flatnonzero((g["kind"][n] != 0) & (rec["mad"] != g["mad"][n])) if "mad" in g.files else []
    print("frame %d recon %s | class mismatches %d %s | mad mismatches %d %s" % (n, "OK" if ok else "DIFF", len(kb), [(int(a) % mbw, int(a) // mbw) for a in kb[:4]], len(mad_bad), [(int(a) % mbw, int(a) // mbw) for a in mad_bad[:6]]), flush=True)
