import sys, time, os
import numpy as np
sys.path.insert(0, ".")
from hartallo_b200 import lib as hl, synth
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1088)
gen = sys.argv[3] if len(sys.argv) > 3 else "g1"
nfr = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t0 = time.time()
g = synth.make(gen, W, H, 3)
fr = [g.next() for _ in range(nfr)]
print("frames generated %.2fs" % (time.time() - t0), flush=True)
enc = hl.Encoder(W, H, qp=31, me_range=32)
print("encoder created %.2fs" % (time.time() - t0), flush=True)
for n in range(nfr):
    t1 = time.time()
    try:
        rec, recon = enc.encode(fr[n], want_recon=True)
        print("frame %d: %.3fs classes %s variant %d" % (n, time.time() - t1, np.bincount(rec["mb_class"], minlength=4), hl.load().hlb200_slice_last_variant()), flush=True)
    except Exception as e:
        print("frame %d failed after %.3fs: %s" % (n, time.time() - t1, str(e)[:300]), flush=True)
        try:
            print("status", enc.st.slice_status())
        except Exception as e2:
            print("status:", str(e2)[:300])
        break
