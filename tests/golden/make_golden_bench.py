"""Generates tests/golden/encoder_1080p_bench.json: MD5 of every reconstructed picture of the UNMODIFIED reference encoder (oracle/_ref/hl_ref_driver) on the
sequence bench.py's stream 0 encodes (1920x1088, G1 seed 12345, QP 31, ME +-32, 1 ref; IDR + P pictures), so that bench.py can check its own output after the
timed region (`"parity_checked": true`).  Run in the build container:  python tests/golden/make_golden_bench.py [frames]"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reftrace as rt  # noqa: E402

W, H, QP, ME = 1920, 1088, 31, 32
if __name__ == "__main__":
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    pre = "/tmp/golden_bench"
    r = subprocess.run([rt.DRIVER, "--size", str(W), str(H), "--frames", str(frames), "--qp", str(QP), "--me-range", str(ME), "--refs", "1", "--gen", "g1",
                        "--recon", pre + ".recon", "--out", pre + ".264"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    s = json.loads(r.stdout.strip().splitlines()[-1])
    # the stream is a sequence of NAL units: the bitstream of the first n pictures is a prefix of it ("frame i layer 0: t ms, B bytes so far" on stderr)
    ends = [int(ln.split(",")[1].split()[0]) for ln in r.stderr.splitlines() if ln.startswith("frame ")]
    stream = open(pre + ".264", "rb").read()
    assert len(ends) == frames and ends[-1] == len(stream) == s["bytes"]
    fb = W * H * 3 // 2
    recon = np.memmap(pre + ".recon", np.uint8, "r").reshape(frames, fb)
    out = {"config": {"w": W, "h": H, "qp": QP, "me_range": ME, "refs": 1, "gen": "g1", "seed": 12345, "frames": frames},
           "bitstream_md5": s["md5"], "bitstream_bytes": s["bytes"],
           "recon_md5": [hashlib.md5(recon[n].tobytes()).hexdigest() for n in range(frames)],
           # bitstream_prefix[n - 1] = (bytes, MD5) of the bitstream of the first n pictures: what an encode of n pictures of this sequence must emit
           "bitstream_prefix": [[e, hashlib.md5(stream[:e]).hexdigest()] for e in ends]}
    json.dump(out, open(os.path.join(HERE, "encoder_1080p_bench.json"), "w"), indent=1)
    os.remove(pre + ".recon"); os.remove(pre + ".264")
    print("bench golden:", frames, "frames", s["md5"])
