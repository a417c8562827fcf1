"""Generates tests/golden/kernels.npz from the UNMODIFIED reference kernels (oracle/_ref/libref_kernels.so, built from
/root/reference by oracle/build_ref.sh).  Run in the build container: python tests/golden/make_golden.py
The fixtures are small and committed; the GPU box and the CPU suite use them where the reference is absent."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle_lib import load_ref  # noqa: E402
from test_oracle_pinned import luma_cases, random_levels, stress_plane, PARTS  # noqa: E402


def main():
    r = load_ref()
    rng = np.random.default_rng(20261018)
    w, h = 64, 48
    out = {}
    plane = stress_plane(rng, h, w)
    cases = luma_cases(rng, 224, w, h)
    res = []
    for (xl, yl, pw, ph, mvx, mvy) in cases:
        b = np.zeros(256, np.uint8)
        assert r.ref_interp_luma(plane, w, h, xl, yl, pw, ph, mvx, mvy, b) == 0
        m = np.zeros((16, 16), np.uint8)
        m[:ph, :pw] = b.reshape(16, 16)[:ph, :pw]
        res.append(m.reshape(-1))
    out.update(luma_plane=plane, luma_w=w, luma_h=h, luma_cases=np.array(cases, np.int32), luma_out=np.array(res))
    u, v = stress_plane(rng, h // 2, w // 2), stress_plane(rng, h // 2, w // 2)
    ccases, cres = [], []
    for i in range(160):
        pw, ph = PARTS[i % 7]
        xl = int(rng.integers(0, (w - pw) // 4 + 1)) * 4
        yl = int(rng.integers(0, (h - ph) // 4 + 1)) * 4
        mvx, mvy = (int(rng.integers(-500, 500)), int(rng.integers(-500, 500))) if i % 4 == 0 else (int(rng.integers(-70, 70)), int(rng.integers(-70, 70)))
        ru, rv = np.zeros(256, np.int32), np.zeros(256, np.int32)
        assert r.ref_interp_chroma(u, v, w, h, xl, yl, pw // 2, ph // 2, mvx, mvy, ru, rv) == 0
        m = np.zeros((8, 8), np.uint8)
        m[:ph // 2, :pw // 2] = ru.reshape(16, 16)[:ph // 2, :pw // 2]
        ccases.append((xl, yl, pw, ph, mvx, mvy))
        cres.append(m.reshape(-1))
    out.update(chroma_u=u, chroma_cases=np.array(ccases, np.int32), chroma_out=np.array(cres))
    tq = {k: [] for k in ("res", "qp", "intra", "w", "z", "r")}
    for it in range(400):
        resd = rng.integers(-255, 256, 16).astype(np.int32)
        if it % 7 == 0:
            resd[:] = rng.choice([-255, 255, 0], 16)
        qp, intra = int(rng.integers(12, 52)), int(rng.integers(0, 2))
        wv, zv, rv = np.zeros(16, np.int32), np.zeros(16, np.int32), np.zeros(16, np.int32)
        r.ref_fwd4x4(resd, wv)
        r.ref_quant4x4(qp, intra, wv, zv)
        r.ref_dequant_inv4x4(qp, 1 - intra, 1, 0, 0, zv, rv)
        for k, val in zip(("res", "qp", "intra", "w", "z", "r"), (resd, qp, intra, wv, zv, rv)):
            tq[k].append(val)
    out.update({"tq_" + k: np.array(v, np.int32) for k, v in tq.items()})
    lvs, nAs, nBs, outs = [], [], [], []
    while len(lvs) < 600:
        lv = random_levels(rng)
        if not lv.any():
            continue
        nA, nB = int(rng.integers(-1, 17)), int(rng.integers(-1, 17))
        s, t = np.zeros(1, np.int32), np.zeros(1, np.int32)
        b = r.ref_cavlc_luma_bits(lv, nA, nB, s, t)
        lvs.append(lv); nAs.append(nA); nBs.append(nB); outs.append([b, int(s[0]), int(t[0])])
    out.update(cavlc_lv=np.array(lvs, np.int32), cavlc_nA=np.array(nAs, np.int32), cavlc_nB=np.array(nBs, np.int32), cavlc_out=np.array(outs, np.int32))
    sa, sb, so, sto = [], [], [], []
    for it in range(300):
        a, b = stress_plane(rng, 16, 16)[:4, :4].copy(), stress_plane(rng, 16, 16)[:4, :4].copy()
        sa.append(a.reshape(-1)); sb.append(b.reshape(-1))
        so.append(r.ref_sad4x4(a.reshape(-1), 4, b.reshape(-1), 4)); sto.append(r.ref_satd4x4(a.reshape(-1), 4, b.reshape(-1), 4))
    out.update(sad_a=np.array(sa, np.uint8), sad_b=np.array(sb, np.uint8), sad_out=np.array(so, np.int32), satd_out=np.array(sto, np.int32))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
