"""Pins the oracle restatement (oracle/hl_oracle.c) against the reference's own kernels.

Two legs:
  * live   -- against oracle/_ref/libref_kernels.so (the unmodified reference built by oracle/build_ref.sh), when it
              exists (build container; prebuilt file on the GPU box);
  * golden -- against tests/golden/kernels.npz, generated from the reference by tests/golden/make_golden.py and
              committed, so the pin survives where the reference is absent.

Named after the reference harnesses these replace: source/test_codec_h264_interpol.c (cpp-vs-variant equivalence on a
plane where about half the pixels are < 34), source/test_codec_264_transf.c, source/test_math.c.
"""
import os

import numpy as np
import pytest

from oracle_lib import have_ref, load_oracle, load_ref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kernels.npz")


def stress_plane(rng, h, w):
    """about half the pixels < 34 (as source/test_codec_h264_interpol.c:1310-1314), plus 0/255 patches"""
    a = rng.integers(0, 256, (h, w))
    lo = rng.integers(0, 34, (h, w))
    p = np.where(rng.integers(0, 2, (h, w)) == 1, lo, a).astype(np.uint8)
    for _ in range(12):
        y, x = rng.integers(0, h - 8), rng.integers(0, w - 8)
        p[y:y + 8, x:x + 8] = 0 if rng.integers(0, 2) else 255
    return p


PARTS = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]


def luma_cases(rng, n, w, h):
    out = []
    for i in range(n):
        pw, ph = PARTS[i % 7]
        xl = int(rng.integers(0, (w - pw) // 4 + 1)) * 4
        yl = int(rng.integers(0, (h - ph) // 4 + 1)) * 4
        if i % 5 == 0:  # far outside the picture: exercises the origin clip
            mvx, mvy = int(rng.integers(-4 * (w + 80), 4 * (w + 80))), int(rng.integers(-4 * (h + 80), 4 * (h + 80)))
        else:
            mvx, mvy = int(rng.integers(-140, 140)), int(rng.integers(-140, 140))
        out.append((xl, yl, pw, ph, mvx, mvy))
    return out


@pytest.mark.skipif(not have_ref(), reason="reference kernels not built here")
def test_interp_luma_all_positions_live():
    o, r = load_oracle(), load_ref()
    rng = np.random.default_rng(1)
    w, h = 176, 144
    plane = stress_plane(rng, h, w)
    seen = set()
    for (xl, yl, pw, ph, mvx, mvy) in luma_cases(rng, 1500, w, h):
        a = np.zeros(256, np.uint8)
        b = np.zeros(256, np.uint8)
        o.hlo_interp_luma(plane, w, h, xl, yl, pw, ph, mvx, mvy, a)
        assert r.ref_interp_luma(plane, w, h, xl, yl, pw, ph, mvx, mvy, b) == 0
        a2, b2 = a.reshape(16, 16)[:ph, :pw], b.reshape(16, 16)[:ph, :pw]
        assert np.array_equal(a2, b2), (xl, yl, pw, ph, mvx, mvy)
        seen.add((mvx & 3, mvy & 3))
    assert len(seen) == 16


@pytest.mark.skipif(not have_ref(), reason="reference kernels not built here")
def test_interp_chroma_live():
    o, r = load_oracle(), load_ref()
    rng = np.random.default_rng(2)
    w, h = 176, 144
    u, v = stress_plane(rng, h // 2, w // 2), stress_plane(rng, h // 2, w // 2)
    for i in range(1200):
        pw, ph = PARTS[i % 7]
        xl = int(rng.integers(0, (w - pw) // 4 + 1)) * 4
        yl = int(rng.integers(0, (h - ph) // 4 + 1)) * 4
        mvx, mvy = (int(rng.integers(-900, 900)), int(rng.integers(-900, 900))) if i % 4 == 0 else (int(rng.integers(-70, 70)), int(rng.integers(-70, 70)))
        ru, rv = np.zeros(256, np.int32), np.zeros(256, np.int32)
        assert r.ref_interp_chroma(u, v, w, h, xl, yl, pw // 2, ph // 2, mvx, mvy, ru, rv) == 0
        for plane, rr in ((u, ru), (v, rv)):
            a = np.zeros(64, np.uint8)
            o.hlo_interp_chroma(plane, w // 2, h // 2, xl, yl, pw // 2, ph // 2, mvx, mvy, a)
            assert np.array_equal(a.reshape(8, 8)[:ph // 2, :pw // 2].astype(np.int32), rr.reshape(16, 16)[:ph // 2, :pw // 2]), (xl, yl, pw, ph, mvx, mvy)


@pytest.mark.skipif(not have_ref(), reason="reference kernels not built here")
def test_transform_quant_live():
    o, r = load_oracle(), load_ref()
    rng = np.random.default_rng(3)
    for it in range(3000):
        res = rng.integers(-255, 256, 16).astype(np.int32)
        if it % 7 == 0:
            res[:] = rng.choice([-255, 255, 0], 16)
        qp = int(rng.integers(12, 52))
        intra = int(rng.integers(0, 2))
        wa, wb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_fwd4x4(res, wa)
        r.ref_fwd4x4(res, wb)
        assert np.array_equal(wa, wb)
        za, zb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_quant4x4(qp, intra, wa, za)
        r.ref_quant4x4(qp, intra, wa, zb)
        assert np.array_equal(za, zb)
        for (luma, i16, keep) in ((1, 0, 0), (1, 1, 1), (0, 0, 1)):
            ra, rb = np.zeros(16, np.int32), np.zeros(16, np.int32)
            o.hlo_dequant_inv4x4(qp, keep, za, ra)
            r.ref_dequant_inv4x4(qp, 1 - intra, luma, i16, 0, za, rb)
            assert np.array_equal(ra, rb), (qp, luma, i16)
        # DC paths
        dc = rng.integers(-4080, 4081, 16).astype(np.int32)
        ha, hb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_hadamard4x4_dc_luma(dc, ha)
        r.ref_hadamard4x4_dc_luma(dc, hb)
        assert np.array_equal(ha, hb)
        qa, qb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_quant_dc(qp, 1, ha, qa, 16)
        r.ref_quant_dc_luma(qp, 1, ha, qb)
        assert np.array_equal(qa, qb)
        sa, sb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_scale_luma_dc(qp, qa, sa)
        r.ref_scale_luma_dc(qp, qa, sb)
        assert np.array_equal(sa, sb)
        c4 = rng.integers(-4080, 4081, 4).astype(np.int32)
        h4a, q4a = np.zeros(4, np.int32), np.zeros(4, np.int32)
        h4b, q4b = np.zeros(4, np.int32), np.zeros(4, np.int32)
        o.hlo_hadamard2x2(c4, h4a)
        o.hlo_quant_dc(qp, intra, h4a, q4a, 4)
        r.ref_hadamard2x2_quant_dc_chroma(qp, intra, c4, h4b, q4b)
        assert np.array_equal(h4a, h4b) and np.array_equal(q4a, q4b)
        d4a, d4b = np.zeros(4, np.int32), np.zeros(4, np.int32)
        o.hlo_scale_chroma_dc(qp, q4a, d4a)
        r.ref_scale_chroma_dc(qp, q4a, d4b)
        assert np.array_equal(d4a, d4b)


@pytest.mark.skipif(not have_ref(), reason="reference kernels not built here")
def test_math_live():
    o, r = load_oracle(), load_ref()
    rng = np.random.default_rng(4)
    for it in range(2000):
        a = stress_plane(rng, 16, 16)
        b = stress_plane(rng, 16, 16)
        assert o.hlo_sad4x4(a, 16, b, 16) == r.ref_sad4x4(a, 16, b, 16)
        assert o.hlo_satd4x4(a, 16, b, 16) == r.ref_satd4x4(a, 16, b, 16)
        assert o.hlo_ssd4x4(a, 16, b, 16) == r.ref_ssd4x4(a, 16, b, 16)
        pa = a.ctypes.data + 16 * 3 + 2   # an 8x8 block whose 3x3 support stays inside the 16x16 plane
        assert o.hlo_homogeneity8x8(pa, 16) == r.ref_homogeneity8x8(pa, 16)
        pred = rng.choice(np.array([0, 1, 3, 250, 255, 128, 200, 10], np.uint8), 16).astype(np.uint8)
        res = rng.integers(-300, 300, 16).astype(np.int32)
        oa, ob = np.zeros(16, np.uint8), np.zeros(16, np.uint8)
        o.hlo_addclip_u8xi32(pred, res, oa)
        r.ref_addclip_u8xi32(pred, res, ob)
        assert np.array_equal(oa, ob)
        ia, ib = np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_addclip_i32(pred.astype(np.int32), res, ia)
        r.ref_addclip_i32(pred.astype(np.int32), res, ib)
        assert np.array_equal(ia, ib)
    # the wrap of SURVEY F7: the probe inputs of the survey
    pred = np.array([250, 3, 255, 0, 200, 10] + [0] * 10, np.uint8)
    res = np.array([10, -10, 1, -1, 100, -20] + [0] * 10, np.int32)
    out = np.zeros(16, np.uint8)
    o.hlo_addclip_u8xi32(pred, res, out)
    assert out[:6].tolist() == [4, 249, 0, 255, 44, 246]


def random_levels(rng):
    kind = rng.integers(0, 6)
    lv = np.zeros(16, np.int32)
    if kind == 0:
        lv[rng.integers(0, 16)] = rng.choice([-1, 1])
    elif kind == 1:
        n = rng.integers(1, 17)
        idx = rng.choice(16, n, replace=False)
        lv[idx] = rng.choice([-1, 1], n)
    elif kind == 2:
        n = rng.integers(1, 17)
        idx = rng.choice(16, n, replace=False)
        lv[idx] = rng.integers(-4, 5, n)
    elif kind == 3:
        n = rng.integers(1, 17)
        idx = rng.choice(16, n, replace=False)
        lv[idx] = rng.integers(-60, 61, n)
    elif kind == 4:
        lv[:] = rng.integers(-400, 401, 16)
    else:
        n = rng.integers(1, 6)
        lv[:n] = rng.integers(-3, 4, n)
    return lv


@pytest.mark.skipif(not have_ref(), reason="reference kernels not built here")
def test_cavlc_bits_live():
    o, r = load_oracle(), load_ref()
    rng = np.random.default_rng(5)
    n_checked = 0
    for it in range(6000):
        lv = random_levels(rng)
        if not lv.any():
            continue
        nA = int(rng.integers(-1, 17))
        nB = int(rng.integers(-1, 17))
        sa, ta = np.zeros(1, np.int32), np.zeros(1, np.int32)
        sb, tb = np.zeros(1, np.int32), np.zeros(1, np.int32)
        ba = o.hlo_cavlc_bits(lv, 16, o.hlo_nC(nA, nB), sa, ta)
        bb = r.ref_cavlc_luma_bits(lv, nA, nB, sb, tb)
        assert (ba, sa[0], ta[0]) == (bb, sb[0], tb[0]), (lv.tolist(), nA, nB)
        n_checked += 1
    assert n_checked > 5000


def test_golden_kernels():
    """oracle vs committed golden vectors generated from the reference (tests/golden/make_golden.py)"""
    o = load_oracle()
    g = np.load(GOLDEN)
    plane, w, h = g["luma_plane"], int(g["luma_w"]), int(g["luma_h"])
    for case, want in zip(g["luma_cases"], g["luma_out"]):
        xl, yl, pw, ph, mvx, mvy = [int(v) for v in case]
        a = np.zeros(256, np.uint8)
        o.hlo_interp_luma(plane, w, h, xl, yl, pw, ph, mvx, mvy, a)
        assert np.array_equal(a.reshape(16, 16)[:ph, :pw], want.reshape(16, 16)[:ph, :pw])
    u = g["chroma_u"]
    for case, want in zip(g["chroma_cases"], g["chroma_out"]):
        xl, yl, pw, ph, mvx, mvy = [int(v) for v in case]
        a = np.zeros(64, np.uint8)
        o.hlo_interp_chroma(u, w // 2, h // 2, xl, yl, pw // 2, ph // 2, mvx, mvy, a)
        assert np.array_equal(a.reshape(8, 8)[:ph // 2, :pw // 2], want.reshape(8, 8)[:ph // 2, :pw // 2])
    for res, qp, intra, wv, zv, rv in zip(g["tq_res"], g["tq_qp"], g["tq_intra"], g["tq_w"], g["tq_z"], g["tq_r"]):
        wa, za, ra = np.zeros(16, np.int32), np.zeros(16, np.int32), np.zeros(16, np.int32)
        o.hlo_fwd4x4(res, wa)
        o.hlo_quant4x4(int(qp), int(intra), wa, za)
        o.hlo_dequant_inv4x4(int(qp), 0, za, ra)
        assert np.array_equal(wa, wv) and np.array_equal(za, zv) and np.array_equal(ra, rv)
    for lv, nA, nB, want in zip(g["cavlc_lv"], g["cavlc_nA"], g["cavlc_nB"], g["cavlc_out"]):
        s, t = np.zeros(1, np.int32), np.zeros(1, np.int32)
        b = o.hlo_cavlc_bits(lv, 16, o.hlo_nC(int(nA), int(nB)), s, t)
        assert [b, int(s[0]), int(t[0])] == want.tolist()
    for a, b, sad, satd in zip(g["sad_a"], g["sad_b"], g["sad_out"], g["satd_out"]):
        assert o.hlo_sad4x4(a, 4, b, 4) == sad and o.hlo_satd4x4(a, 4, b, 4) == satd
