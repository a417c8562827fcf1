"""GPU parity of the fused residual -> transform -> quant -> dequant -> inverse -> reconstruction kernel
(hlb200_tq_recon) and of the SAD / SATD kernel against the oracle.

Counterpart of the reference's source/test_codec_264_transf.c (inverse transform equivalence on a fixed matrix) and
source/test_math.c (SAD), widened to the whole inter-macroblock residual path incl. chroma DC and the chroma
single-coefficient elimination (rdo.c:2599-2649).  Bit-exact."""
import numpy as np
import pytest

from oracle_lib import chroma_qp, load_oracle_mb

pytestmark = pytest.mark.gpu


def oracle_tq_frame(o, src, pred, w, h, qp):
    from hartallo_b200 import lib as hl
    nmb = (w // 16) * (h // 16)
    mbw = w // 16
    sy, su, sv = [np.ascontiguousarray(a) for a in (src[:w * h], src[w * h:w * h * 5 // 4], src[w * h * 5 // 4:])]
    py, pu, pv = [np.ascontiguousarray(a) for a in (pred[:w * h], pred[w * h:w * h * 5 // 4], pred[w * h * 5 // 4:])]
    ry, ru, rv = np.zeros_like(sy), np.zeros_like(su), np.zeros_like(sv)
    coeffs = np.zeros(nmb, hl.MB_COEFFS)
    qpc = chroma_qp(qp)
    for mb in range(nmb):
        ll = np.zeros(256, np.int16)
        dc = np.zeros(8, np.int16)
        ac = np.zeros(128, np.int16)
        c4, cdc, cac = np.zeros(1, np.int32), np.zeros(2, np.int32), np.zeros(2, np.int32)
        o.hlo_recon_inter_mb(sy, su, sv, py, pu, pv, w, mb % mbw, mb // mbw, qp, qpc, 0, ll, dc, ac, c4, cdc, cac, ry, ru, rv)
        coeffs[mb]["luma_level"] = ll.reshape(16, 16)
        coeffs[mb]["chroma_dc_level"] = dc.reshape(2, 4)
        coeffs[mb]["chroma_ac_level"] = ac.reshape(2, 4, 16)
        coeffs[mb]["cbp_luma4x4"] = c4[0]
        coeffs[mb]["cbp_chroma_dc4x4"] = cdc
        coeffs[mb]["cbp_chroma_ac4x4"] = cac
    return coeffs, np.concatenate([ry, ru, rv])


@pytest.mark.parametrize("qp", [12, 23, 24, 31, 38, 51])
@pytest.mark.parametrize("kind", ["near", "far", "extreme"])
def test_tq_recon_parity(qp, kind):
    from hartallo_b200 import lib as hl
    from hartallo_b200 import synth
    w, h = 176, 144
    rng = np.random.default_rng(qp * 7 + len(kind))
    src = synth.G2(w, h, seed=2).next()
    if kind == "near":      # prediction = source + small noise: many zero / single-coefficient blocks
        pred = np.clip(src.astype(np.int32) + rng.integers(-3, 4, src.size) * (rng.integers(0, 4, src.size) == 0), 0, 255).astype(np.uint8)
    elif kind == "far":
        pred = synth.G2(w, h, seed=5).next()
    else:                   # saturated: residuals of +-255
        pred = np.where(rng.integers(0, 2, src.size) == 1, 0, 255).astype(np.uint8)
    st = hl.Stream(w, h, 1)
    st.upload_frame(src)
    gc, grec = st.tq_recon(qp, pred)
    oc, orec = oracle_tq_frame(load_oracle_mb(), src, pred, w, h, qp)
    assert np.array_equal(grec, orec)
    for f in ("luma_level", "cbp_luma4x4", "cbp_chroma_dc4x4", "cbp_chroma_ac4x4"):
        assert np.array_equal(gc[f], oc[f]), f
    # chroma DC levels are only defined where the plane has a DC flag (the reference leaves them stale otherwise)
    for c in range(2):
        sel = oc["cbp_chroma_dc4x4"][:, c] != 0
        assert np.array_equal(gc["chroma_dc_level"][sel, c], oc["chroma_dc_level"][sel, c])
    assert np.array_equal(gc["chroma_ac_level"][..., :15], oc["chroma_ac_level"][..., :15])
    if kind == "near" and qp >= 31:  # the content must exercise the all-zero / eliminated paths
        assert (oc["cbp_chroma_ac4x4"] == 0).any() and (oc["cbp_luma4x4"] == 0).any()
    st.close()


def test_sad_satd_parity():
    from hartallo_b200 import lib as hl
    from hartallo_b200 import synth
    w, h = 352, 288
    o = load_oracle_mb()
    a = synth.G2(w, h, seed=1).next()
    b = synth.G2(w, h, seed=9).next()
    st = hl.Stream(w, h, 1)
    st.upload_frame(a)
    ya, yb = a[:w * h].reshape(h, w), np.ascontiguousarray(b[:w * h].reshape(h, w))
    gs, gt, gq = st.sad4x4(yb), st.sad4x4(yb, satd=True), st.sad4x4(yb, satd=2)
    gh = st.homogeneity8x8()
    ya = np.ascontiguousarray(ya)
    for by in range(0, h // 4, 3):
        for bx in range(w // 4):
            pa, pb = ya[by * 4:, bx * 4:], yb[by * 4:, bx * 4:]
            assert gs[by, bx] == o.hlo_sad4x4(np.ascontiguousarray(pa[:4, :4]), 4, np.ascontiguousarray(pb[:4, :4]), 4)
            assert gt[by, bx] == o.hlo_satd4x4(np.ascontiguousarray(pa[:4, :4]), 4, np.ascontiguousarray(pb[:4, :4]), 4)
            assert gq[by, bx] == o.hlo_ssd4x4(np.ascontiguousarray(pa[:4, :4]), 4, np.ascontiguousarray(pb[:4, :4]), 4)
    # edge-map homogeneity of every 8x8 block (hl_math.c:470): interior blocks against the oracle, border blocks flagged
    assert (gh[0] == -1).all() and (gh[-1] == -1).all() and (gh[:, 0] == -1).all() and (gh[:, -1] == -1).all()
    for by in range(1, h // 8 - 1):
        for bx in range(1, w // 8 - 1):
            assert gh[by, bx] == o.hlo_homogeneity8x8(ya.ctypes.data + by * 8 * w + bx * 8, w), (by, bx)
    # size-independent property at full size: SAD(a, a) == 0 everywhere
    assert not st.sad4x4(np.ascontiguousarray(ya)).any()
    st.close()
