"""GPU parity of the ME candidate cost (hlb200_me_cost = hl_codec_264_me_ds_mb_compute_cost_mode, me_ds.c:527) for
integer-, half- and quarter-pel candidates of every partition size.

Counterpart of the reference's source/test_codec_h264_pel.c (which only dumps a half-pel picture for eyeballing): here
every candidate's distortion, residual bit count, Single_ctr, CBP and per-block TotalCoeff are compared bit-for-bit."""
import numpy as np
import pytest

from oracle_lib import load_oracle_mb

pytestmark = pytest.mark.gpu

SHAPES = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]


@pytest.mark.parametrize("qp", [12, 26, 31, 44])
def test_me_cost_parity(qp):
    from hartallo_b200 import lib as hl
    from hartallo_b200 import synth
    w, h = 176, 144
    rng = np.random.default_rng(qp)
    g = synth.G2(w, h, seed=4)
    ref, src = g.next(), g.next()
    st = hl.Stream(w, h, 1)
    st.upload_slot(0, ref)
    st.upload_frame(src)
    n = 900
    c = np.zeros(n, hl.ME_CAND)
    for i in range(n):
        pw, ph = SHAPES[i % 7]
        c[i]["mb_x"], c[i]["mb_y"] = rng.integers(0, w // 16), rng.integers(0, h // 16)
        c[i]["part_w"], c[i]["part_h"] = pw, ph
        c[i]["part_x"], c[i]["part_y"] = rng.integers(0, (16 - pw) // 4 + 1) * 4, rng.integers(0, (16 - ph) // 4 + 1) * 4
        # translation of the content between the two frames is (3,-2): candidates cluster around it + outliers
        if i % 6 == 0:
            c[i]["mv_x"], c[i]["mv_y"] = rng.integers(-600, 600), rng.integers(-600, 600)
        else:
            c[i]["mv_x"], c[i]["mv_y"] = 12 + rng.integers(-6, 7), -8 + rng.integers(-6, 7)
    got = st.me_cost(0, qp, c)
    o = load_oracle_mb()
    sy, ry = np.ascontiguousarray(src[:w * h]), np.ascontiguousarray(ref[:w * h])
    nz_seen = 0
    for i in range(n):
        d, b, s, cb = (np.zeros(1, np.int32) for _ in range(4))
        tc, t1 = np.zeros(16, np.uint8), np.zeros(16, np.uint8)
        k = c[i]
        o.hlo_me_cost(sy, ry, w, h, qp, int(k["mb_x"]), int(k["mb_y"]), int(k["part_x"]), int(k["part_y"]), int(k["part_w"]), int(k["part_h"]),
                      int(k["mv_x"]), int(k["mv_y"]), d, b, s, cb, tc, t1)
        assert (got[i]["dist"], got[i]["bits_rest"], got[i]["single_ctr"], got[i]["cbp_luma4x4"]) == (d[0], b[0], s[0], cb[0]), i
        assert np.array_equal(got[i]["total_coeff"], tc) and np.array_equal(got[i]["trailing_ones"], t1), i
        nz_seen += int(cb[0] != 0)
    assert nz_seen > 50
    st.close()
