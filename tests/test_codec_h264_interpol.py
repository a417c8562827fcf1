"""GPU parity of the interpolation kernels (hlb200_interp_luma / hlb200_interp_chroma) against the oracle.

Counterpart of the reference's source/test_codec_h264_interpol.c (cpp-vs-variant equivalence of the 16 luma kernels on
a plane where about half the pixels are < 34), extended to the u8 4x4 kernels the encoder really uses, to every
partition shape, to chroma, and to MVs far outside the picture (origin clip).  Bit-exact."""
import numpy as np
import pytest

from gpu_util import random_motion
from oracle_lib import load_oracle, oracle_predict_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,seed", [(176, 144, 1), (352, 288, 2), (64, 48, 3)])
def test_interp_frame_parity(w, h, seed):
    from hartallo_b200 import lib as hl
    from hartallo_b200 import synth
    from test_oracle_pinned import stress_plane
    rng = np.random.default_rng(seed)
    ref = np.concatenate([stress_plane(rng, h, w).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1)])
    st = hl.Stream(w, h, 1)
    st.upload_slot(0, ref)
    o = load_oracle()
    nmb = (w // 16) * (h // 16)
    for far in (0, 5):
        motion = random_motion(rng, nmb, far_every=far)
        gy = st.interp_luma(0, motion)
        gu, gv = st.interp_chroma(0, motion)
        oy, ou, ov = oracle_predict_frame(o, ref, w, h, motion)
        assert np.array_equal(gy, oy)
        assert np.array_equal(gu, ou) and np.array_equal(gv, ov)
    st.close()


def test_interp_all_16_positions_16x16():
    """every fractional position on whole 16x16 partitions, incl. positions straddling each picture border"""
    from hartallo_b200 import lib as hl
    from test_oracle_pinned import stress_plane
    w, h = 64, 64
    rng = np.random.default_rng(7)
    ref = np.concatenate([stress_plane(rng, h, w).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1)])
    st = hl.Stream(w, h, 1)
    st.upload_slot(0, ref)
    o = load_oracle()
    nmb = 16
    for frac in range(16):
        for (ix, iy) in ((0, 0), (-9, 3), (5, -11), (30, 30), (-40, -40), (70, 2)):
            motion = np.zeros(nmb, hl.MB_MOTION)
            motion["mv"][:, :, :, 0] = ix * 4 + (frac & 3)
            motion["mv"][:, :, :, 1] = iy * 4 + (frac >> 2)
            gy = st.interp_luma(0, motion)
            oy, _, _ = oracle_predict_frame(o, ref, w, h, motion)
            assert np.array_equal(gy, oy), (frac, ix, iy)
    st.close()
