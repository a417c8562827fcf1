"""helpers shared by the -m gpu tests"""
import numpy as np

from hartallo_b200 import lib as hl


def random_motion(rng, nmb, mv_range=140, far_every=0):
    m = np.zeros(nmb, hl.MB_MOTION)
    m["part_mode"] = rng.integers(0, 4, nmb)
    m["sub_mode"] = rng.integers(0, 4, (nmb, 4))
    m["mv"] = rng.integers(-mv_range, mv_range + 1, (nmb, 4, 4, 2))
    if far_every:
        idx = np.arange(0, nmb, far_every)
        m["mv"][idx] = rng.integers(-9000, 9001, (len(idx), 4, 4, 2))
    return m
