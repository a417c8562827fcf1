"""Packed-instruction formulations of the search's 4x4 primitives (hartallo_b200/csrc/hlb_fast.cuh: dot-product 6-tap filter and transform, 16-bit-pair vertical
filter, saturating pack, packed SAD) against the plain formulations of hlb_prims.cuh, which tests/test_oracle_pinned.py pins against the reference through the oracle.
CPU: tools/emu/check_fast.cpp (the C++ twins of the instructions); GPU: hlb200_dev_selftest (the instructions themselves)."""
import ctypes as C

import pytest


@pytest.mark.gpu
def test_packed_primitives_on_device():
    from hartallo_b200 import lib as hl
    lib = hl.load()
    hl.check(lib.hlb200_init(0), "init")
    bad = C.c_int(-1)
    hl.check(lib.hlb200_dev_selftest(96, 7, C.byref(bad)), "selftest")
    assert bad.value == 0, "%d packed-primitive results differ from the plain formulation on the device" % bad.value
