"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/hlb200.h
declares; struct layouts seen by Python equal the C ones; with no CUDA device the library reports HL_ERROR_SYSTEM
instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import __graft_entry__ as g
    g.build()
    from hartallo_b200 import lib
    return lib


def test_exports_every_declared_symbol():
    lib = _lib()
    l = lib.load()
    hdr = open(os.path.join(ROOT, "include", "hlb200.h")).read()
    names = re.findall(r"HLB200_API\s+[\w\s\*]+?\b(hlb200_\w+)\s*\(", hdr)
    assert len(names) >= 24
    for n in names:
        assert hasattr(l, n), n


def test_struct_layouts_match_c():
    lib = _lib()
    src = '#include "hlb200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",sizeof(hlb200_mb_motion_t),' \
          'sizeof(hlb200_mb_coeffs_t),sizeof(hlb200_me_cand_t),sizeof(hlb200_me_cost_t),sizeof(hlb200_mb_record_t),sizeof(hlb200_slice_params_t),' \
          'offsetof(hlb200_mb_record_t,mv),offsetof(hlb200_mb_record_t,luma_level),sizeof(hlb200_svc_mb_state_t),' \
          'offsetof(hlb200_svc_mb_state_t,chroma_dc_level),sizeof(hlb200_svc_base_mb_t),offsetof(hlb200_svc_base_mb_t,mv),sizeof(hlb200_svc_layer_geom_t));return 0;}'
    exe = "/tmp/hlb200_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = [int(v) for v in subprocess.check_output([exe]).split()]
    want = [lib.MB_MOTION.itemsize, lib.MB_COEFFS.itemsize, lib.ME_CAND.itemsize, lib.ME_COST.itemsize, lib.MB_RECORD.itemsize, ctypes.sizeof(lib.SliceParams),
            lib.MB_RECORD.fields["mv"][1], lib.MB_RECORD.fields["luma_level"][1], lib.SVC_STATE.itemsize, lib.SVC_STATE.fields["chroma_dc_level"][1],
            lib.SVC_BASE_MB.itemsize, lib.SVC_BASE_MB.fields["mv"][1], lib.SVC_GEOM.itemsize]
    assert got == want


def test_no_cpu_fallback():
    lib = _lib()
    l = lib.load()
    if l.hlb200_device_count() > 0:
        return  # on the GPU box the compute tests cover the path
    assert l.hlb200_init(0) == 13  # HL_ERROR_SYSTEM
    assert b"no CPU fallback" in l.hlb200_last_error() or b"failed" in l.hlb200_last_error()
    try:
        lib.Stream(64, 48)
        assert False, "must raise without a GPU"
    except lib.Hlb200Error:
        pass


def test_synth_matches_driver_generators():
    """hartallo_b200/synth.py == generators inside oracle/ref_driver.c (only where the reference driver exists)"""
    drv = os.path.join(ROOT, "oracle", "_ref", "hl_ref_driver")
    if not os.path.exists(drv):
        return
    from hartallo_b200 import synth
    for gen, seed in (("g1", 1), ("g2", 3), ("g3", 7)):
        subprocess.check_call([drv, "--size", "64", "48", "--frames", "2", "--gen", gen, "--seed", str(seed), "--dump-input", "/tmp/hlb_in.yuv"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = np.fromfile("/tmp/hlb_in.yuv", np.uint8).reshape(2, -1)
        g = synth.make(gen, 64, 48, seed)
        for i in range(2):
            assert np.array_equal(g.next(), ref[i])


def test_argument_errors_are_reported_not_executed():
    """error behaviour of the boundary (HL_ERROR_INVALID_PARAMETER = 5 in the reference's enum order is what the glue maps): null pointers,
    sizes that are not a multiple of 16, QP outside 0..51, picture batches of zero or too many pictures are refused before anything is
    launched -- so this runs without a GPU"""
    lib = _lib()
    l = lib.load()
    inv = l.hlb200_dev_interp_luma(None, 64, 48, None, None, None)
    assert inv != 0
    assert l.hlb200_dev_interp_luma_batch(1, 64, 48, 0, 0, 1, 1, None) == inv          # no pictures
    assert l.hlb200_dev_interp_luma_batch(1, 60, 48, 1, 0, 1, 1, None) == inv          # width not a multiple of 16
    assert l.hlb200_dev_interp_chroma_batch(1, 1, 64, 48, 70000, 0, 1, 1, 1, None) == inv
    assert l.hlb200_dev_tq_recon_batch(1, 1, 1, 1, 1, 1, 64, 48, 1, 0, 52, 0, 1, 1, 1, 1, None) == inv   # QP 52
    assert l.hlb200_dev_svc_inter_recon_batch(1, 1, 1, 1, 1, 1, 64, 48, 1, 0, 31, 0, 1, None, 1, 1, 1, 1, None) == inv   # no state array
    assert l.hlb200_dev_svc_inter_recon_batch(1, 1, 1, 1, 1, 1, 64, 40, 1, 0, 31, 0, 1, 1, 1, 1, 1, 1, None) == inv      # height not a multiple of 16
    assert l.hlb200_svc_layer_picture(None, 0, 1, 31, 0, 1, None, None, None, 1) == inv                                  # no layer context
    assert l.hlb200_dev_svc_bl_recon_batch(1, 1, 1, 1, 1, 1, 64, 48, 0, 0, 31, 0, 1, 1, 1, 1, 1, None) == inv            # no pictures
    assert l.hlb200_dev_svc_derive_motion_batch(1, None, 64, 48, 1, 1, 1, 1, None) == inv                                # no geometry
    assert l.hlb200_dev_svc_derive_motion_batch(1, 1, 64, 40, 1, 1, 1, 1, None) == inv                                   # height not a multiple of 16
    assert l.hlb200_svc_layer_picture_derived(None, 0, 1, 31, 0, 1, 1, None, None, 1) == inv                             # no layer context
    assert l.hlb200_slice_encode_batch_async(None, None, 1) == inv
    assert l.hlb200_frame_set_device(None, 1, 1, 1) == inv
    prev = l.hlb200_slice_set_variant(1)
    assert l.hlb200_slice_set_variant(7) == 1      # out-of-range selects "automatic" and reports the previous setting
    assert l.hlb200_slice_set_variant(prev) == -1
