"""ctypes bindings for the oracle (TEST INFRASTRUCTURE ONLY).

* ``oracle/libhl_oracle.so``            -- CPU restatement (oracle/hl_oracle.c), always available (built by build()).
* ``oracle/_ref/libref_kernels.so``     -- the real reference kernels behind plain-pointer entry points; exists only
                                           where oracle/build_ref.sh could see the reference tree (and travels to the
                                           GPU box as a prebuilt file).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def _ensure_oracle():
    so = os.path.join(ORACLE_DIR, "libhl_oracle.so")
    src = os.path.join(ORACLE_DIR, "hl_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libhl_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_SIGS = {
    "interp_luma": (None, [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]),
    "fwd4x4": (None, [i32p, i32p]),
    "quant4x4": (None, [C.c_int, C.c_int, i32p, i32p]),
    "hadamard4x4_dc_luma": (None, [i32p, i32p]),
    "scale_luma_dc": (None, [C.c_int, i32p, i32p]),
    "scale_chroma_dc": (None, [C.c_int, i32p, i32p]),
    "sad4x4": (C.c_int, [u8p, C.c_int, u8p, C.c_int]),
    "satd4x4": (C.c_int, [u8p, C.c_int, u8p, C.c_int]),
    "addclip_u8xi32": (None, [u8p, i32p, u8p]),
    "addclip_i32": (None, [i32p, i32p, i32p]),
}


def load_oracle():
    lib = C.CDLL(_ensure_oracle())
    for name, (res, args) in _SIGS.items():
        f = getattr(lib, "hlo_" + name)
        f.restype, f.argtypes = res, args
    lib.hlo_interp_chroma.restype = None
    lib.hlo_interp_chroma.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    lib.hlo_quant_dc.restype = None
    lib.hlo_quant_dc.argtypes = [C.c_int, C.c_int, i32p, i32p, C.c_int]
    lib.hlo_dequant_inv4x4.restype = None
    lib.hlo_dequant_inv4x4.argtypes = [C.c_int, C.c_int, i32p, i32p]
    lib.hlo_dequant4x4.restype = None
    lib.hlo_dequant4x4.argtypes = [C.c_int, C.c_int, i32p, i32p]
    lib.hlo_inv4x4.restype = None
    lib.hlo_inv4x4.argtypes = [i32p, i32p]
    lib.hlo_hadamard2x2.restype = None
    lib.hlo_hadamard2x2.argtypes = [i32p, i32p]
    lib.hlo_zigzag.restype = None
    lib.hlo_zigzag.argtypes = [i32p, i32p]
    lib.hlo_inv_zigzag.restype = None
    lib.hlo_inv_zigzag.argtypes = [i32p, i32p]
    lib.hlo_cavlc_bits.restype = C.c_int
    lib.hlo_cavlc_bits.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p]
    lib.hlo_nC.restype = C.c_int
    lib.hlo_nC.argtypes = [C.c_int, C.c_int]
    lib.hlo_trial_luma4x4.restype = C.c_int
    lib.hlo_trial_luma4x4.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p]
    return lib


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_kernels.so"))


def load_ref():
    lib = C.CDLL(os.path.join(REF_DIR, "libref_kernels.so"))
    lib.ref_init.restype = C.c_int
    assert lib.ref_init() == 0
    lib.ref_interp_luma.restype = C.c_int
    lib.ref_interp_luma.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    lib.ref_interp_chroma.restype = C.c_int
    lib.ref_interp_chroma.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p]
    lib.ref_fwd4x4.argtypes = [i32p, i32p]
    lib.ref_quant4x4.argtypes = [C.c_int, C.c_int, i32p, i32p]
    lib.ref_dequant_inv4x4.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p]
    lib.ref_hadamard4x4_dc_luma.argtypes = [i32p, i32p]
    lib.ref_quant_dc_luma.argtypes = [C.c_int, C.c_int, i32p, i32p]
    lib.ref_scale_luma_dc.argtypes = [C.c_int, i32p, i32p]
    lib.ref_hadamard2x2_quant_dc_chroma.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p]
    lib.ref_scale_chroma_dc.argtypes = [C.c_int, i32p, i32p]
    lib.ref_sad4x4.restype = C.c_int
    lib.ref_sad4x4.argtypes = [u8p, C.c_int, u8p, C.c_int]
    lib.ref_satd4x4.restype = C.c_int
    lib.ref_satd4x4.argtypes = [u8p, C.c_int, u8p, C.c_int]
    lib.ref_addclip_u8xi32.argtypes = [u8p, i32p, u8p]
    lib.ref_addclip_i32.argtypes = [i32p, i32p, i32p]
    lib.ref_cavlc_luma_bits.restype = C.c_int
    lib.ref_cavlc_luma_bits.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p]
    for n in ("ref_fwd4x4", "ref_quant4x4", "ref_dequant_inv4x4", "ref_hadamard4x4_dc_luma", "ref_quant_dc_luma", "ref_scale_luma_dc",
              "ref_hadamard2x2_quant_dc_chroma", "ref_scale_chroma_dc", "ref_addclip_u8xi32", "ref_addclip_i32"):
        getattr(lib, n).restype = None
    return lib
