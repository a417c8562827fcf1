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
    "ssd4x4": (C.c_int, [u8p, C.c_int, u8p, C.c_int]),
    "homogeneity8x8": (C.c_int, [C.c_void_p, C.c_int]),
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
    lib.ref_ssd4x4.restype = C.c_int
    lib.ref_ssd4x4.argtypes = [u8p, C.c_int, u8p, C.c_int]
    lib.ref_homogeneity8x8.restype = C.c_int
    lib.ref_homogeneity8x8.argtypes = [C.c_void_p, C.c_int]
    lib.ref_addclip_u8xi32.argtypes = [u8p, i32p, u8p]
    lib.ref_addclip_i32.argtypes = [i32p, i32p, i32p]
    lib.ref_cavlc_luma_bits.restype = C.c_int
    lib.ref_cavlc_luma_bits.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p]
    for n in ("ref_fwd4x4", "ref_quant4x4", "ref_dequant_inv4x4", "ref_hadamard4x4_dc_luma", "ref_quant_dc_luma", "ref_scale_luma_dc",
              "ref_hadamard2x2_quant_dc_chroma", "ref_scale_chroma_dc", "ref_addclip_u8xi32", "ref_addclip_i32"):
        getattr(lib, n).restype = None
    return lib


i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


def load_oracle_mb():
    """adds the macroblock-level oracle functions (hlo_recon_inter_mb, hlo_me_cost)"""
    lib = load_oracle()
    lib.hlo_recon_inter_mb.restype = None
    lib.hlo_recon_inter_mb.argtypes = [u8p, u8p, u8p, u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i16p, i16p, i16p, i32p, i32p, i32p, u8p, u8p, u8p]
    lib.hlo_me_cost.restype = None
    lib.hlo_me_cost.argtypes = [u8p, u8p] + [C.c_int] * 11 + [i32p, i32p, i32p, i32p, u8p, u8p]
    return lib


QPC = list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]


def chroma_qp(qp, offset=0):
    return QPC[min(51, max(0, qp + offset))]


def partitions(part_mode, sub_mode):
    """list of (mbPartIdx, subMbPartIdx, ox, oy, w, h) -- partition tables of source/h264/hl_codec_264_rdo.c:711-809"""
    if part_mode == 0:
        return [(0, 0, 0, 0, 16, 16)]
    if part_mode == 1:
        return [(0, 0, 0, 0, 16, 8), (1, 0, 0, 8, 16, 8)]
    if part_mode == 2:
        return [(0, 0, 0, 0, 8, 16), (1, 0, 8, 0, 8, 16)]
    out = []
    for p in range(4):
        px, py = (p & 1) * 8, (p >> 1) * 8
        sm = int(sub_mode[p])
        if sm == 0:
            out.append((p, 0, px, py, 8, 8))
        elif sm == 1:
            out += [(p, 0, px, py, 8, 4), (p, 1, px, py + 4, 8, 4)]
        elif sm == 2:
            out += [(p, 0, px, py, 4, 8), (p, 1, px + 4, py, 4, 8)]
        else:
            out += [(p, s, px + (s & 1) * 4, py + (s >> 1) * 4, 4, 4) for s in range(4)]
    return out


def oracle_predict_frame(o, ref_yuv, w, h, motion):
    """prediction planes of a whole frame from a per-MB motion field, via the oracle's interpolation"""
    ry = np.ascontiguousarray(ref_yuv[:w * h])
    ru = np.ascontiguousarray(ref_yuv[w * h:w * h * 5 // 4])
    rv = np.ascontiguousarray(ref_yuv[w * h * 5 // 4:])
    py = np.zeros((h, w), np.uint8)
    pu = np.zeros((h // 2, w // 2), np.uint8)
    pv = np.zeros((h // 2, w // 2), np.uint8)
    mbw = w // 16
    tmp = np.zeros(256, np.uint8)
    tc = np.zeros(64, np.uint8)
    for mb in range(len(motion)):
        mbx, mby = mb % mbw, mb // mbw
        m = motion[mb]
        for (p, s, ox, oy, pw, ph) in partitions(int(m["part_mode"]), m["sub_mode"]):
            mvx, mvy = int(m["mv"][p, s, 0]), int(m["mv"][p, s, 1])
            xl, yl = mbx * 16 + ox, mby * 16 + oy
            o.hlo_interp_luma(ry, w, h, xl, yl, pw, ph, mvx, mvy, tmp)
            py[yl:yl + ph, xl:xl + pw] = tmp.reshape(16, 16)[:ph, :pw]
            for plane, dst in ((ru, pu), (rv, pv)):
                o.hlo_interp_chroma(plane, w // 2, h // 2, xl, yl, pw // 2, ph // 2, mvx, mvy, tc)
                dst[yl // 2:yl // 2 + ph // 2, xl // 2:xl // 2 + pw // 2] = tc.reshape(8, 8)[:ph // 2, :pw // 2]
    return py, pu, pv
