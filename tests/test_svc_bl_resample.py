"""Intra_Base resampling for SVC enhancement-layer I pictures (first half): the I_BL prediction planes from the reference layer's reconstruction,
_hl_codec_264_decode_svc_resample_intra_colour_comps (source/h264/hl_codec_264_decode_svc.c:2864-3200) with the sample locations of utils.c:1064-1157.

CPU tier: the oracle restatement (oracle/hl_oracle.c: hlo_svc_resample_intra_plane) against what the UNMODIFIED reference computed -- the committed fixture
(tests/golden/svc_inter.npz holds a 3-layer encode: reconstruction of layer 1 and the prediction the reference resampled from it for layer 2) and live 3-layer
encodes where oracle/_ref exists; the device source (hartallo_b200/csrc/hlb_svc.cuh: svc_resample_px, compiled as C++ by tools/emu) against the oracle.
GPU tier: hlb200_dev_svc_resample_intra_batch through the C-ABI against the fixture and against the oracle on random pictures (dyadic and non-dyadic ratios)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import reftrace
import svc_util
from oracle_lib import load_oracle_mb
from test_svc_inter import _emu_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "svc_inter.npz")


def oracle_resample(ref_yuv, rw, rh, w, h, level_idc=0):
    o = load_oracle_mb()
    o.hlo_svc_resample_intra_yuv.restype = None
    o.hlo_svc_resample_intra_yuv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    ref_yuv = np.ascontiguousarray(ref_yuv, np.uint8)
    out = np.zeros(w * h * 3 // 2, np.uint8)
    o.hlo_svc_resample_intra_yuv(ref_yuv.ctypes.data, rw, rh, w, h, level_idc, out.ctypes.data)
    return out


def emu_resample(ref_yuv, rw, rh, w, h, level_idc=0):
    e = _emu_lib()
    e.svc_emu_resample_plane.restype = C.c_int
    e.svc_emu_resample_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    ref_yuv = np.ascontiguousarray(ref_yuv, np.uint8)
    out = np.zeros(w * h * 3 // 2, np.uint8)
    ro, oo = [0, rw * rh, rw * rh * 5 // 4], [0, w * h, w * h * 5 // 4]
    for pl in range(3):
        c = pl != 0
        assert e.svc_emu_resample_plane(ref_yuv.ctypes.data + ro[pl], rw >> c, rh >> c, out.ctypes.data + oo[pl], w >> c, h >> c, int(c), level_idc) == 0
    return out


def golden_pair():
    """(reconstruction of layer 1, 64x64) -> (prediction of layer 2, 128x128) of the fixture's 3-layer I access unit"""
    g = np.load(GOLDEN, allow_pickle=True)
    lo, hi = g["g2_3layer.0.meta"], g["g2_3layer.1.meta"]
    assert (int(lo[3]), int(lo[4]), int(lo[5])) == (0, 16, 1) and (int(hi[3]), int(hi[4]), int(hi[5])) == (0, 32, 1)   # frame 0, DQId 16 / 32, I_BL pictures
    return svc_util.planes_of_mb(g["g2_3layer.0.rec"], int(lo[0]), int(lo[1])), (int(lo[0]), int(lo[1])), g["g2_3layer.1.ref"], (int(hi[0]), int(hi[1]))


def random_yuv(rng, w, h):
    p = rng.integers(0, 256, w * h * 3 // 2).astype(np.uint8)
    p[rng.random(p.size) < 0.2] = 255     # runs of extremes next to noise: both clip branches of (G-305), the negative taps of Table G-9
    p[rng.random(p.size) < 0.2] = 0
    return p


SIZES = [((16, 16), (32, 32)), ((64, 48), (128, 96)), ((176, 144), (352, 288)), ((48, 32), (80, 64)), ((32, 32), (32, 32)), ((64, 16), (96, 48)), ((16, 64), (128, 80))]


def test_oracle_vs_golden():
    base, (rw, rh), pred, (w, h) = golden_pair()
    assert np.array_equal(oracle_resample(base, rw, rh, w, h), pred)


@pytest.mark.skipif(not os.path.exists(reftrace.DRIVER), reason="oracle/_ref/hl_ref_driver only exists where the reference tree is available")
@pytest.mark.parametrize("args", [("96", "64", "g1", "5", "24"), ("160", "96", "g2", "7", "36"), ("32", "64", "g2", "11", "45"), ("176", "144", "g1", "1", "31")])
def test_oracle_vs_live_reference(tmp_path, args):
    """3 spatial layers: the prediction the reference resampled for layer 2's I picture must equal the oracle's resampling of layer 1's reconstruction"""
    tr = str(tmp_path / "t.trace")
    subprocess.run([reftrace.DRIVER, "--size", args[0], args[1], "--layers", "3", "--frames", "2", "--gen", args[2], "--seed", args[3], "--qp", args[4], "--trace", tr],
                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    pics = {p["dqid"]: p for p in svc_util.bl_pictures_from_trace(tr)}
    lo, hi = pics[16], pics[32]
    base = svc_util.planes_of_mb(lo["rec"], lo["w"], lo["h"])
    # level_idc as hl_codec_264_utils_guess_level (utils.c:15) assigns it to the top layer: above 30 once the picture exceeds 720x480
    level = 31 if (hi["w"] > 720 or hi["h"] > 480) else 30
    assert np.array_equal(oracle_resample(base, lo["w"], lo["h"], hi["w"], hi["h"], level), hi["ref"])


def test_device_source_on_cpu_vs_golden():
    base, (rw, rh), pred, (w, h) = golden_pair()
    assert np.array_equal(emu_resample(base, rw, rh, w, h), pred)


def _pow2(v):
    return v & (v - 1) == 0


@pytest.mark.parametrize("rsz,sz", SIZES + [((352, 288), (704, 576)), ((176, 144), (240, 208)), ((48, 80), (112, 96))])
def test_device_source_on_cpu_vs_oracle_random(rsz, sz):
    """both fixed-point precisions of (G-43): level_idc <= 30 (16 bits) and above (31 - ceil(log2(dimension)); what QCIF -> CIF -> 4CIF uses for its top layer),
    the latter wherever the reference's int32 arithmetic is defined (no power-of-two luma / chroma reference dimension)"""
    rng = np.random.default_rng(rsz[0] * 7 + sz[1])
    base = random_yuv(rng, *rsz)
    assert np.array_equal(emu_resample(base, rsz[0], rsz[1], sz[0], sz[1]), oracle_resample(base, rsz[0], rsz[1], sz[0], sz[1]))
    if not any(_pow2(d) for d in (rsz[0], rsz[1], rsz[0] // 2, rsz[1] // 2)):
        assert np.array_equal(emu_resample(base, rsz[0], rsz[1], sz[0], sz[1], 31), oracle_resample(base, rsz[0], rsz[1], sz[0], sz[1], 31))


def test_abi_arguments():
    """refused before anything is launched (runs without a GPU): null planes, sizes that are not multiples of 16, down-sampling, level_idc above 30 (no
    reference behaviour to pin that precision on), a frame stride that breaks the word stores"""
    from hartallo_b200 import lib as hl
    l = hl.load()
    inv = l.hlb200_dev_svc_resample_intra_batch(None, 4, 4, 64, 48, 4, 4, 4, 128, 96, 1, 0, 0, 0, None)
    assert inv != 0
    assert l.hlb200_dev_svc_resample_intra_batch(4, 4, 4, 60, 48, 4, 4, 4, 128, 96, 1, 0, 0, 0, None) == inv
    assert l.hlb200_dev_svc_resample_intra_batch(4, 4, 4, 64, 48, 4, 4, 4, 32, 96, 1, 0, 0, 0, None) == inv
    assert l.hlb200_dev_svc_resample_intra_batch(4, 4, 4, 64, 48, 4, 4, 4, 128, 96, 1, 0, 0, 40, None) == inv
    assert l.hlb200_dev_svc_resample_intra_batch(4, 4, 4, 64, 48, 4, 4, 4, 128, 96, 2, 4608, 18433, 0, None) == inv
    assert l.hlb200_dev_svc_resample_intra_batch(4, 4, 4, 64, 48, 4, 4, 4, 128, 96, 0, 0, 0, 0, None) == inv
    assert l.hlb200_svc_layer_picture_resampled(None, 1, 31, 0, 4, 4, 4, 64, 48, 0, 4) == inv     # no layer context


def gpu_resample(bases, rsz, sz):
    """list of reference-layer pictures of one size -> ONE launch of hlb200_dev_svc_resample_intra_batch"""
    import torch
    from hartallo_b200 import lib as hl
    lib = hl.load()
    (rw, rh), (w, h) = rsz, sz
    n = len(bases)
    dev = torch.device("cuda:0")
    d_ref = torch.from_numpy(np.stack(bases)).to(dev)
    rfb, fb = rw * rh * 3 // 2, w * h * 3 // 2
    d_out = torch.zeros(n * fb, dtype=torch.uint8, device=dev)
    r, o = d_ref.data_ptr(), d_out.data_ptr()
    hl.check(lib.hlb200_dev_svc_resample_intra_batch(r, r + rw * rh, r + rw * rh * 5 // 4, rw, rh, o, o + w * h, o + w * h * 5 // 4, w, h, n, rfb, fb, 0,
                                                     torch.cuda.current_stream().cuda_stream), "svc_resample_intra_batch")
    torch.cuda.synchronize()
    return d_out.cpu().numpy().reshape(n, fb)


@pytest.mark.gpu
def test_gpu_vs_golden():
    base, rsz, pred, sz = golden_pair()
    assert np.array_equal(gpu_resample([base], rsz, sz)[0], pred)


@pytest.mark.gpu
@pytest.mark.parametrize("rsz,sz,n", [(a, b, 1 + i % 3) for i, (a, b) in enumerate(SIZES)] + [((960, 544), (1920, 1088), 2)])
def test_gpu_vs_oracle_random(rsz, sz, n):
    rng = np.random.default_rng(rsz[0] * 7 + sz[1] + n)
    bases = [random_yuv(rng, *rsz) for _ in range(n)]
    got = gpu_resample(bases, rsz, sz)
    for i, b in enumerate(bases):
        assert np.array_equal(got[i], oracle_resample(b, rsz[0], rsz[1], sz[0], sz[1])), i


@pytest.mark.gpu
def test_gpu_layer_picture_resampled_vs_golden():
    """hlb200_svc_layer_picture_resampled (what host/hlb200_glue.c calls for an enhancement-layer I picture): the fixture's layer-2 I picture from layer 1's
    reconstruction must give the coefficients and the reconstruction the reference left, and exactly what the host-resampled entry point gives"""
    from hartallo_b200 import lib as hl
    base, (rw, rh), pred, (w, h) = golden_pair()
    p = [q for q in svc_util.load_golden() if q["name"] == "g2_3layer.1"][0]
    st = hl.Stream(w, h, 1)
    st.upload_frame(p["src"])
    coef, rec = st.svc_layer_picture_resampled(p["qp"], base, rw, rh)
    st2 = hl.Stream(w, h, 1)
    st2.upload_frame(p["src"])
    coef2, rec2 = st2.svc_layer_picture(p["qp"], pred_yuv=pred)
    assert coef.tobytes() == coef2.tobytes() and np.array_equal(rec, rec2)
    assert svc_util.compare_picture(p, coef, rec, None, "GPU resampled layer picture") == (w // 16) * (h // 16)
    st.close(); st2.close()


@pytest.mark.gpu
def test_gpu_layer_picture_resampled_from_another_context():
    """hlb200_svc_layer_picture_resampled_from: the reference layer's reconstruction is read from a frame store of the lower layer's device context (the hand-off
    between the layers of an access unit) -- same coefficients and reconstruction as the host-buffer entry point; with a second GPU in the box the lower
    layer's context lives there and the planes travel GPU to GPU (cudaMemcpyPeerAsync)"""
    from hartallo_b200 import lib as hl
    lib = hl.load()
    base, (rw, rh), pred, (w, h) = golden_pair()
    p = [q for q in svc_util.load_golden() if q["name"] == "g2_3layer.1"][0]
    st = hl.Stream(w, h, 1)
    st.upload_frame(p["src"])
    coef, rec = st.svc_layer_picture_resampled(p["qp"], base, rw, rh)
    for ref_dev in range(min(2, lib.hlb200_device_count())):
        low = hl.Stream(rw, rh, 1, device=ref_dev)      # hlb200_init(ref_dev): the lower layer's context on that GPU
        low.upload_slot(1, base)
        hl.check(lib.hlb200_init(0), "hlb200_init")     # back to the enhancement layer's GPU
        up = hl.Stream(w, h, 1)
        up.upload_frame(p["src"])
        coef2, rec2 = up.svc_layer_picture_resampled_from(p["qp"], low, 1)
        assert coef.tobytes() == coef2.tobytes() and np.array_equal(rec, rec2), ref_dev
        up.close()
        hl.check(lib.hlb200_init(ref_dev), "hlb200_init")
        low.close()
        hl.check(lib.hlb200_init(0), "hlb200_init")
    st.close()
