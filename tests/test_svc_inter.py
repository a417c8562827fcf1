"""SVC enhancement-layer inter macroblocks (base_mode_flag = 1) -- SURVEY 8a row a14, hl_codec_264_rdo_mb_guess_best_inter_pred_svc (rdo.c:1273-1521).

CPU tier: the oracle restatement (oracle/hl_oracle.c: hlo_recon_svc_inter_mb, prediction by hlo_interp_luma/chroma) and the device source of the
fused kernel run lane by lane on the CPU (tools/emu/svc_emu.cpp compiles hartallo_b200/csrc/hlb_svc.cuh as C++) against
 * the committed golden fixture tests/golden/svc_inter.npz (made by tests/golden/make_golden_svc.py from the unmodified reference), and
 * a live multi-layer encode of the reference where oracle/_ref exists (this container).
GPU tier: hlb200_dev_svc_inter_recon_batch through the C-ABI against the same fixture, against the oracle on random pictures / motion fields
(all partition shapes, vectors far outside the picture, random carried state), and a picture batch in one launch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import reftrace
import svc_util
from oracle_lib import chroma_qp, i16p, i32p, load_oracle_mb, oracle_predict_frame, u8p
from svc_util import MB_COEFFS, MB_MOTION, SVC_STATE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle():
    o = load_oracle_mb()
    o.hlo_recon_svc_inter_mb.restype = None
    o.hlo_recon_svc_inter_mb.argtypes = [u8p] * 6 + [C.c_int] * 5 + [i16p, i16p, i16p, i32p, i32p, i32p, u8p, u8p, u8p]
    o.hlo_recon_svc_bl_mb.restype = None
    o.hlo_recon_svc_bl_mb.argtypes = o.hlo_recon_svc_inter_mb.argtypes
    return o


def oracle_picture(o, src, ref, w, h, qp, qpc, motion, state_in, bl=0):
    """whole picture through the oracle: returns (coeffs, rec_yuv, state_out); bl: `ref` holds the prediction planes (I_BL, rdo.c:301: same residual coding)"""
    ysz, csz = w * h, w * h // 4
    nmb = (w // 16) * (h // 16)
    inherit = {}
    if bl:
        py, pu, pv = (np.ascontiguousarray(ref[:ysz]), np.ascontiguousarray(ref[ysz:ysz + csz]), np.ascontiguousarray(ref[ysz + csz:]))
    else:
        py, pu, pv = oracle_predict_frame(o, ref, w, h, motion)
        # macroblocks that inherit the prediction of an earlier macroblock of the picture (svc_util.mark_inherited): copy it over, code them as I_BL
        inherit = {a: int(m["pad"][1]) | (int(m["pad"][2]) << 8) for a, m in enumerate(motion) if m["pad"][0] & 1}
        if inherit:
            mbs = svc_util.mb_of_planes(py.reshape(-1), pu.reshape(-1), pv.reshape(-1), w, h)
            for a, b in inherit.items():
                assert not (motion[b]["pad"][0] & 1)
                mbs[a] = mbs[b]
            pl = svc_util.planes_of_mb(mbs, w, h)
            py, pu, pv = np.ascontiguousarray(pl[:ysz]), np.ascontiguousarray(pl[ysz:ysz + csz]), np.ascontiguousarray(pl[ysz + csz:])
    sy, su, sv = (np.ascontiguousarray(src[:ysz]), np.ascontiguousarray(src[ysz:ysz + csz]), np.ascontiguousarray(src[ysz + csz:]))
    ry, ru, rv = np.zeros(ysz, np.uint8), np.zeros(csz, np.uint8), np.zeros(csz, np.uint8)
    coeffs, state = np.zeros(nmb, MB_COEFFS), state_in.copy()
    c4, cdc, cac = np.zeros(1, np.int32), np.zeros(2, np.int32), np.zeros(2, np.int32)
    for mb in range(nmb):
        ll, dc = np.zeros(256, np.int16), np.ascontiguousarray(state[mb]["chroma_dc_level"].reshape(-1))
        ac = np.ascontiguousarray(state[mb]["chroma_ac_level"].reshape(-1))
        (o.hlo_recon_svc_bl_mb if (bl or mb in inherit) else o.hlo_recon_svc_inter_mb)(sy, su, sv, py.reshape(-1), pu.reshape(-1), pv.reshape(-1), w, mb % (w // 16), mb // (w // 16), qp, qpc, ll, dc, ac, c4, cdc, cac, ry, ru, rv)
        c = coeffs[mb]
        c["luma_level"], c["chroma_dc_level"], c["chroma_ac_level"] = ll.reshape(16, 16), dc.reshape(2, 4), ac.reshape(2, 4, 16)
        c["cbp_luma4x4"], c["cbp_chroma_dc4x4"], c["cbp_chroma_ac4x4"] = int(c4[0]), cdc, cac
        state[mb]["chroma_ac_level"], state[mb]["chroma_dc_level"] = ac.reshape(2, 4, 16), dc.reshape(2, 4)
    return coeffs, np.concatenate([ry, ru, rv]), state


_emu = None


def _emu_lib():
    """tools/emu/libsvc_emu.so: hartallo_b200/csrc/hlb_svc.cuh compiled as plain C++ (built on demand; g++ only)"""
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tools", "emu")
        subprocess.check_call(["make", "-C", d, "libsvc_emu.so"], stdout=subprocess.DEVNULL)
        _emu = C.CDLL(os.path.join(d, "libsvc_emu.so"))
        _emu.svc_emu_recon_batch.restype = C.c_int
        _emu.svc_emu_recon_batch.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int] + [C.c_void_p] * 6
    return _emu


def emu_picture(src, ref, w, h, qp, motion, state_in, bl=0):
    ysz, csz = w * h, w * h // 4
    src, ref, motion = np.ascontiguousarray(src), np.ascontiguousarray(ref), np.ascontiguousarray(motion)
    rec, state, coeffs = np.zeros_like(src), state_in.copy(), np.zeros(len(motion), MB_COEFFS)
    s, r, o = src.ctypes.data, ref.ctypes.data, rec.ctypes.data
    rc = _emu_lib().svc_emu_recon_batch(bl, s, s + ysz, s + ysz + csz, r, r + ysz, r + ysz + csz, w, h, 1, 0, qp, 0, motion.ctypes.data, state.ctypes.data,
                                              coeffs.ctypes.data, o, o + ysz, o + ysz + csz)
    assert rc == 0
    return coeffs, rec, state


def _fill_invalid(p):
    """macroblocks without reference behaviour get zero motion (any valid motion will do; they are not compared)"""
    return p["motion"].copy()


# ------------------------------------------------------------------ CPU tier ------------------------------------------------------------------
def test_golden_fixture_shape():
    pics = svc_util.load_golden()
    assert len(pics) == 13 and sum(p["kind"] for p in pics) == 4     # 9 P pictures (base-mode inter) + 4 I pictures (I_BL)
    modes = set()
    for p in pics:
        assert p["qp"] in (24, 31, 36) and len(p["motion"]) == (p["w"] // 16) * (p["h"] // 16)
        modes |= set(int(m) for m, v in zip(p["motion"]["part_mode"], p["valid"]) if v and not p["kind"])
    assert modes == {0, 1, 2, 3}                                    # 16x16, 16x8, 8x16, 8x8 all occur
    assert sum(int((p["motion"]["pad"][:, 0] & 1).sum()) for p in pics) >= 8   # and macroblocks without partitions that inherit an earlier prediction
    assert any(p["dqid"] == 32 for p in pics)                       # third spatial layer


def test_oracle_vs_golden():
    o = _oracle()
    n = 0
    for p in svc_util.load_golden():
        coeffs, rec, state = oracle_picture(o, p["src"], p["ref"], p["w"], p["h"], p["qp"], chroma_qp(p["qp"]), _fill_invalid(p), p["state_in"], p["kind"])
        n += svc_util.compare_picture(p, coeffs, rec, state, "oracle")
    assert n >= 540


def test_device_source_on_cpu_vs_golden():
    n = 0
    for p in svc_util.load_golden():
        coeffs, rec, state = emu_picture(p["src"], p["ref"], p["w"], p["h"], p["qp"], _fill_invalid(p), p["state_in"], p["kind"])
        n += svc_util.compare_picture(p, coeffs, rec, state, "hlb_svc.cuh on the CPU")
    assert n >= 540


@pytest.mark.skipif(not reftrace.have_driver(), reason="oracle/_ref/hl_ref_driver only exists where the reference tree is available")
@pytest.mark.parametrize("args", [["--size", "176", "144", "--layers", "2", "--frames", "3", "--gen", "g1"],
                                  ["--size", "80", "64", "--layers", "3", "--frames", "3", "--gen", "g2", "--seed", "11", "--qp", "28"],
                                  ["--size", "96", "80", "--layers", "2", "--frames", "4", "--gen", "g2", "--seed", "4", "--qp", "40"]])
def test_live_reference(tmp_path, args):
    """oracle and device source against a fresh multi-layer encode of the reference (QCIF -> CIF is the lower half of BASELINE.json configs[3])"""
    tr = str(tmp_path / "svc.trace")
    svc_util.run_driver_svc(args, tr)
    o = _oracle()
    n = 0
    for p in svc_util.bl_pictures_from_trace(tr) + svc_util.pictures_from_trace(tr):
        c1, r1, s1 = oracle_picture(o, p["src"], p["ref"], p["w"], p["h"], p["qp"], p["qpc"], p["motion"], p["state_in"], p["kind"])
        n += svc_util.compare_picture(p, c1, r1, s1, "oracle")
        c2, r2, s2 = emu_picture(p["src"], p["ref"], p["w"], p["h"], p["qp"], p["motion"], p["state_in"], p["kind"])
        svc_util.compare_picture(p, c2, r2, s2, "hlb_svc.cuh on the CPU")
    assert n > 100


@pytest.mark.skipif(not reftrace.have_driver(), reason="oracle/_ref/hl_ref_driver only exists where the reference tree is available")
def test_svc_fuzz_vs_live_reference():
    """a few random multi-layer configurations of tools/emu/fuzz_svc.py (oracle + device source against the reference run live; bitstream MD5 through the
    glue hook on G1 content)"""
    out = subprocess.run([os.sys.executable, os.path.join(ROOT, "tools", "emu", "fuzz_svc.py"), "6", "500"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and ", 0 mismatches" in out.stdout, out.stdout[-800:]


@pytest.mark.skipif(not reftrace.have_driver(), reason="oracle/_ref/hl_ref_driver only exists where the reference tree is available")
def test_stale_prediction_model(tmp_path):
    """what the reference's scratch blocks really hold when it codes a macroblock without partitions (trace tag 10) == the prediction of the last macroblock
    with partitions, as svc_util.mark_inherited / SvcPredSrc model it"""
    tr = str(tmp_path / "svc.trace")
    svc_util.run_driver_svc(["--size", "176", "144", "--layers", "2", "--frames", "4", "--gen", "g2"], tr)
    o = _oracle()
    n = 0
    for p in svc_util.pictures_from_trace(tr):
        if not p["stale"]:
            continue
        w, h = p["w"], p["h"]
        plain = p["motion"].copy()
        plain["pad"] = 0
        py, pu, pv = oracle_predict_frame(o, p["ref"], w, h, plain)
        mbs = svc_util.mb_of_planes(py.reshape(-1), pu.reshape(-1), pv.reshape(-1), w, h)
        for a, held in p["stale"].items():
            m = p["motion"][a]
            if m["pad"][0] & 1:
                assert np.array_equal(held, mbs[int(m["pad"][1]) | (int(m["pad"][2]) << 8)]), a
                n += 1
    assert n >= 8


def _random_case(rng, w, h, far_every=0):
    from test_oracle_pinned import stress_plane
    nmb = (w // 16) * (h // 16)

    def frame():
        return np.concatenate([stress_plane(rng, h, w).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1)])
    ref = frame()
    # source = reference moved a little + noise, so that zero / small / large residuals all occur
    src = ref.copy()
    ysz = w * h
    src[:ysz] = np.roll(ref[:ysz].reshape(h, w), (1, 2), (0, 1)).reshape(-1)
    noise = rng.integers(-3, 4, src.shape)
    mask = rng.random(src.shape) < 0.3
    src = np.clip(src.astype(np.int32) + noise * mask, 0, 255).astype(np.uint8)
    flat = rng.integers(0, nmb, max(1, nmb // 6))             # some macroblocks identical to the reference (zero residual -> stale state is read)
    mbw = w // 16
    m = np.zeros(nmb, MB_MOTION)
    m["part_mode"] = rng.integers(0, 4, nmb)
    m["sub_mode"] = rng.integers(0, 4, (nmb, 4))
    m["mv"] = rng.integers(-40, 41, (nmb, 4, 4, 2))
    if far_every:
        idx = np.arange(0, nmb, far_every)
        m["mv"][idx] = rng.integers(-9000, 9001, (len(idx), 4, 4, 2))
    for a in flat:
        m["part_mode"][a] = 0
        m["mv"][a] = 0
        x, y = (a % mbw) * 16, (a // mbw) * 16
        sy, ry_ = src[:ysz].reshape(h, w), ref[:ysz].reshape(h, w)
        sy[y:y + 16, x:x + 16] = ry_[y:y + 16, x:x + 16]
        for pl in range(2):
            o = ysz + pl * (ysz // 4)
            sc, rc = src[o:o + ysz // 4].reshape(h // 2, w // 2), ref[o:o + ysz // 4].reshape(h // 2, w // 2)
            sc[y // 2:y // 2 + 8, x // 2:x // 2 + 8] = rc[y // 2:y // 2 + 8, x // 2:x // 2 + 8]
            if rng.random() < 0.5:
                sc[y // 2:y // 2 + 4, x // 2:x // 2 + 4] += 9          # one block with a DC-only difference: its three neighbours read their stale AC levels
    sources = set()
    for a in rng.integers(1, nmb, max(1, nmb // 8)):           # macroblocks that inherit the prediction of an earlier one (never of another inheriting one)
        b = int(rng.integers(0, a))
        if not (m["pad"][b][0] & 1) and int(a) not in sources:   # (a macroblock somebody inherits from must keep its own partitions)
            m["pad"][a] = (1, b & 255, b >> 8)
            sources.add(b)
    st = np.zeros(nmb, SVC_STATE)
    st["chroma_ac_level"][:, :, :, :15] = rng.integers(-2, 3, (nmb, 2, 4, 15)) * (rng.random((nmb, 2, 4, 15)) < 0.2)
    st["chroma_dc_level"] = rng.integers(-3, 4, (nmb, 2, 4))
    return src, ref, m, st


@pytest.mark.parametrize("w,h,qp,far,bl", [(64, 48, 28, 0, 0), (48, 64, 12, 3, 0), (80, 32, 44, 0, 0), (64, 64, 30, 0, 1)])
def test_device_source_on_cpu_vs_oracle_random(w, h, qp, far, bl):
    """all partition shapes (also the sub-8x8 ones the dyadic inter-layer derivation never produces), out-of-picture vectors, carried state"""
    rng = np.random.default_rng(w * 131 + qp)
    src, ref, m, st = _random_case(rng, w, h, far)
    c1, r1, s1 = oracle_picture(_oracle(), src, ref, w, h, qp, chroma_qp(qp), m, st, bl)
    c2, r2, s2 = emu_picture(src, ref, w, h, qp, m, st, bl)
    assert np.array_equal(r1, r2)
    assert c1.tobytes() == c2.tobytes()
    assert s1.tobytes() == s2.tobytes()
    assert (c1["cbp_luma4x4"] != 0).any() and (c1["cbp_luma4x4"] != 0xffff).any()


# ------------------------------------------------------------------ GPU tier ------------------------------------------------------------------
def gpu_pictures(pics_in, size, qp, bl=0):
    """pics_in: list of (src, ref, motion, state) of ONE size (w, h) -> one launch of hlb200_dev_svc_inter_recon_batch / _bl_recon_batch over all of them"""
    import torch
    from hartallo_b200 import lib as hl
    lib = hl.load()
    assert hl.MB_COEFFS.itemsize == MB_COEFFS.itemsize and hl.MB_MOTION.itemsize == MB_MOTION.itemsize and hl.SVC_STATE.itemsize == SVC_STATE.itemsize
    n = len(pics_in)
    dev = torch.device("cuda:0")
    srcs, refs = np.stack([p[0] for p in pics_in]), np.stack([p[1] for p in pics_in])
    motion, state = np.concatenate([p[2] for p in pics_in]), np.concatenate([p[3] for p in pics_in])
    w, h = size
    ysz, csz = w * h, w * h // 4
    fb = ysz + 2 * csz
    nmb = (w // 16) * (h // 16)
    d_src, d_ref = torch.from_numpy(srcs).to(dev), torch.from_numpy(refs).to(dev)
    d_rec = torch.zeros_like(d_src)
    d_motion = torch.from_numpy(np.ascontiguousarray(motion).view(np.uint8)).to(dev)
    d_state = torch.from_numpy(np.ascontiguousarray(state).view(np.uint8).copy()).to(dev)
    d_coef = torch.zeros(n * nmb * MB_COEFFS.itemsize, dtype=torch.uint8, device=dev)
    s, r, o = d_src.data_ptr(), d_ref.data_ptr(), d_rec.data_ptr()
    sp = torch.cuda.current_stream().cuda_stream
    if bl:
        hl.check(lib.hlb200_dev_svc_bl_recon_batch(s, s + ysz, s + ysz + csz, r, r + ysz, r + ysz + csz, w, h, n, fb, qp, 0, d_state.data_ptr(), d_coef.data_ptr(),
                                                   o, o + ysz, o + ysz + csz, sp), "svc_bl_recon_batch")
    else:
        hl.check(lib.hlb200_dev_svc_inter_recon_batch(s, s + ysz, s + ysz + csz, r, r + ysz, r + ysz + csz, w, h, n, fb, qp, 0, d_motion.data_ptr(), d_state.data_ptr(),
                                                      d_coef.data_ptr(), o, o + ysz, o + ysz + csz, sp), "svc_inter_recon_batch")
    torch.cuda.synchronize()
    coef = d_coef.cpu().numpy().view(MB_COEFFS).reshape(n, nmb)
    st = d_state.cpu().numpy().view(SVC_STATE).reshape(n, nmb)
    return coef, d_rec.cpu().numpy(), st


@pytest.mark.gpu
def test_gpu_vs_golden():
    n = 0
    for p in svc_util.load_golden():
        coef, rec, st = gpu_pictures([(p["src"], p["ref"], _fill_invalid(p), p["state_in"])], (p["w"], p["h"]), p["qp"], p["kind"])
        n += svc_util.compare_picture(p, coef[0], rec[0], st[0], "GPU")
    assert n >= 540


@pytest.mark.gpu
def test_gpu_layer_picture_host_api_vs_golden():
    """hlb200_svc_layer_picture (host buffers, one context per layer, the per-macroblock state carried inside the context from picture to picture): the
    pictures of each fixture encode in coding order, every layer in its own context -- the call sequence of host/hlb200_glue.c"""
    from hartallo_b200 import lib as hl
    ctxs, n = {}, 0
    for p in svc_util.load_golden():
        key = (p["name"].split(".")[0], p["dqid"])
        if key not in ctxs:
            ctxs[key] = hl.Stream(p["w"], p["h"], 1)
        st = ctxs[key]
        st.upload_frame(p["src"])
        if p["kind"]:
            coef, rec = st.svc_layer_picture(p["qp"], pred_yuv=p["ref"])
        else:
            st.upload_slot(0, p["ref"])
            coef, rec = st.svc_layer_picture(p["qp"], motion=_fill_invalid(p).view(hl.MB_MOTION))
        # the context's carried state equals the reference's only while every macroblock so far had reference behaviour; the fixture's state_in is what the
        # reference really held, so compare just the pictures for which both agree (all of them except after a picture with host-coded macroblocks)
        if getattr(st, "svc_state_ok", True):
            n += svc_util.compare_picture(p, coef, rec, None, "GPU host API")
        st.svc_state_ok = getattr(st, "svc_state_ok", True) and bool(p["valid"].all())
    for st in ctxs.values():
        st.close()
    assert n >= 500


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,qp,n,bl", [(64, 48, 28, 1, 0), (176, 144, 31, 5, 0), (48, 64, 12, 3, 0), (1920, 1088, 31, 1, 0), (176, 144, 26, 3, 1), (1920, 1088, 31, 2, 1)])
def test_gpu_vs_oracle_random(w, h, qp, n, bl):
    """bl = 1: the second plane set is used as the prediction itself (I_BL entry point)"""
    rng = np.random.default_rng(w + 7 * qp + n)
    cases = [_random_case(rng, w, h, far_every=3 if i == 1 or n == 1 else 0) for i in range(n)]
    coef, rec, st = gpu_pictures(cases, (w, h), qp, bl)
    o = _oracle()
    for i, (src, ref, m, s0) in enumerate(cases):
        if w >= 1920:   # full size: the oracle on a band of macroblock rows is enough for a picture whose rows are independent; the rest against the CPU run of the device source
            c2, r2, s2 = emu_picture(src, ref, w, h, qp, m, s0, bl)
            assert np.array_equal(rec[i], r2) and coef[i].tobytes() == c2.tobytes() and st[i].tobytes() == s2.tobytes()
            continue
        c1, r1, s1 = oracle_picture(o, src, ref, w, h, qp, chroma_qp(qp), m, s0, bl)
        assert np.array_equal(rec[i], r1), i
        assert coef[i].tobytes() == c1.tobytes(), i
        assert st[i].tobytes() == s1.tobytes(), i


# ------------------------------------------------------ the drop-in hook for enhancement layers ------------------------------------------------------
import json  # noqa: E402

SVC_STREAMS = json.load(open(os.path.join(ROOT, "tests", "golden", "svc_bitstream.json")))
GLUE_CHECK = os.path.join(ROOT, "oracle", "_ref", "hl_svc_glue_check")
GLUE_FULL = os.path.join(ROOT, "oracle", "_ref", "hl_glue_check_full")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "hl_ref_driver")
B200_ENCODER = os.path.join(ROOT, "oracle", "_ref", "hl_b200_encoder")


def _encode(exe, args):
    out = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr[-800:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(GLUE_CHECK), reason="oracle/_ref/hl_svc_glue_check only exists where the reference tree is available")
@pytest.mark.parametrize("name", sorted(SVC_STREAMS))
def test_svc_glue_hook_bitstream_md5_cpu(name):
    """host/hlb200_glue.c's enhancement-layer hook (derivation pre-pass -> ONE picture call -> the reference's own loop and CAVLC writer) with the device source
    compiled as C++ standing in for the library (tools/emu/svc_shim.cpp), base layer on the reference's CPU path: the multi-layer bitstream must equal the
    all-reference one byte for byte (QCIF -> CIF -> 4CIF is BASELINE.json configs[3])"""
    g = SVC_STREAMS[name]
    got = _encode(GLUE_CHECK, g["args"])
    assert (got["bytes"], got["md5"]) == (g["bytes"], g["md5"]), (got, g)


@pytest.mark.skipif(not os.path.exists(GLUE_FULL), reason="oracle/_ref/hl_glue_check_full only exists where the reference tree is available")
@pytest.mark.parametrize("name", sorted(SVC_STREAMS))
def test_whole_glue_bitstream_md5_cpu(name):
    """the WHOLE glue object that is linked into hl_b200_encoder (base-layer hook + enhancement-layer hook), with the per-macroblock sources of both kernels
    compiled as C++ standing in for the library (tools/emu/svc_shim.cpp -DSVC_SHIM_WITH_SLICE): no layer goes through the reference's decision functions, and
    the multi-layer bitstream still equals the all-reference one.  This is the CPU-tier twin of test_svc_bitstream_md5_drop_in."""
    g = SVC_STREAMS[name]
    got = _encode(GLUE_FULL, g["args"])
    assert (got["bytes"], got["md5"]) == (g["bytes"], g["md5"]), (got, g)


@pytest.mark.skipif(not os.path.exists(GLUE_FULL), reason="oracle/_ref/hl_glue_check_full only exists where the reference tree is available")
def test_single_layer_glue_bitstream_md5_cpu():
    """CPU-tier twin of test_encoder.py::test_bitstream_md5_drop_in: the AVC drop-in (slice hook + decision records copied into the reference's macroblock
    objects + the reference's own CAVLC writer) with the slice kernel's source standing in for the library, on every golden encoder configuration"""
    import glob
    n = 0
    for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "encoder_*.npz"))):
        g = np.load(f)
        w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
        if w * h * frames > 352 * 288 * 4:     # keep the CPU tier short: the big ones run on the GPU
            continue
        refs = int(g["refs"]) if "refs" in g.files else 1
        early, deblock = (int(g["flags"][0]), int(g["flags"][1])) if "flags" in g.files else (0, 0)
        got = _encode(GLUE_FULL, ["--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp), "--me-range", str(me_range), "--refs", str(refs), "--gen", str(g["gen"]),
                                  "--seed", str(seed), "--early-term", str(early), "--deblock", str(deblock)])
        assert got["md5"] == str(g["bitstream_md5"]), (f, got)
        n += 1
    assert n >= 3


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200_ENCODER), reason="oracle/_ref/hl_b200_encoder not built (needs the reference tree at build time)")
@pytest.mark.parametrize("name", sorted(SVC_STREAMS))
def test_svc_bitstream_md5_drop_in(name):
    """the same through the real library: every layer on the B200 (base layer: slice kernel; enhancement layers: k_svc_inter_recon), the reference's
    unmodified host code around it"""
    g = SVC_STREAMS[name]
    got = _encode(B200_ENCODER, g["args"])
    assert (got["bytes"], got["md5"]) == (g["bytes"], g["md5"]), (got, g)


# ------------------------------------------------------ what the drop-in refuses ------------------------------------------------------
def _run_expect_refusal(exe, args, needle):
    out = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode != 0, "the drop-in produced a stream where it must refuse: " + out.stdout[-200:]
    assert "not implemented" in out.stderr.lower() and needle in out.stderr, out.stderr[-600:]


@pytest.mark.skipif(not os.path.exists(GLUE_FULL), reason="oracle/_ref/hl_glue_check_full only exists where the reference tree is available")
@pytest.mark.parametrize("args,needle", [
    (["--size", "48", "48", "--layers", "2", "--frames", "2", "--deblock", "1"], "deblock_flag"),   # inter-layer deblocking (deblock.c:175-186) is not reproduced
    (["--size", "48", "16", "--layers", "3", "--frames", "2", "--gen", "g2", "--seed", "1865", "--qp", "22"], "Intra_Base"),   # 12-macroblock enhancement I picture (layer.c:202)
    (["--size", "48", "48", "--layers", "3", "--frames", "4", "--gen", "g2", "--seed", "21", "--qp", "30"], "status bits 8"),   # HLB200_SVC_DERIVE_NO_PRED_SOURCE: coded against an earlier picture's scratch memory
])
def test_glue_refuses_what_it_does_not_reproduce(args, needle):
    """No silent divergence and no CPU fallback: deblocking of streams with SVC layers, enhancement-layer I pictures too small for the reference's own window
    array and macroblocks the reference codes against stale scratch memory all end in HL_ERROR_NOT_IMPLEMENTED (host/hlb200_glue.c)."""
    _run_expect_refusal(GLUE_FULL, args, needle)


@pytest.mark.skipif(not os.path.exists(GLUE_FULL) or not os.path.exists(REF_DRIVER), reason="needs oracle/_ref (built where the reference tree is available)")
@pytest.mark.parametrize("args", [
    ["--size", "176", "144", "--frames", "4", "--gen", "g2", "--seed", "8", "--qp", "33", "--defaults"],
    ["--size", "96", "80", "--frames", "5", "--gen", "g3", "--seed", "2", "--qp", "22", "--defaults"],
    ["--size", "64", "48", "--frames", "3", "--gen", "g1", "--defaults"],
])
def test_library_defaults_drop_in_cpu(args):
    """hl_codec_create's own settings (deblock_flag = 1, me_early_term_flag = 1, hl_types.h:67,69) through the drop-in give the reference's stream byte for byte
    (CPU tier: the kernels' per-macroblock source stands in for the library; the GPU twin is test_encoder.py::test_bitstream_md5_drop_in on the *_defaults goldens)"""
    ref = _encode(REF_DRIVER, args)
    got = _encode(GLUE_FULL, args)
    assert (got["bytes"], got["md5"]) == (ref["bytes"], ref["md5"]), (got, ref)


def test_glue_has_no_reference_cpu_path():
    """the product glue neither calls the reference's own SVC decision function nor its CPU resampling (round-1 finding)"""
    src = open(os.path.join(ROOT, "host", "hlb200_glue.c")).read()
    assert "__real_hl_codec_264_rdo_mb_guess_best_inter_pred_svc" not in src
    assert "_hl_codec_264_decode_svc_resample_intra_colour_comps(" not in src
