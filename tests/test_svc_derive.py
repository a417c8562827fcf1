"""SVC inter-layer motion derivation on the device (second half) -- hl_codec_264_utils_derivation_process_initialisation_svc (utils.c:1225: G.8.6.1.1
utils.c:1677, G.8.6.1.2 utils.c:1779, G.8.6.1.3 utils.c:1986) + hl_codec_264_utils_derivation_process_for_mv_comps_and_ref_indices_svc (utils.c:1498), the calls
rdo.c:1318-1346 makes for every enhancement-layer macroblock of a P picture before it predicts.

CPU tier: the oracle restatement (oracle/hl_oracle.c: hlo_svc_derive_mb) and the device source run on the CPU (tools/emu/svc_emu.cpp compiles
hartallo_b200/csrc/hlb_svc_derive.cuh as C++) against
 * the committed golden fixture tests/golden/svc_derive.npz (made by tests/golden/make_golden_svc_derive.py from live runs of the unmodified reference:
   what the derivation read = trace tag 11, what it produced = tag 6),
 * a live multi-layer encode of the reference where oracle/_ref exists (this container), and
 * each other on random reference-layer fields (every partition layout, intra macroblocks, unused partitions, layers of equal and doubled size).
GPU tier: hlb200_dev_svc_derive_motion_batch through the C-ABI against the same fixture and against the oracle on 1080p-size random fields in a picture batch, and
hlb200_svc_layer_picture_derived (derivation + fused prediction / residual coding on the device) against hlb200_svc_layer_picture fed with the reference's field."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import reftrace
import svc_util as S
from oracle_lib import load_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VP = C.c_void_p


def _oracle():
    o = load_oracle()
    o.hlo_svc_derive_picture.restype = None
    o.hlo_svc_derive_picture.argtypes = [VP] + [C.c_int] * 10 + [VP, VP]
    return o


_emu = None


def _emu_lib():
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tools", "emu")
        subprocess.check_call(["make", "-C", d, "libsvc_emu.so"], stdout=subprocess.DEVNULL)
        _emu = C.CDLL(os.path.join(d, "libsvc_emu.so"))
        _emu.svc_emu_derive_motion.restype = C.c_int
        _emu.svc_emu_derive_motion.argtypes = [VP, VP, C.c_int, C.c_int, VP, VP, C.POINTER(C.c_int32)]
    return _emu


def emu_derive(base, geom, w, h, had_parts):
    """the device source on the CPU: (motion, status); had_parts is updated in place"""
    base, geom = np.ascontiguousarray(base), np.ascontiguousarray(geom)
    mot, st = np.zeros((w // 16) * (h // 16), S.MB_MOTION), C.c_int32(0)
    rc = _emu_lib().svc_emu_derive_motion(base.ctypes.data, geom.ctypes.data, w, h, had_parts.ctypes.data, mot.ctypes.data, C.byref(st))
    assert rc == 0, rc
    return mot, st.value


def oracle_derive(o, base, geom, w, h, had_parts):
    out, bad = S.oracle_derive(o, base, geom, w, h)
    return S.motion_from_oracle(out, bad, had_parts)


def _streams(pics):
    """pictures grouped by (stream, layer) in coding order: the per-macroblock 'object holds partitions' flag is carried along each group"""
    groups = {}
    for p in pics:
        groups.setdefault((p.get("stream", ""), p["dqid"]), []).append(p)
    return groups.values()


def _check_pictures(pics, derive, what):
    n = refused = 0
    for group in _streams(pics):
        had = np.zeros((group[0]["w"] // 16) * (group[0]["h"] // 16), np.uint8)
        for p in group:
            m, st = derive(p["base"], p["geom"], p["w"], p["h"], had)
            n += S.compare_derived(p, m, st, what)
            refused += st != 0
    return n, refused


def make_geom(ref_w, ref_h, w, h, level_idc=10):
    """RestrictedSpatialResolutionChangeFlag as layer.c:143 sets it for layers without offsets: equal or doubled size in each direction"""
    g = np.zeros(1, S.SVC_GEOM)
    g["ref_width"], g["ref_height"], g["scaled_width"], g["scaled_height"], g["level_idc"] = ref_w, ref_h, w, h, level_idc
    g["restricted"] = int(w in (ref_w, 2 * ref_w) and h in (ref_h, 2 * ref_h))
    return g


# ------------------------------------------------------------------ CPU tier ------------------------------------------------------------------
def test_golden_fixture_shape():
    pics = S.load_derive_golden()
    assert len(pics) == 30
    assert {p["dqid"] for p in pics} == {16, 32}                                  # first and second enhancement layer
    assert sum(int((p["kind"] == 1).sum()) for p in pics) >= 30                   # macroblocks whose base macroblock is intra
    assert {int(v) for p in pics for v in np.unique(p["part_mode"][p["kind"] == 0])} == {0, 1, 2, 3}
    assert all(int(p["geom"]["cropping_change"][0]) == 0 for p in pics)
    assert sum(int(p["geom"]["restricted"][0]) == 0 for p in pics) == 7           # layers scaled 3:2: the general case with its replacement / merging steps
    assert any(int(p["sub_mode"].max()) > 0 for p in pics if int(p["geom"]["restricted"][0]) == 0)


def test_oracle_vs_golden():
    o = _oracle()
    n, refused = _check_pictures(S.load_derive_golden(), lambda *a: oracle_derive(o, *a), "oracle")
    assert n > 2500 and refused >= 1      # includes a picture the glue refuses (macroblocks without any prediction source)


def test_device_source_vs_golden():
    n, refused = _check_pictures(S.load_derive_golden(), emu_derive, "device source on the CPU")
    assert n > 2500 and refused >= 1


@pytest.mark.skipif(not reftrace.have_driver(), reason="needs oracle/_ref/hl_ref_driver (build container)")
def test_live_reference(tmp_path):
    o = _oracle()
    for name, args in [("g2_3layer", ["--size", "48", "32", "--layers", "3", "--frames", "4", "--gen", "g2", "--seed", "77", "--qp", "33"]),
                       ("g1_2layer", ["--size", "96", "64", "--layers", "2", "--frames", "3", "--gen", "g1", "--seed", "4", "--qp", "28"]),
                       # layers scaled 3:2 (ref_driver --scale): the general case of G.8.6.1 with its replacement / merging steps, sub-macroblock partitions
                       ("g2_ess", ["--size", "64", "192", "--layers", "3", "--frames", "3", "--gen", "g2", "--seed", "5147", "--qp", "44", "--scale", "3", "2"]),
                       ("g1_ess", ["--size", "128", "64", "--layers", "2", "--frames", "4", "--gen", "g1", "--seed", "9", "--qp", "27", "--scale", "3", "2"])]:
        tr = str(tmp_path / (name + ".trace"))
        S.run_driver_svc(args + ["--no-levels"], tr)
        pics = S.derive_pictures_from_trace(tr)
        assert pics
        assert _check_pictures(pics, lambda *a: oracle_derive(o, *a), "oracle")[0] > 0
        assert _check_pictures(pics, emu_derive, "device source on the CPU")[0] > 0


@pytest.mark.parametrize("ref_size,size,level_idc", [((64, 48), (128, 96), 10), ((64, 48), (64, 48), 10), ((80, 32), (160, 32), 10), ((48, 80), (48, 160), 31), ((176, 144), (352, 288), 40),
                                                     ((64, 64), (96, 96), 10), ((160, 96), (240, 144), 40), ((96, 64), (128, 80), 10), ((64, 32), (96, 64), 10)])
def test_device_source_vs_oracle_random(ref_size, size, level_idc):
    """random reference-layer fields: the two restatements agree macroblock by macroblock, over three pictures of a layer (carried 'object holds partitions' flags)"""
    o = _oracle()
    rng = np.random.default_rng(ref_size[0] * 7 + size[1])
    (rw, rh), (w, h) = ref_size, size
    g = make_geom(rw, rh, w, h, level_idc)
    had_o, had_e = np.zeros((w // 16) * (h // 16), np.uint8), np.zeros((w // 16) * (h // 16), np.uint8)
    seen = set()
    for pic in range(3):
        base = S.random_base_field(rng, rw, rh, intra_frac=0.15 if pic else 0.05)
        mo, so = oracle_derive(o, base, g, w, h, had_o)
        me, se = emu_derive(base, g, w, h, had_e)
        assert so == se, (pic, so, se)
        assert np.array_equal(mo.view(np.uint8), me.view(np.uint8)), np.nonzero((mo.view(np.uint8).reshape(len(mo), -1) != me.view(np.uint8).reshape(len(me), -1)).any(axis=1))[0][:8]
        assert np.array_equal(had_o, had_e & 1)
        seen |= {int(v) for v in np.unique(mo["part_mode"])} | {10 + int(v) for v in np.unique(mo["sub_mode"])}
    if size == ref_size or not int(g["restricted"][0]):
        assert {0, 3} <= seen and seen & {11, 12, 13}      # equal size / general case: sub-macroblock partitions appear
    else:
        assert {0, 3} <= seen


def test_refusals():
    """geometry the derivation is not pinned for is refused, not approximated"""
    base = S.random_base_field(np.random.default_rng(1), 64, 48)
    for field, value in (("cropping_change", 1),):
        g = make_geom(64, 48, 128, 96)
        g[field] = value
        mot, st = np.zeros(48, S.MB_MOTION), C.c_int32(0)
        had = np.zeros(48, np.uint8)
        assert _emu_lib().svc_emu_derive_motion(base.ctypes.data, g.ctypes.data, 128, 96, had.ctypes.data, mot.ctypes.data, C.byref(st)) == 7
    g = make_geom(64, 48, 128, 96, level_idc=40)      # level_idc > 30 with a power-of-two reference width: the reference's own fixed-point set-up overflows
    assert _emu_lib().svc_emu_derive_motion(base.ctypes.data, g.ctypes.data, 128, 96, had.ctypes.data, mot.ctypes.data, C.byref(st)) == 7


# ------------------------------------------------------------------ GPU tier ------------------------------------------------------------------
def _gpu_derive(lib, torch, base_list, geom, w, h, had_t):
    """hlb200_dev_svc_derive_motion_batch on len(base_list) pictures: (motion arrays, status per picture); had_t = device tensor carried by the caller"""
    from hartallo_b200 import lib as hl
    n, nmb = len(base_list), (w // 16) * (h // 16)
    base = torch.from_numpy(np.concatenate(base_list).view(np.uint8).copy()).cuda()
    motion = torch.zeros(n * nmb * hl.MB_MOTION.itemsize, dtype=torch.uint8, device="cuda")
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    g = np.ascontiguousarray(geom)
    hl.check(lib.hlb200_dev_svc_derive_motion_batch(base.data_ptr(), g.ctypes.data, w, h, n, had_t.data_ptr(), motion.data_ptr(), status.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "hlb200_dev_svc_derive_motion_batch")
    torch.cuda.synchronize()
    return motion.cpu().numpy().view(hl.MB_MOTION).reshape(n, nmb), status.cpu().numpy()


@pytest.mark.gpu
def test_gpu_vs_golden():
    import torch
    from hartallo_b200 import lib as hl
    lib = hl.load()
    n = refused = 0
    for group in _streams(S.load_derive_golden()):
        w, h = group[0]["w"], group[0]["h"]
        had = torch.zeros((w // 16) * (h // 16), dtype=torch.uint8, device="cuda")
        for p in group:
            m, st = _gpu_derive(lib, torch, [p["base"]], p["geom"], w, h, had)
            n += S.compare_derived(p, m[0], int(st[0]), "GPU")
            refused += int(st[0]) != 0
    assert n > 2500 and refused >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("ref_size,size", [((960, 544), (1920, 1088)), ((1920, 1088), (1920, 1088)), ((176, 144), (352, 288)), ((1280, 704), (1920, 1056))])
def test_gpu_vs_oracle_random_batch(ref_size, size):
    """full-size random fields, eight pictures in one launch, two launches (the carried flags of picture k feed picture k of the next launch)"""
    import torch
    from hartallo_b200 import lib as hl
    lib, o = hl.load(), _oracle()
    (rw, rh), (w, h) = ref_size, size
    rng = np.random.default_rng(rw + w)
    g, nmb, npic = make_geom(rw, rh, w, h, 40 if rw == 960 else 10), (w // 16) * (h // 16), 8
    had_t = torch.zeros(npic * nmb, dtype=torch.uint8, device="cuda")
    had_o = np.zeros((npic, nmb), np.uint8)
    for launch in range(2):
        bases = [S.random_base_field(rng, rw, rh, intra_frac=0.02 + 0.05 * launch) for _ in range(npic)]
        m, st = _gpu_derive(lib, torch, bases, g, w, h, had_t)
        for k in range(npic):
            mo, so = oracle_derive(o, bases[k], g, w, h, had_o[k])
            assert int(st[k]) == so, (launch, k, int(st[k]), so)
            assert np.array_equal(mo.view(np.uint8), m[k].view(np.uint8)), (launch, k)
        assert np.array_equal(had_t.cpu().numpy().reshape(npic, nmb) & 1, had_o)


@pytest.mark.gpu
def test_gpu_layer_picture_derived_equals_host_field():
    """derivation + coding on the device (hlb200_svc_layer_picture_derived) = coding with the field the reference derived (hlb200_svc_layer_picture), picture by picture
    along the golden streams of tests/golden/svc_inter.npz whose derivation inputs are in tests/golden/svc_derive.npz"""
    from hartallo_b200 import lib as hl
    inter = {(p["name"].rsplit(".", 1)[0], p["frame"], p["dqid"]): p for p in S.load_golden() if p["kind"] == 0}
    done = 0
    for group in _streams(S.load_derive_golden()):
        w, h = group[0]["w"], group[0]["h"]
        a, b = hl.Stream(w, h), hl.Stream(w, h)
        for d in group:
            p = inter.get((d["stream"], d["frame"], d["dqid"]))
            if p is None:
                continue
            for s in (a, b):
                s.upload_frame(p["src"]); s.upload_slot(0, p["ref"])
            c1, r1, m1, st = a.svc_layer_picture_derived(p["qp"], d["base"], d["geom"])
            if st:
                assert (d["valid"] == 0).any()
                continue
            c0, r0 = b.svc_layer_picture(p["qp"], motion=p["motion"])
            assert np.array_equal(c0.view(np.uint8), c1.view(np.uint8)) and np.array_equal(r0, r1), d["name"]
            S.compare_picture(p, c1, r1, None, "derived on the device")
            done += 1
        a.close(); b.close()
    assert done >= 6
