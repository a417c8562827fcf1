"""SVC enhancement-layer inter macroblocks: the reference's per-macroblock trace (tags 6 / 7 of oracle/ref_driver.c) as arrays in the
layouts of include/hlb200.h, the committed golden fixture made from it, and the comparison shared by the oracle / CPU-emulation / GPU tests."""
import os
import subprocess

import numpy as np

import reftrace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "svc_inter.npz")

# numpy mirrors of hlb200_mb_motion_t / hlb200_mb_coeffs_t / hlb200_svc_mb_state_t (kept here so that the CPU tier does not import the CUDA binding)
MB_MOTION = np.dtype([("part_mode", "u1"), ("sub_mode", "u1", (4,)), ("ref_idx", "i1", (4,)), ("pad", "u1", (3,)), ("mv", "<i2", (4, 4, 2))])
MB_COEFFS = np.dtype([("luma_level", "<i2", (16, 16)), ("chroma_dc_level", "<i2", (2, 4)), ("chroma_ac_level", "<i2", (2, 4, 16)),
                      ("cbp_luma4x4", "<u2"), ("cbp_chroma_dc4x4", "u1", (2,)), ("cbp_chroma_ac4x4", "u1", (2,)), ("pad", "u1", (2,))])
SVC_STATE = np.dtype([("chroma_ac_level", "<i2", (2, 4, 16)), ("chroma_dc_level", "<i2", (2, 4))])

# (name, driver arguments): small multi-layer encodes whose enhancement P pictures make the fixture
CONFIGS = [
    ("g2_2layer", ["--size", "64", "48", "--layers", "2", "--frames", "4", "--gen", "g2", "--seed", "3"]),
    ("g2_3layer", ["--size", "32", "32", "--layers", "3", "--frames", "3", "--gen", "g2", "--seed", "9", "--qp", "36"]),
    ("g1_2layer_q24", ["--size", "48", "48", "--layers", "2", "--frames", "3", "--gen", "g1", "--qp", "24"]),
]


def run_driver_svc(args, trace_path):
    subprocess.run([reftrace.DRIVER] + args + ["--trace", trace_path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)


def pictures_from_trace(path):
    """list of dicts, one per enhancement-layer P picture of the trace, in the layouts of the C-ABI"""
    t = reftrace.parse(path)
    pics, order = {}, []
    for r in t.get(7, []):
        W, H = int(r[4]), int(r[5])
        n = W * H * 3 // 2
        b = r[6:].view(np.uint8)
        nmb = (W // 16) * (H // 16)
        key = (int(r[2]), int(r[3]))
        order.append(key)
        pics[key] = dict(frame=key[0], dqid=key[1], w=W, h=H, qp=-1, src=b[:n].copy(), ref=b[n:2 * n].copy(), motion=np.zeros(nmb, MB_MOTION),
                         valid=np.zeros(nmb, np.uint8), state_in=np.zeros(nmb, SVC_STATE), expect=np.zeros(nmb, MB_COEFFS), rec=np.zeros((nmb, 384), np.uint8), seen=0)
    for r in t.get(6, []):
        p = pics[(int(r[2]), int(r[3]))]
        a = int(r[4])
        qp, qpc = int(r[5]), int(r[6])
        assert p["qp"] in (-1, qp) and int(r[7]) == qpc
        p["qp"], p["qpc"] = qp, qpc
        nparts, pw, ph = int(r[8]), int(r[9]), int(r[10])
        nsub, pflag, ridx, mv = r[11:15], r[23:27], r[27:31], r[31:63].reshape(4, 4, 2)
        k = 63
        e = p["expect"][a]
        e["cbp_luma4x4"] = int(r[k]); e["cbp_chroma_dc4x4"] = r[k + 1:k + 3]; e["cbp_chroma_ac4x4"] = r[k + 3:k + 5]; k += 7
        e["luma_level"] = r[k:k + 256].reshape(16, 16); k += 256
        e["chroma_dc_level"] = r[k:k + 8].reshape(2, 4); k += 8
        e["chroma_ac_level"] = r[k:k + 128].reshape(2, 4, 16); k += 128
        p["rec"][a] = r[k:k + 384]; k += 384
        p["state_in"][a]["chroma_ac_level"] = r[k:k + 128].reshape(2, 4, 16); k += 128
        e_type = int(r[k]); pwh = r[k + 1:k + 9].reshape(4, 2); k += 9
        p["state_in"][a]["chroma_dc_level"] = r[k:k + 8].reshape(2, 4)
        mode = {(1, 16, 16): 0, (2, 16, 8): 1, (2, 8, 16): 2, (4, 8, 8): 3}.get((nparts, pw, ph), -1)
        # macroblocks the reference predicts from real motion; the others (base macroblock intra: no partition, predFlagL0 = 0) are coded by the
        # reference against the prediction of an earlier macroblock: mark_inherited() below
        ok = mode >= 0 and all(int(nsub[i]) >= 1 and int(pflag[i]) == 1 and int(ridx[i]) == 0 and int(pwh[i, 0]) == pw and int(pwh[i, 1]) == ph for i in range(nparts))
        ok = ok and e_type in (301, 302, 303, 304)
        p["valid"][a] = 1 if ok else 0
        p["kind"] = 0
        if ok:
            m = p["motion"][a]
            m["part_mode"] = mode
            for i in range(nparts):
                m["mv"][i, 0] = mv[i, 0]
        p["seen"] += 1
    stale = {(int(r[2]), int(r[3]), int(r[4])): r[5:5 + 384].copy() for r in t.get(10, [])}
    out = []
    for key in order:
        p = pics[key]
        assert p["seen"] == len(p["valid"]), "trace holds %d of %d macroblocks of picture %r" % (p["seen"], len(p["valid"]), key)
        mark_inherited(p["motion"], p["valid"])
        p["stale"] = {a: stale[(key[0], key[1], a)] for a in range(len(p["valid"])) if (key[0], key[1], a) in stale}   # what the scratch blocks really held
        out.append(p)
    return out


def mark_inherited(motion, valid):
    """A macroblock without partitions (base macroblock intra inside a P picture) is coded by the reference against the prediction its scratch blocks still
    hold: that of the last macroblock WITH partitions.  Inside one picture that is expressible: pad[0] bit 0 + pad[1..2] = address of
    that macroblock (hlb_svc.cuh: SvcPredSrc); such a macroblock then has reference behaviour (valid = 1).  One that precedes every macroblock with partitions
    of its picture inherits from an earlier picture (or from the I_BL function's temporaries) and stays without (valid = 0)."""
    last = -1
    for a in range(len(valid)):
        if valid[a]:
            if not (motion[a]["pad"][0] & 1):
                last = a
        elif last >= 0:
            motion[a]["part_mode"] = 0
            motion[a]["pad"] = (1, last & 255, last >> 8)
            valid[a] = 1


def planes_of_mb(mbs, w, h):
    """inverse of mb_of_planes: (nmb, 384) macroblock samples -> tight Y|U|V"""
    mbw, mbh = w // 16, h // 16
    y = mbs[:, :256].reshape(mbh, mbw, 16, 16).transpose(0, 2, 1, 3).reshape(-1)
    u = mbs[:, 256:320].reshape(mbh, mbw, 8, 8).transpose(0, 2, 1, 3).reshape(-1)
    v = mbs[:, 320:].reshape(mbh, mbw, 8, 8).transpose(0, 2, 1, 3).reshape(-1)
    return np.concatenate([y, u, v]).astype(np.uint8)


def bl_pictures_from_trace(path):
    """enhancement-layer I pictures (I_BL macroblocks, tags 8 / 9): same dicts, `ref` holds the PREDICTION planes (resampled base layer), kind = 1"""
    t = reftrace.parse(path)
    pics, order = {}, []
    for r in t.get(9, []):
        W, H = int(r[4]), int(r[5])
        n = W * H * 3 // 2
        nmb = (W // 16) * (H // 16)
        key = (int(r[2]), int(r[3]))
        order.append(key)
        pics[key] = dict(frame=key[0], dqid=key[1], w=W, h=H, qp=-1, kind=1, src=r[6:].view(np.uint8)[:n].copy(), pred_mb=np.zeros((nmb, 384), np.int32),
                         motion=np.zeros(nmb, MB_MOTION), valid=np.ones(nmb, np.uint8), state_in=np.zeros(nmb, SVC_STATE), expect=np.zeros(nmb, MB_COEFFS),
                         rec=np.zeros((nmb, 384), np.uint8), seen=0)
    for r in t.get(8, []):
        p = pics[(int(r[2]), int(r[3]))]
        a = int(r[4])
        p["qp"], p["qpc"] = int(r[5]), int(r[6])
        k = 8
        e = p["expect"][a]
        e["cbp_luma4x4"] = int(r[k]); e["cbp_chroma_dc4x4"] = r[k + 1:k + 3]; e["cbp_chroma_ac4x4"] = r[k + 3:k + 5]; k += 7
        e["luma_level"] = r[k:k + 256].reshape(16, 16); k += 256
        e["chroma_dc_level"] = r[k:k + 8].reshape(2, 4); k += 8
        e["chroma_ac_level"] = r[k:k + 128].reshape(2, 4, 16); k += 128
        p["rec"][a] = r[k:k + 384]; k += 384
        p["state_in"][a]["chroma_ac_level"] = r[k:k + 128].reshape(2, 4, 16); k += 128
        p["state_in"][a]["chroma_dc_level"] = r[k:k + 8].reshape(2, 4); k += 8
        p["pred_mb"][a] = r[k:k + 384]
        p["seen"] += 1
    out = []
    for key in order:
        p = pics[key]
        assert p["seen"] == len(p["valid"])
        assert p["pred_mb"].min() >= 0 and p["pred_mb"].max() <= 255     # the resampling process clips to the sample range (G.8.6.2.3)
        p["ref"] = planes_of_mb(p.pop("pred_mb"), p["w"], p["h"])
        out.append(p)
    return out


_KEYS = ("src", "ref", "motion", "valid", "state_in", "expect", "rec")


def save_golden(named_pics, path=GOLDEN):
    d, index = {}, []
    for name, pics in named_pics:
        for i, p in enumerate(pics):
            tag = "%s.%d" % (name, i)
            index.append(tag)
            d[tag + ".meta"] = np.array([p["w"], p["h"], p["qp"], p["frame"], p["dqid"], p.get("kind", 0)], np.int32)
            for k in _KEYS:
                d[tag + "." + k] = p[k].view(np.uint8) if p[k].dtype.names else p[k]
    d["index"] = np.array(index)
    np.savez_compressed(path, **d)


def load_golden(path=GOLDEN):
    z = np.load(path)
    out = []
    for tag in z["index"]:
        tag = str(tag)
        w, h, qp, frame, dqid, kind = (int(v) for v in z[tag + ".meta"])
        p = dict(name=tag, w=w, h=h, qp=qp, frame=frame, dqid=dqid, kind=kind)   # kind 0: base-mode inter picture, 1: I_BL picture (ref = prediction planes)
        for k, dt in (("src", None), ("ref", None), ("motion", MB_MOTION), ("valid", None), ("state_in", SVC_STATE), ("expect", MB_COEFFS), ("rec", None)):
            a = z[tag + "." + k]
            p[k] = a.view(dt).reshape(-1) if dt is not None else a
        out.append(p)
    return out


def mb_of_planes(y, u, v, w, h):
    """tight planes -> (nmb, 384) macroblock samples (16x16 Y, 8x8 Cb, 8x8 Cr) as the trace stores them"""
    mbw, mbh = w // 16, h // 16
    yy = y.reshape(mbh, 16, mbw, 16).transpose(0, 2, 1, 3).reshape(mbw * mbh, 256)
    uu = u.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbw * mbh, 64)
    vv = v.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbw * mbh, 64)
    return np.concatenate([yy, uu, vv], axis=1)


def compare_picture(p, coeffs, rec_yuv, state_out=None, what=""):
    """coeffs (MB_COEFFS per macroblock) / rec_yuv (tight Y|U|V) of an implementation against the reference's picture `p`; valid macroblocks only"""
    w, h = p["w"], p["h"]
    ysz, csz = w * h, w * h // 4
    got = mb_of_planes(rec_yuv[:ysz], rec_yuv[ysz:ysz + csz], rec_yuv[ysz + csz:], w, h)
    n = 0
    for a in np.nonzero(p["valid"])[0]:
        e, g = p["expect"][a], coeffs[a]
        for f in ("cbp_luma4x4", "cbp_chroma_dc4x4", "cbp_chroma_ac4x4", "luma_level", "chroma_dc_level", "chroma_ac_level"):
            assert np.array_equal(e[f], g[f]), "%s %s: macroblock %d field %s\nref %s\ngot %s" % (what, p.get("name", ""), a, f, e[f], g[f])
        assert np.array_equal(p["rec"][a], got[a]), "%s %s: reconstruction of macroblock %d" % (what, p.get("name", ""), a)
        if state_out is not None:
            assert np.array_equal(state_out[a]["chroma_ac_level"], e["chroma_ac_level"]) and np.array_equal(state_out[a]["chroma_dc_level"], e["chroma_dc_level"]), a
        n += 1
    return n
