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
        sw, sh = r[15:19], r[19:23]
        SUB = {(8, 8): (0, 1), (8, 4): (1, 2), (4, 8): (2, 2), (4, 4): (3, 4)}   # (SubMbPartWidth, SubMbPartHeight) -> (sub_mode, NumSubMbPart)
        ok = mode >= 0 and all(int(nsub[i]) >= 1 and int(pflag[i]) == 1 and int(ridx[i]) == 0 for i in range(nparts))
        if ok and mode == 3:    # P_8x8: sub-macroblock partitions of every shape (the general, non-dyadic case produces them); partWidth/Height[i][0] = the sub-partition's
            ok = all(SUB.get((int(sw[i]), int(sh[i])), (0, -1))[1] == int(nsub[i]) and int(pwh[i, 0]) == int(sw[i]) and int(pwh[i, 1]) == int(sh[i]) for i in range(4))
        elif ok:
            ok = all(int(pwh[i, 0]) == pw and int(pwh[i, 1]) == ph for i in range(nparts))
        ok = ok and e_type in (301, 302, 303, 304)
        p["valid"][a] = 1 if ok else 0
        p["kind"] = 0
        if ok:
            m = p["motion"][a]
            m["part_mode"] = mode
            for i in range(nparts):
                ns = int(nsub[i]) if mode == 3 else 1
                if mode == 3:
                    m["sub_mode"][i] = SUB[(int(sw[i]), int(sh[i]))][0]
                m["mv"][i, :ns] = mv[i, :ns]
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


# ---- inter-layer motion derivation (second half): tag 11 of oracle/ref_driver.c = what the derivation reads, tag 6 = what it produced -----------------
SVC_BASE_MB = np.dtype([("flags", "u1"), ("part_w", "u1"), ("part_h", "u1"), ("sub_w", "u1", (4,)), ("sub_h", "u1", (4,)), ("pred_flag", "i1", (4,)), ("ref_idx", "i1", (4,)),
                        ("pad", "u1"), ("mv", "<i2", (4, 4, 2))])   # hlb200_svc_base_mb_t
SVC_GEOM = np.dtype([("ref_width", "<i4"), ("ref_height", "<i4"), ("scaled_width", "<i4"), ("scaled_height", "<i4"), ("left_offset", "<i4"), ("top_offset", "<i4"),
                     ("level_idc", "<i4"), ("restricted", "<i4"), ("cropping_change", "<i4")])   # hlb200_svc_layer_geom_t
E_TYPE_I_BL = 426   # HL_CODEC_264_MB_TYPE_SVC_I_BL (hl_codec_264_defs.h:474)
DERIVE_BAD_REF, DERIVE_UNSUPPORTED, DERIVE_STALE_PARTS, DERIVE_NO_PRED_SOURCE = 1, 2, 4, 8


def base_mbs_from_words(w):
    """(n, 53) int32 words of a tag-11 record -> SVC_BASE_MB array"""
    b = np.zeros(len(w), SVC_BASE_MB)
    b["flags"] = (w[:, 0] != 0) * 1 + (w[:, 1] != 0) * 2 + (w[:, 2] != 0) * 4
    b["part_w"], b["part_h"] = w[:, 3], w[:, 4]
    b["sub_w"], b["sub_h"], b["pred_flag"], b["ref_idx"] = w[:, 5:9], w[:, 9:13], w[:, 13:17], w[:, 17:21]
    b["mv"] = w[:, 21:53].reshape(-1, 4, 4, 2)
    return b


def derive_pictures_from_trace(path):
    """one dict per enhancement-layer P picture, in coding order: geom / base = the derivation's inputs (tag 11); kind (0 inter, 1 base macroblock intra), part_mode,
    sub_mode, ref_idx, mv, nparts, nsub = what the reference derived per macroblock (tag 6: NumMbPart, MbPartWidth/Height, SubMbPartWidth/Height, refIdxL0, mvL0; only the
    entries below NumMbPart / NumSubMbPart are meaningful, the macroblock objects are never reset); stale_parts = NumSubMbPart[0] the object held (tag 6, unchanged by an
    intra derivation); motion / valid = the glue's view of it (pictures_from_trace)"""
    t = reftrace.parse(path)
    glue_view = {(p["frame"], p["dqid"]): p for p in pictures_from_trace(path)}
    out = []
    for r in t.get(11, []):
        key = (int(r[2]), int(r[3]))
        gv = glue_view[key]
        nmb = len(gv["valid"])
        g = np.zeros(1, SVC_GEOM)
        for i, f in enumerate(SVC_GEOM.names):
            g[f] = int(r[4 + i]) if f != "cropping_change" else int(r[12])
        g["restricted"] = int(r[11])
        nref = int(r[14])
        d = dict(frame=key[0], dqid=key[1], w=gv["w"], h=gv["h"], geom=g, spatial_change=int(r[13]), base=base_mbs_from_words(r[15:15 + 53 * nref].reshape(nref, 53)),
                 kind=np.zeros(nmb, np.uint8), part_mode=np.zeros(nmb, np.uint8), sub_mode=np.zeros((nmb, 4), np.uint8), ref_idx=np.zeros((nmb, 4), np.int8),
                 mv=np.zeros((nmb, 4, 4, 2), np.int16), nparts=np.zeros(nmb, np.uint8), nsub=np.zeros((nmb, 4), np.uint8), stale_parts=np.zeros(nmb, np.uint8),
                 motion=gv["motion"], valid=gv["valid"])
        out.append(d)
    by_key = {(d["frame"], d["dqid"]): d for d in out}
    for r in t.get(6, []):
        d = by_key.get((int(r[2]), int(r[3])))
        if d is None:
            continue
        a = int(r[4])
        nparts, pw, ph = int(r[8]), int(r[9]), int(r[10])
        nsub, sw, sh, ridx, mv = r[11:15], r[15:19], r[19:23], r[27:31], r[31:63].reshape(4, 4, 2)
        e_type = int(r[63 + 7 + 256 + 8 + 128 + 384 + 128])
        if e_type == E_TYPE_I_BL:
            d["kind"][a] = 1
            d["stale_parts"][a] = int(nsub[0])
            continue
        mode = {(1, 16, 16): 0, (2, 16, 8): 1, (2, 8, 16): 2, (4, 8, 8): 3}[(nparts, pw, ph)]
        d["part_mode"][a], d["nparts"][a] = mode, nparts
        for p in range(nparts):
            d["ref_idx"][a, p] = int(ridx[p])
            d["nsub"][a, p] = int(nsub[p])
            if mode == 3:
                d["sub_mode"][a, p] = {(8, 8): 0, (8, 4): 1, (4, 8): 2, (4, 4): 3}[(int(sw[p]), int(sh[p]))]
            d["mv"][a, p, :int(nsub[p])] = mv[p, :int(nsub[p])]
    return out


def compare_derived(d, motion, status, what=""):
    """motion (MB_MOTION per macroblock) / status bits of an implementation of the derivation against the reference's picture `d`.  Every inter macroblock: partition layout,
    refIdxL0 and mvL0 of every (sub-)partition; every macroblock the glue codes (valid): the inherited-prediction marker; the status bits = the reasons the glue refuses"""
    n = 0
    inter_ok = (d["kind"] == 0) & np.array([all(int(d["ref_idx"][a, p]) == 0 for p in range(int(d["nparts"][a]))) for a in range(len(d["kind"]))])
    for a in range(len(d["kind"])):
        m = motion[a]
        tag = "%s frame %d dqid %d macroblock %d" % (what, d["frame"], d["dqid"], a)
        if d["kind"][a] == 0:
            assert int(m["part_mode"]) == int(d["part_mode"][a]), tag + ": part_mode %d, reference %d" % (m["part_mode"], d["part_mode"][a])
            for p in range(int(d["nparts"][a])):
                assert int(m["ref_idx"][p]) == int(d["ref_idx"][a, p]), tag + ": refIdxL0[%d]" % p
                if d["part_mode"][a] == 3:
                    assert int(m["sub_mode"][p]) == int(d["sub_mode"][a, p]), tag + ": sub_mode[%d]" % p
                ns = int(d["nsub"][a, p])
                assert np.array_equal(m["mv"][p, :ns], d["mv"][a, p, :ns]), tag + ": mvL0[%d] %s, reference %s" % (p, m["mv"][p, :ns].tolist(), d["mv"][a, p, :ns].tolist())
            assert not (int(m["pad"][0]) & 1), tag
        # the glue's view (pictures_from_trace / mark_inherited) is only defined for macroblocks it would code: inter ones inside its pinned set and intra-base ones;
        # an inter macroblock outside the set (sub-macroblock partitions in the general case) makes the glue refuse the picture
        if d["valid"][a] and (d["kind"][a] == 1 or inter_ok[a]) and not ((d["kind"] == 0) & ~inter_ok)[:a].any():
            g = d["motion"][a]
            assert np.array_equal(m["pad"], g["pad"]), tag + ": inherited-prediction marker %s, glue %s" % (m["pad"].tolist(), g["pad"].tolist())
            if not (int(g["pad"][0]) & 1):
                assert int(m["part_mode"]) == int(g["part_mode"]) and np.array_equal(m["sub_mode"], g["sub_mode"]) and np.array_equal(m["mv"], g["mv"]), tag
        n += 1
    expect = 0
    if ((d["kind"] == 0) & ~inter_ok).any():
        expect |= DERIVE_UNSUPPORTED
    if ((d["kind"] == 1) & (d["stale_parts"] != 0)).any():
        expect |= DERIVE_STALE_PARTS
    first_ok = np.nonzero(inter_ok)[0]
    intra = np.nonzero(d["kind"] == 1)[0]
    if len(intra) and (not len(first_ok) or intra[0] < first_ok[0]):
        expect |= DERIVE_NO_PRED_SOURCE
    assert status == expect, "%s frame %d dqid %d: status %d, expected %d" % (what, d["frame"], d["dqid"], status, expect)
    return n


DERIVE_GOLDEN = os.path.join(ROOT, "tests", "golden", "svc_derive.npz")


def load_derive_golden(path=DERIVE_GOLDEN):
    """pictures of tests/golden/svc_derive.npz (made by tests/golden/make_golden_svc_derive.py) in coding order; pictures of one stream share the name prefix"""
    z = np.load(path)
    out = []
    for tag in z["index"]:
        tag = str(tag)
        w, h, frame, dqid, sc = (int(v) for v in z[tag + ".meta"])
        p = dict(name=tag, stream=tag.rsplit(".", 1)[0], w=w, h=h, frame=frame, dqid=dqid, spatial_change=sc)
        for k, dt in (("geom", SVC_GEOM), ("base", SVC_BASE_MB), ("motion", MB_MOTION)):
            p[k] = z[tag + "." + k].view(dt).reshape(-1)
        for k in ("kind", "part_mode", "sub_mode", "ref_idx", "mv", "nparts", "nsub", "stale_parts", "valid"):
            p[k] = z[tag + "." + k]
        out.append(p)
    return out


def base_words(base):
    """SVC_BASE_MB array -> (n, 53) int32 words in the order of trace tag 11 (what oracle/hl_oracle.c: hlo_svc_derive_mb reads)"""
    w = np.zeros((len(base), 53), np.int32)
    f = base["flags"].astype(np.int32)
    w[:, 0], w[:, 1], w[:, 2] = f & 1, (f >> 1) & 1, (f >> 2) & 1
    w[:, 3], w[:, 4] = base["part_w"], base["part_h"]
    w[:, 5:9], w[:, 9:13], w[:, 13:17], w[:, 17:21] = base["sub_w"], base["sub_h"], base["pred_flag"], base["ref_idx"]
    w[:, 21:53] = base["mv"].reshape(len(base), 32)
    return w


def oracle_derive(olib, base, geom, w, h):
    """the oracle's derivation of a picture -> (MB_MOTION array in the C-ABI's layout incl. inherited-prediction markers, kind, status bits); had_parts is the caller's"""
    import ctypes as C
    g = geom[0]
    nmb = (w // 16) * (h // 16)
    bw = np.ascontiguousarray(base_words(base))
    out, bad = np.zeros((nmb, 52), np.int32), np.zeros(nmb, np.uint8)
    olib.hlo_svc_derive_picture(bw.ctypes.data_as(C.c_void_p), int(g["ref_width"]), int(g["ref_height"]), int(g["scaled_width"]), int(g["scaled_height"]), int(g["left_offset"]),
                                int(g["top_offset"]), int(g["level_idc"]), int(g["restricted"]), w, h, out.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p))
    return out, bad


def motion_from_oracle(out, bad, had_parts):
    """oracle records -> (MB_MOTION array, status bits) with the classification / inheritance rules of host/hlb200_glue.c (restated here in numpy, not shared with the device)"""
    nmb = len(out)
    m, status, kind = np.zeros(nmb, MB_MOTION), 0, np.zeros(nmb, np.uint8)
    for a in range(nmb):
        o = out[a]
        if bad[a]:
            status |= DERIVE_BAD_REF; kind[a] = 2
            continue
        if o[0]:
            kind[a] = 1
            if had_parts[a]:
                status |= DERIVE_STALE_PARTS
            continue
        n, pw, ph = int(o[1]), int(o[2]), int(o[3])
        mode = {(1, 16, 16): 0, (2, 16, 8): 1, (2, 8, 16): 2, (4, 8, 8): 3}[(n, pw, ph)]
        m[a]["part_mode"] = mode
        sup = True
        for p in range(n):
            ns = int(o[4 + p])
            m[a]["ref_idx"][p] = int(o[16 + p])
            if mode == 3:
                m[a]["sub_mode"][p] = {(8, 8): 0, (8, 4): 1, (4, 8): 2, (4, 4): 3}[(int(o[8 + p]), int(o[12 + p]))]
            m[a]["mv"][p, :ns] = o[20:52].reshape(4, 4, 2)[p, :ns]
            sup = sup and int(o[16 + p]) == 0
        had_parts[a] = 1
        if not sup:
            status |= DERIVE_UNSUPPORTED; kind[a] = 2
    last = -1
    for a in range(nmb):
        if kind[a] == 0:
            last = a
        elif kind[a] == 1:
            if last < 0 or last >= 65536:
                status |= DERIVE_NO_PRED_SOURCE
            else:
                m[a]["pad"] = (1, last & 255, last >> 8)
    return m, status


def random_base_field(rng, ref_w, ref_h, intra_frac=0.1, zero_mv_frac=0.3):
    """random reference-layer macroblock fields with every partition layout, intra macroblocks, unused partitions (predFlagL0 = 0) and non-zero reference indices"""
    n = (ref_w // 16) * (ref_h // 16)
    b = np.zeros(n, SVC_BASE_MB)
    mode = rng.integers(0, 4, n)
    b["part_w"], b["part_h"] = np.array([16, 16, 8, 8])[mode], np.array([16, 8, 16, 8])[mode]
    sm = rng.integers(0, 4, (n, 4))
    b["sub_w"], b["sub_h"] = np.array([8, 8, 4, 4])[sm], np.array([8, 4, 8, 4])[sm]
    b["sub_w"][mode != 3], b["sub_h"][mode != 3] = b["part_w"][mode != 3, None], b["part_h"][mode != 3, None]
    b["flags"] = (mode == 3) * 4
    b["pred_flag"] = (rng.random((n, 4)) > 0.05)
    b["ref_idx"] = np.where(rng.random((n, 4)) > 0.1, 0, rng.integers(0, 3, (n, 4)))
    mv = rng.integers(-64, 65, (n, 4, 4, 2))
    same = rng.random(n) < zero_mv_frac          # macroblocks whose partitions all move alike (they merge to larger partitions in the enhancement layer)
    mv[same] = mv[same][:, :1, :1, :]
    b["mv"] = mv
    intra = rng.random(n) < intra_frac
    b["flags"][intra] = np.where(rng.random(int(intra.sum())) < 0.5, 3, 2)   # Intra16x16 (recognised by e_type) or Intra4x4 of a P picture (flags_type only: motion is derived from it)
    return b
