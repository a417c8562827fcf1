"""Generates tests/golden/encoder_*.npz from the UNMODIFIED reference encoder (oracle/_ref/hl_ref_driver, built from
/root/reference by oracle/build_ref.sh): per-frame reconstruction planes' MD5, per-MB decision fields, bitstream MD5.
Run in the build container:  python tests/golden/make_golden_encoder.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reftrace as rt  # noqa: E402

CONFIGS = [  # name, gen, seed, w, h, frames, qp, me_range[, max_ref_frame[, me_early_term_flag, deblock_flag]]
    ("g2_qcif", "g2", 5, 176, 144, 5, 30, 16),
    ("g1_qcif", "g1", 1, 176, 144, 5, 31, 16),
    ("g2_small_q12", "g2", 7, 64, 48, 6, 12, 64),
    ("g2_cif_q38", "g2", 2, 352, 288, 3, 38, 8),
    ("g2_qcif_ref4", "g2", 9, 176, 144, 6, 29, 24, 4),
    ("g1_cif_10", "g1", 12345, 352, 288, 10, 31, 16),      # BASELINE.json configs[0]: CIF, 1 ref, quarter-pel ME +-16, 10 frames
    # the configuration bench.py quotes its number on (BASELINE.json configs[1] at the reference's real feature set): 1920x1088, QP 31, ME +-32, 1 ref,
    # G1 with the seed of bench stream 0, IDR + 7 P pictures; the same on all-inter content (G2) and with max_ref_frame = 4 (configs[2])
    ("g1_1080p_q31", "g1", 12345, 1920, 1088, 8, 31, 32),
    ("g2_1080p_q31", "g2", 3, 1920, 1088, 4, 31, 32),
    ("g1_1080p_ref4", "g1", 12345, 1920, 1088, 6, 31, 32, 4),
    # the library's own defaults (hl_types.h:67,69): early termination (homogeneous-block mode mask, rdo.c:889-935) and the in-loop deblocking filter
    # (deblock.c:192) -- G3 content straddles the homogeneity thresholds; the flags one at a time and together, and the defaults at the bench size
    ("g3_cif_early", "g3", 4, 352, 288, 4, 20, 16, 1, 1, 0),
    ("g2_qcif_deblock", "g2", 5, 176, 144, 5, 30, 16, 1, 0, 1),
    ("g3_cif_defaults", "g3", 6, 352, 288, 5, 26, 16, 1, 1, 1),
    ("g1_1080p_defaults", "g1", 12345, 1920, 1088, 4, 31, 32, 1, 1, 1),
]


def kind_of(e_type):
    return {306: 0, 301: 1, 302: 1, 303: 1, 304: 1, 305: 1, 101: 3}.get(e_type, 2)


def build(name, gen, seed, w, h, frames, qp, me_range, refs=1, early_term=0, deblock=0):
    pre = "/tmp/golden_" + name
    s = rt.run_driver(pre, w, h, frames, gen=gen, seed=seed, qp=qp, me_range=me_range, refs=refs, early_term=early_term, deblock=deblock)
    t = rt.parse(pre + ".trace")
    nmb = (w // 16) * (h // 16)
    rec = {}
    for r in t[1]:
        d = rt.mb_record(r)
        rec[(d["frame"], d["addr"])] = d
    st = {(d["frame"], d["addr"]): d for d in map(rt.state_record, t[5])}
    fb = w * h * 3 // 2
    recon = np.fromfile(pre + ".recon", np.uint8).reshape(frames, fb)
    out = dict(config=np.array([w, h, frames, qp, me_range, seed], np.int32), refs=np.array(refs, np.int32), flags=np.array([early_term, deblock], np.int32), gen=np.array(gen), bitstream_md5=np.array(s["md5"]),
               recon_md5=np.array([hashlib.md5(recon[n].tobytes()).hexdigest() for n in range(frames)]))
    kind = np.zeros((frames, nmb), np.uint8)
    mb_type = np.zeros((frames, nmb), np.uint8)
    mv = np.zeros((frames, nmb, 4, 4, 2), np.int16)
    mvd = np.zeros((frames, nmb, 4, 4, 2), np.int16)
    nparts = np.zeros((frames, nmb, 5), np.uint8)   # NumMbPart, NumSubMbPart[4]
    cbp = np.zeros((frames, nmb, 3), np.uint8)      # coded_block_pattern, luma, chroma
    tc_luma = np.zeros((frames, nmb, 16), np.uint8)
    i4 = np.zeros((frames, nmb, 16), np.uint8)
    lvl_md5 = np.zeros((frames, nmb), "U32" if frames * nmb < 20000 else "U8")   # big pictures: the first 8 hex digits of each level digest
    mad = np.zeros((frames, nmb), np.int32)                                          # *pi_mad of the decision that stood (rdo.c:1216: best distortion)
    for n in range(frames):
        for a in range(nmb):
            r, q = rec[(n, a)], st[(n, a)]
            k = kind_of(q["e_type"])
            kind[n, a] = k
            mb_type[n, a] = r["mb_type"] if k else 5
            mv[n, a] = q["mv"]
            mvd[n, a] = r["mvd"]
            nparts[n, a, 0] = q["num_mb_part"]
            nparts[n, a, 1:] = q["num_sub"]
            cbp[n, a] = (r["coded_block_pattern"], q["cbp_luma"], q["cbp_chroma"])
            tc_luma[n, a] = q["tc_luma"]
            i4[n, a] = q["i4_mode"]
            # levels the writer consumes for this macroblock type
            if k == 1:
                lv = r["luma_level"] * ((r["cbp_luma4x4"] >> np.arange(16)) & 1)[:, None]
            elif k == 3:
                lv = r["luma_level"]
            elif k == 2:
                lv = np.concatenate([r["i16_dc"].reshape(1, 16), r["i16_ac"]])
            else:
                lv = np.zeros((1, 1), np.int32)
            lvl_md5[n, a] = hashlib.md5(lv.astype(np.int16).tobytes()).hexdigest()[:lvl_md5.dtype.itemsize // 4]
            mad[n, a] = r["mad"]
    out.update(kind=kind, mb_type=mb_type, mv=mv, mvd=mvd, nparts=nparts, cbp=cbp, tc_luma=tc_luma, i4_mode=i4, level_md5=lvl_md5, mad=mad)
    np.savez_compressed(os.path.join(HERE, "encoder_%s.npz" % name), **out)
    print(name, s["md5"], "kinds", np.bincount(kind.reshape(-1), minlength=4))


if __name__ == "__main__":
    only = sys.argv[1:]
    for c in CONFIGS:
        if not only or c[0] in only:
            build(*c)
