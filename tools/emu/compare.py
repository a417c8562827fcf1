#!/usr/bin/env python
"""Diffs the CPU emulation (tools/emu/emu) against a trace of the reference encoder, macroblock by macroblock.
usage: compare.py --gen g2 --size 352 288 --frames 4 [--seed 1] [--qp 31] [--me-range 16]"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reftrace as rt  # noqa: E402
from hartallo_b200 import lib as hl  # noqa: E402
from hartallo_b200 import synth  # noqa: E402

MBSTATE = np.dtype([("kind", "u1"), ("part_mode", "u1"), ("sub_mode", "u1", (4,)), ("cbp_luma", "u1"), ("cbp_chroma", "u1"), ("tc_luma", "u1", (16,)),
                    ("tc_cac", "u1", (2, 4)), ("ref_idx", "i1", (4,)), ("i4_mode", "u1", (16,)), ("last_sctr", "u1"), ("pad", "u1", (3,)),
                    ("mv", "<i2", (4, 4, 2)), ("chroma_ac", "<i2", (2, 4, 16)), ("chroma_dc", "<i2", (2, 4))])

E_TYPE_KIND = {306: 0, 301: 1, 302: 1, 303: 1, 304: 1, 305: 1}


def ref_kind(e_type):
    if e_type in E_TYPE_KIND:
        return E_TYPE_KIND[e_type]
    return 3 if e_type == 101 else 2   # I_NXN = 101, I_16X16_* follow


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen", default="g2")
    ap.add_argument("--size", type=int, nargs=2, default=[352, 288])
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--qp", type=int, default=31)
    ap.add_argument("--me-range", type=int, default=16)
    ap.add_argument("--max-report", type=int, default=8)
    ap.add_argument("--refs", type=int, default=1)
    a = ap.parse_args()
    w, h = a.size
    nmb = (w // 16) * (h // 16)
    pre = "/tmp/cmp_%s_%dx%d" % (a.gen, w, h)
    g = synth.make(a.gen, w, h, a.seed)
    with open(pre + ".yuv", "wb") as f:
        for _ in range(a.frames):
            f.write(g.next().tobytes())
    s = rt.run_driver(pre + "_ref", w, h, a.frames, gen=a.gen, seed=a.seed, qp=a.qp, me_range=a.me_range, refs=a.refs)
    subprocess.check_call([os.path.join(ROOT, "tools", "emu", "emu"), "--size", str(w), str(h), "--frames", str(a.frames), "--qp", str(a.qp),
                           "--me-range", str(a.me_range), "--refs", str(a.refs), "--in", pre + ".yuv", "--out", pre + "_emu"], stdout=subprocess.DEVNULL)
    t = rt.parse(pre + "_ref.trace")
    ref_rec = {}
    for r in t[1]:
        d = rt.mb_record(r)
        ref_rec[(d["frame"], d["addr"])] = d   # the last record of an MB wins (in P pictures the intra record precedes the inter one)
    ref_st = {(d["frame"], d["addr"]): d for d in map(rt.state_record, t[5])}
    fb = w * h * 3 // 2
    ref_recon = np.fromfile(pre + "_ref.recon", np.uint8).reshape(-1, fb)
    emu_recon = np.fromfile(pre + "_emu.recon", np.uint8).reshape(-1, fb)
    emu_rec = np.fromfile(pre + "_emu.rec", hl.MB_RECORD).reshape(-1, nmb)
    emu_st = np.fromfile(pre + "_emu.st", MBSTATE).reshape(-1, nmb)
    assert MBSTATE.itemsize == 392, MBSTATE.itemsize
    bad = 0
    for n in range(a.frames):
        fbad = 0
        for mb in range(nmb):
            rr, rs, er, es = ref_rec[(n, mb)], ref_st[(n, mb)], emu_rec[n, mb], emu_st[n, mb]
            diffs = []
            k = ref_kind(rs["e_type"])
            if k != es["kind"]:
                diffs.append("kind ref=%d(e_type %d) emu=%d" % (k, rs["e_type"], es["kind"]))
            else:
                if k in (0, 1):
                    npart = rs["num_mb_part"]
                    for p in range(npart):
                        for q in range(int(rs["num_sub"][p])):
                            if tuple(rs["mv"][p, q]) != tuple(es["mv"][p, q]):
                                diffs.append("mv[%d][%d] ref=%s emu=%s" % (p, q, rs["mv"][p, q], es["mv"][p, q]))
                if k == 1:
                    rm = {(1, 16, 16): 0, (2, 16, 8): 1, (2, 8, 16): 2}.get((rs["num_mb_part"], rs["part_w"], rs["part_h"]), 3)
                    if rm != es["part_mode"]:
                        diffs.append("part_mode ref=%d emu=%d" % (rm, es["part_mode"]))
                    if not np.array_equal(rr["mvd"][0, 0], er["mvd"][0, 0]) and rm == 0:
                        diffs.append("mvd ref=%s emu=%s" % (rr["mvd"][0, 0], er["mvd"][0, 0]))
                if rs["cbp_luma"] != es["cbp_luma"] or rs["cbp_chroma"] != es["cbp_chroma"]:
                    diffs.append("cbp ref=%d/%d emu=%d/%d" % (rs["cbp_luma"], rs["cbp_chroma"], es["cbp_luma"], es["cbp_chroma"]))
                if k != 0 and rr["mb_type"] != er["mb_type"]:
                    diffs.append("mb_type ref=%d emu=%d" % (rr["mb_type"], er["mb_type"]))
                if k == 3 and not np.array_equal(rs["i4_mode"], es["i4_mode"]):
                    diffs.append("i4_mode ref=%s emu=%s" % (rs["i4_mode"], es["i4_mode"]))
                if k == 2 and rr["i16_mode"] != er["i16_pred_mode"]:
                    diffs.append("i16 mode ref=%d emu=%d" % (rr["i16_mode"], er["i16_pred_mode"]))
                if k != 0 and rr["mad"] != er["mad"]:
                    diffs.append("mad ref=%d emu=%d" % (rr["mad"], er["mad"]))
            # TotalCoeffsLuma[] of a P_Skip macroblock is dead state (neighbours count 0 for a skipped macroblock, the next picture gates it by CodedBlockPatternLuma = 0):
            # the kernel does not run the trials that would only update it (hlb_mbcore.cuh: me_pskip_tail), so it is not compared there
            if es["kind"] != 0 and not np.array_equal(rs["tc_luma"], es["tc_luma"]):
                diffs.append("tc_luma ref=%s emu=%s" % (rs["tc_luma"], es["tc_luma"]))
            if not np.array_equal(rs["tc_cac"], es["tc_cac"]):
                diffs.append("tc_cac ref=%s emu=%s" % (rs["tc_cac"].reshape(-1), es["tc_cac"].reshape(-1)))
            if not np.array_equal(rs["chroma_ac"][..., :15], es["chroma_ac"][..., :15]):
                diffs.append("chroma_ac differs")
            if rs["last_single_ctr"] is not None and mb == nmb - 1:
                pass
            # reconstruction of this macroblock
            mbx, mby = mb % (w // 16), mb // (w // 16)
            ry = ref_recon[n, :w * h].reshape(h, w)[mby * 16:mby * 16 + 16, mbx * 16:mbx * 16 + 16]
            ey = emu_recon[n, :w * h].reshape(h, w)[mby * 16:mby * 16 + 16, mbx * 16:mbx * 16 + 16]
            if not np.array_equal(ry, ey):
                diffs.append("luma recon differs (%d px)" % int((ry != ey).sum()))
            for c, off in ((0, w * h), (1, w * h * 5 // 4)):
                rc = ref_recon[n, off:off + w * h // 4].reshape(h // 2, w // 2)[mby * 8:mby * 8 + 8, mbx * 8:mbx * 8 + 8]
                ec = emu_recon[n, off:off + w * h // 4].reshape(h // 2, w // 2)[mby * 8:mby * 8 + 8, mbx * 8:mbx * 8 + 8]
                if not np.array_equal(rc, ec):
                    diffs.append("chroma %d recon differs (%d px)" % (c, int((rc != ec).sum())))
            if diffs:
                fbad += 1
                if bad + fbad <= a.max_report:
                    print("frame %d mb %d (%d,%d): " % (n, mb, mbx, mby) + "; ".join(diffs))
        print("frame %d: %d / %d macroblocks differ; recon equal: %s" % (n, fbad, nmb, np.array_equal(ref_recon[n], emu_recon[n])))
        bad += fbad
    print("ref md5", s["md5"])
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
