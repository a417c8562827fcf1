"""Parser of the int32 record stream written by oracle/_ref/hl_ref_driver --trace (layouts documented in oracle/ref_driver.c)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "hl_ref_driver")
MB_REC_HDR = 188


def have_driver():
    return os.path.exists(DRIVER) and os.access(DRIVER, os.X_OK)


def run_driver(out_prefix, w, h, frames, gen="g1", seed=1, qp=31, me_range=16, refs=1, cand=False, state=True, levels=True, early_term=0, deblock=0):
    """runs the reference encoder; returns dict(json summary) and writes <prefix>.trace/.recon/.264"""
    import json
    cmd = [DRIVER, "--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp), "--me-range", str(me_range), "--refs", str(refs), "--gen", gen,
           "--seed", str(seed), "--trace", out_prefix + ".trace", "--recon", out_prefix + ".recon", "--out", out_prefix + ".264"]
    if early_term:
        cmd += ["--early-term", "1"]
    if deblock:
        cmd += ["--deblock", "1"]
    if cand:
        cmd.append("--trace-cand")
    if state:
        cmd.append("--trace-state")
    if not levels:
        cmd.append("--no-levels")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def parse(path):
    """returns dict tag -> list of int32 arrays (whole record incl. [tag, n])"""
    a = np.fromfile(path, np.int32)
    out = {1: [], 2: [], 3: [], 4: [], 5: []}
    i = 0
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        out.setdefault(tag, []).append(a[i:i + n])
        i += n
    return out


def mb_record(r):
    """tag-1 record -> dict"""
    d = dict(frame=int(r[2]), addr=int(r[3]), is_p=int(r[4]), which=int(r[5]), e_type=int(r[6]), mb_type=int(r[7]), flags_type=int(r[8]), num_mb_part=int(r[9]),
             part_w=int(r[10]), part_h=int(r[11]), num_sub=r[12:16].copy(), sub_mb_type=r[16:20].copy(), sub_w=r[20:24].copy(), sub_h=r[24:28].copy(),
             mv=r[28:60].reshape(4, 4, 2).copy(), mvd=r[60:92].reshape(4, 4, 2).copy(), ref_idx=r[92:96].copy(), pred_flag=r[96:100].copy(),
             coded_block_pattern=int(r[100]), cbp_luma=int(r[101]), cbp_chroma=int(r[102]), cbp_luma4x4=int(r[103]), cbp_dc4x4=r[104:106].copy(),
             cbp_ac4x4=r[106:108].copy(), i16_mode=int(r[108]), i4_mode=r[109:125].copy(), chroma_mode=int(r[125]), prev_i4=r[126:142].copy(),
             rem_i4=r[142:158].copy(), qp_delta=int(r[158]), qpy=int(r[159]), qpc=r[160:162].copy(), mad=int(r[162]), err=int(r[163]),
             tc_luma=r[164:180].copy(), tc_cac=r[180:188].reshape(2, 4).copy())
    if len(r) > MB_REC_HDR:
        k = MB_REC_HDR
        d["luma_level"] = r[k:k + 256].reshape(16, 16).copy(); k += 256
        d["chroma_dc"] = r[k:k + 8].reshape(2, 4).copy(); k += 8
        d["chroma_ac"] = r[k:k + 128].reshape(2, 4, 16).copy(); k += 128
        d["i16_dc"] = r[k:k + 16].copy(); k += 16
        d["i16_ac"] = r[k:k + 256].reshape(16, 16).copy()
    return d


def state_record(r):
    """tag-5 record -> dict"""
    k = 2
    d = dict(frame=int(r[2]), addr=int(r[3]), e_type=int(r[4]), flags_type=int(r[5]), num_mb_part=int(r[6]), part_w=int(r[7]), part_h=int(r[8]))
    k = 9
    d["num_sub"] = r[k:k + 4].copy(); k += 4
    d["sub_w"] = r[k:k + 4].copy(); k += 4
    d["sub_h"] = r[k:k + 4].copy(); k += 4
    d["cbp_luma"], d["cbp_chroma"] = int(r[k]), int(r[k + 1]); k += 2
    d["ref_idx"] = r[k:k + 4].copy(); k += 4
    d["pred_flag"] = r[k:k + 4].copy(); k += 4
    d["mv"] = r[k:k + 32].reshape(4, 4, 2).copy(); k += 32
    d["tc_luma"] = r[k:k + 16].copy(); k += 16
    d["tc_cac"] = r[k:k + 8].reshape(2, 4).copy(); k += 8
    d["i4_mode"] = r[k:k + 16].copy(); k += 16
    d["chroma_ac"] = r[k:k + 128].reshape(2, 4, 16).copy(); k += 128
    d["last_single_ctr"] = int(r[k])
    return d


def me_record(r):
    """tag-2 record -> dict"""
    d = dict(frame=int(r[2]), addr=int(r[3]), mode=int(r[4]), ref=int(r[5]), pskip=int(r[6]), num_mb_part=int(r[7]), num_sub=r[8:12].copy())
    k = 12
    d["mv"] = r[k:k + 32].reshape(4, 4, 2).copy(); k += 32
    d["mvp"] = r[k:k + 32].reshape(4, 4, 2).copy(); k += 32
    d["dist"] = r[k:k + 16].reshape(4, 4).copy(); k += 16
    d["sctr"] = r[k:k + 16].reshape(4, 4).copy(); k += 16
    d["cbp"] = r[k:k + 16].reshape(4, 4).copy(); k += 16
    d["cost"] = r[k:k + 32].copy().view(np.float64).reshape(4, 4); k += 32
    d["tc_luma"] = r[k:k + 16].copy()
    return d
