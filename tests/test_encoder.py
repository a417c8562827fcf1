"""Slice-level parity (hlb200_slice_encode = the per-MB decide+reconstruct loop of hl_codec_264_nal_slice_data_encode):
macroblock types, partitions, motion vectors, mvd, coded block patterns, quantised levels and reconstructed pictures must be
identical to the reference encoder's.  Counterpart of the reference's source/test_encoder.c (which only prints fps).

Golden data = outputs of the UNMODIFIED reference C path (tests/golden/make_golden_encoder.py).  Where oracle/_ref/hl_ref_driver
is present the reference is additionally run live on fresh seeds.

* CPU (`-m "not gpu"`): the control flow shared with the kernel, compiled as plain C++ (tools/emu), against the golden data.
* GPU (`-m gpu`): the CUDA path through the C-ABI.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import reftrace as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CONFIGS = ["g2_qcif", "g1_qcif", "g2_small_q12", "g2_cif_q38", "g2_qcif_ref4", "g1_cif_10", "g3_cif_early", "g2_qcif_deblock", "g3_cif_defaults"]
# the configuration bench.py quotes its number on (1920x1088, QP 31, ME +-32, G1 seed of bench stream 0), its all-inter twin (G2) and max_ref_frame = 4: GPU tier only
CONFIGS_1080P = ["g1_1080p_q31", "g2_1080p_q31", "g1_1080p_ref4", "g1_1080p_defaults"]


def refs_of(g):
    """max_ref_frame of a golden configuration (1 unless stored)"""
    return int(g["refs"]) if "refs" in g.files else 1


def flags_of(g):
    """(me_early_term_flag, deblock_flag) of a golden configuration (0, 0 unless stored: source/test_encoder.c:140,146)"""
    return (int(g["flags"][0]), int(g["flags"][1])) if "flags" in g.files else (0, 0)


def frames_of(gen, seed, w, h, n):
    from hartallo_b200 import synth
    g = synth.make(gen, w, h, seed)
    return [g.next() for _ in range(n)]


def level_md5(kind, rec):
    if kind == 1:
        lv = rec["luma_level"].astype(np.int16) * ((int(rec["cbp_luma4x4"]) >> np.arange(16)) & 1)[:, None].astype(np.int16)
    elif kind == 3:
        lv = rec["luma_level"].astype(np.int16)
    elif kind == 2:
        lv = np.concatenate([rec["i16_dc_level"].reshape(1, 16), rec["i16_ac_level"]]).astype(np.int16)
    else:
        lv = np.zeros((1, 1), np.int16)
    return hashlib.md5(lv.tobytes()).hexdigest()


def _first(mask):
    """index of the first True of a per-macroblock mask (for the assertion message)"""
    i = np.flatnonzero(mask.reshape(mask.shape[0], -1).any(axis=1))
    return int(i[0]) if len(i) else -1


def check_frame(g, n, rec, recon, what):
    """rec: MB_RECORD array of frame n; recon: tight planes.  Vectorised over the macroblocks of the picture (a 1080p picture has 8,160 of them)."""
    assert hashlib.md5(recon.tobytes()).hexdigest() == str(g["recon_md5"][n]), "%s: reconstruction of frame %d differs" % (what, n)
    kind = g["kind"][n].astype(np.int64)
    bad = rec["mb_class"] != kind
    assert not bad.any(), (what, n, _first(bad), "class")
    coded = kind != 0
    bad = coded & (rec["mb_type"] != g["mb_type"][n])
    assert not bad.any(), (what, n, _first(bad), "mb_type")
    npart, nsub = g["nparts"][n][:, 0].astype(np.int64), g["nparts"][n][:, 1:5].astype(np.int64)
    valid = (np.arange(4)[None, :, None] < npart[:, None, None]) & (np.arange(4)[None, None, :] < nsub[:, :, None])   # (nmb, part, sub)
    inter = valid & np.isin(kind, (0, 1))[:, None, None]
    bad = inter[..., None] & (rec["mv"] != g["mv"][n])
    assert not bad.any(), (what, n, _first(bad), "mv")
    bad = (valid & (kind == 1)[:, None, None])[..., None] & (rec["mvd"] != g["mvd"][n])
    assert not bad.any(), (what, n, _first(bad), "mvd")
    got_cbp = np.stack([rec["coded_block_pattern"], rec["cbp_luma"], rec["cbp_chroma"]], axis=1)
    bad = got_cbp != g["cbp"][n]
    assert not bad.any(), (what, n, _first(bad), "cbp")
    bad = (kind == 3)[:, None] & (rec["i4_pred_mode"] != g["i4_mode"][n])
    assert not bad.any(), (what, n, _first(bad), "i4 modes")
    if "mad" in g.files:   # *pi_mad of the decision that stood: the search's best distortion (north star: "identical MVs and costs")
        bad = coded & (rec["mad"] != g["mad"][n])
        assert not bad.any(), (what, n, _first(bad), "mad", int(rec["mad"][_first(bad)]), int(g["mad"][n][_first(bad)]))
    # level digests: what the writer consumes for the macroblock's type
    gold = g["level_md5"][n]
    glen = len(str(gold[0]))
    mask16 = ((rec["cbp_luma4x4"].astype(np.int64)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.int16)
    inter_lv = np.ascontiguousarray(rec["luma_level"].astype(np.int16) * mask16[:, :, None])
    i4_lv = np.ascontiguousarray(rec["luma_level"].astype(np.int16))
    i16_lv = np.ascontiguousarray(np.concatenate([rec["i16_dc_level"].reshape(-1, 1, 16), rec["i16_ac_level"]], axis=1).astype(np.int16))
    zero = hashlib.md5(np.zeros((1, 1), np.int16).tobytes()).hexdigest()[:glen]
    for a in range(len(rec)):
        k = kind[a]
        d = zero if k == 0 else hashlib.md5((inter_lv[a] if k == 1 else (i4_lv[a] if k == 3 else i16_lv[a])).tobytes()).hexdigest()[:glen]
        assert d == str(gold[a]), (what, n, a, "levels")


def run_emu(w, h, frames, qp, me_range, yuv_frames, tag, refs=1, flags=(0, 0)):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools", "emu"), "emu"], stdout=subprocess.DEVNULL)
    from hartallo_b200 import lib as hl
    pre = "/tmp/test_emu_" + tag
    with open(pre + ".yuv", "wb") as f:
        for fr in yuv_frames:
            f.write(fr.tobytes())
    subprocess.check_call([os.path.join(ROOT, "tools", "emu", "emu"), "--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp),
                           "--me-range", str(me_range), "--refs", str(refs), "--early-term", str(flags[0]), "--deblock", str(flags[1]), "--in", pre + ".yuv", "--out", pre],
                          stdout=subprocess.DEVNULL)
    nmb = (w // 16) * (h // 16)
    rec = np.fromfile(pre + ".rec", hl.MB_RECORD).reshape(frames, nmb)
    recon = np.fromfile(pre + ".recon", np.uint8).reshape(frames, -1)
    return rec, recon


@pytest.mark.parametrize("name", CONFIGS)
def test_control_flow_vs_reference_cpu(name):
    g = np.load(os.path.join(GOLD, "encoder_%s.npz" % name))
    w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
    fr = frames_of(str(g["gen"]), seed, w, h, frames)
    rec, recon = run_emu(w, h, frames, qp, me_range, fr, name, refs_of(g), flags_of(g))
    for n in range(frames):
        check_frame(g, n, rec[n], recon[n], "emu/" + name)


@pytest.fixture(params=["cta", "warp"])
def variant(request):
    """both slice-kernel variants (one CTA / one warp per macroblock) must give the reference's results"""
    from hartallo_b200 import lib as hl
    prev = hl.load().hlb200_slice_set_variant(0 if request.param == "cta" else 1)
    yield request.param
    hl.load().hlb200_slice_set_variant(prev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONFIGS + CONFIGS_1080P)
def test_slice_encode_vs_reference(name, variant):
    from hartallo_b200 import lib as hl
    g = np.load(os.path.join(GOLD, "encoder_%s.npz" % name))
    w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
    fr = frames_of(str(g["gen"]), seed, w, h, frames)
    early, deblock = flags_of(g)
    enc = hl.Encoder(w, h, qp=qp, me_range=me_range, refs=refs_of(g), early_term=early, deblock=deblock)
    for n in range(frames):
        rec, recon = enc.encode(fr[n], want_recon=True)
        check_frame(g, n, rec, recon, "gpu/" + name)
    enc.close()


@pytest.mark.gpu
def test_slice_encode_batch_of_streams(variant):
    """several independent streams (different sizes and contents) in one launch give the same pictures as one by one"""
    from hartallo_b200 import lib as hl
    names = ["g2_qcif", "g1_qcif", "g2_small_q12"]
    gs = [np.load(os.path.join(GOLD, "encoder_%s.npz" % n)) for n in names]
    encs, frs = [], []
    for g in gs:
        w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
        encs.append(hl.Encoder(w, h, qp=qp, me_range=me_range))
        frs.append(frames_of(str(g["gen"]), seed, w, h, frames))
    for n in range(5):
        ps = hl.encode_batch(encs, [f[n] for f in frs])
        for i, (e, g) in enumerate(zip(encs, gs)):
            hl.check(e.st.lib.hlb200_stream_sync(encs[0].st.ctx), "sync")
            rec = np.zeros(e.st.nmb, hl.MB_RECORD)
            hl.check(e.st.lib.hlb200_records_download(e.st.ctx, hl.ptr(rec)), "records_download")
            check_frame(g, n, rec, e.st.download_slot(ps[i].cur_slot), "batch/" + names[i])
    for e in encs:
        e.close()


@pytest.mark.gpu
@pytest.mark.skipif(not rt.have_driver(), reason="oracle/_ref/hl_ref_driver not built")
def test_slice_encode_vs_live_reference():
    """fresh seed, reference run live beside the device path: reconstruction MD5 of every picture"""
    from hartallo_b200 import lib as hl
    w, h, frames, qp, me_range, seed = 176, 144, 4, 27, 24, 11
    rt.run_driver("/tmp/live_ref", w, h, frames, gen="g2", seed=seed, qp=qp, me_range=me_range, levels=False)
    ref = np.fromfile("/tmp/live_ref.recon", np.uint8).reshape(frames, -1)
    fr = frames_of("g2", seed, w, h, frames)
    enc = hl.Encoder(w, h, qp=qp, me_range=me_range)
    for n in range(frames):
        _, recon = enc.encode(fr[n], want_recon=True)
        assert np.array_equal(recon, ref[n]), n
    enc.close()


@pytest.mark.gpu
@pytest.mark.skipif(not rt.have_driver(), reason="oracle/_ref/hl_ref_driver not built")
def test_gpu_fuzz_vs_live_reference():
    """tools/gpu_fuzz.py: random configurations (size, QP, range, generator, seed, max_ref_frame, early termination, deblocking), both kernel variants, against the
    reference encoder run live on the box -- the 32-lane reductions of the search / intra decision have no CPU twin (the emulation harness runs one lane)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_fuzz.py"), "16", "101"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("gen", ["g1", "g2"])
def test_slice_encode_1080p_recon_md5(gen, variant):
    """BASELINE.json's full size (1920x1088, ME +-32): reconstruction MD5 of an IDR and a P picture against the reference's"""
    import json
    from hartallo_b200 import lib as hl
    gold = json.load(open(os.path.join(GOLD, "encoder_1080p.json")))
    c = gold["config"]
    fr = frames_of(gen, c["seed"], c["w"], c["h"], c["frames"])
    enc = hl.Encoder(c["w"], c["h"], qp=c["qp"], me_range=c["me_range"])
    for n in range(c["frames"]):
        _, recon = enc.encode(fr[n], want_recon=True)
        assert hashlib.md5(recon.tobytes()).hexdigest() == gold["recon_md5"][gen][n], (gen, n)
    enc.close()


@pytest.mark.gpu
def test_slice_encode_large_batch_is_deterministic():
    """24 concurrent 1080p streams (3 distinct contents x 8 copies) in one launch, under contention for the persistent CTAs / warps:
    every copy must reconstruct the same picture, both kernel variants must agree, and the stream whose content the reference encoded
    (tests/golden/encoder_1080p.json) must match the reference's MD5 -- a race in the scheduler or in a phase barrier would show here"""
    import json
    from hartallo_b200 import lib as hl
    gold = json.load(open(os.path.join(GOLD, "encoder_1080p.json")))
    c = gold["config"]
    w, h, nfr = c["w"], c["h"], c["frames"]
    contents = [frames_of("g1", c["seed"], w, h, nfr), frames_of("g1", c["seed"] + 1, w, h, nfr), frames_of("g2", c["seed"], w, h, nfr)]
    digests = {}
    for variant in (0, 1):
        prev = hl.load().hlb200_slice_set_variant(variant)
        encs = [hl.Encoder(w, h, qp=c["qp"], me_range=c["me_range"]) for _ in range(24)]
        for n in range(nfr):
            ps = hl.encode_batch(encs, [contents[i % 3][n] for i in range(24)])
            encs[0].st.slice_status()   # drains the launch stream and checks the watchdog words
            md5 = [hashlib.md5(e.st.download_slot(ps[i].cur_slot).tobytes()).hexdigest() for i, e in enumerate(encs)]
            for i in range(24):
                assert md5[i] == md5[i % 3], (variant, n, i)
            assert md5[0] == gold["recon_md5"]["g1"][n], (variant, n)
            assert md5[2] == gold["recon_md5"]["g2"][n], (variant, n)
            digests[(variant, n)] = md5[:3]
        for e in encs:
            e.close()
        hl.load().hlb200_slice_set_variant(prev)
    for n in range(nfr):
        assert digests[(0, n)] == digests[(1, n)], n


B200_ENCODER = os.path.join(ROOT, "oracle", "_ref", "hl_b200_encoder")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200_ENCODER), reason="oracle/_ref/hl_b200_encoder not built (needs the reference tree at build time)")
@pytest.mark.parametrize("name", CONFIGS + CONFIGS_1080P)
def test_bitstream_md5_drop_in(name):
    """The drop-in itself: the reference's unmodified host code (headers, CAVLC writer, DPB bookkeeping) linked with
    host/hlb200_glue.c + libhl_b200.so must emit the same bitstream, byte for byte, as the all-CPU reference
    (counterpart of source/test_encoder.c, which writes ./encoder.264 and checks nothing)."""
    import json
    g = np.load(os.path.join(GOLD, "encoder_%s.npz" % name))
    w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
    out = subprocess.run([B200_ENCODER, "--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp), "--me-range", str(me_range), "--refs", str(refs_of(g)),
                          "--gen", str(g["gen"]), "--seed", str(seed), "--early-term", str(flags_of(g)[0]), "--deblock", str(flags_of(g)[1])],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr[-500:]
    got = json.loads(out.stdout.strip().splitlines()[-1])
    assert got["md5"] == str(g["bitstream_md5"]), (got, str(g["bitstream_md5"]))


@pytest.mark.parametrize("prog", ["check_interp", "check_cavlc", "check_fast"])
def test_compact_primitives_equal_reference_formulation(prog):
    """the loop-form interpolation / register-only CAVLC length used by the kernels == the straightforward formulations
    (which the oracle pins against the reference), on random inputs incl. saturated content"""
    exe = "/tmp/hlb_" + prog
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-x", "c++", os.path.join(ROOT, "tools", "emu", prog + ".cpp"), "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout[-300:]


@pytest.mark.skipif(not rt.have_driver(), reason="oracle/_ref/hl_ref_driver not built (needs the reference tree)")
def test_control_flow_fuzz_vs_live_reference():
    """a few random configurations (size, QP 12..51, search range 1..64, generator, seed, max_ref_frame) of the differential fuzz
    tools/emu/fuzz.py: the CPU build of the kernel's per-macroblock code against the reference encoder run live"""
    out = subprocess.run([os.sys.executable, os.path.join(ROOT, "tools", "emu", "fuzz.py"), "8", "1000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and ", 0 mismatches" in out.stdout, out.stdout[-600:]
