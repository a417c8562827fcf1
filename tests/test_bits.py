"""Device-side CAVLC serialisation (hartallo_b200/csrc/hlb_cavlc.cuh, hlb_bits.cuh, k_bits_* in hlb_slice.cu) and the multi-stream drop-in
(host/hlb200_glue.c batch mode + host/hl_b200_multi.c): the slice data written from the decision records must make the reference's own host code (headers, NAL
assembly, emulation prevention) emit the bitstream of the all-CPU reference, byte for byte -- for one stream and for many streams encoded by ONE launch per picture.

CPU tier: the same sources compiled as C++ stand in for the library (oracle/_ref/hl_glue_check_full, hl_multi_check; tools/emu/svc_shim.cpp).
GPU tier: the library itself (oracle/_ref/hl_b200_encoder is covered by tests/test_encoder.py::test_bitstream_md5_drop_in; hl_b200_multi here)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hl_ref_driver")
MULTI_CPU = os.path.join(ROOT, "oracle", "_ref", "hl_multi_check")
MULTI_GPU = os.path.join(ROOT, "oracle", "_ref", "hl_b200_multi")
GOLD = os.path.join(ROOT, "tests", "golden")


def _json(cmd):
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr[-800:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(MULTI_CPU), reason="oracle/_ref/hl_multi_check only exists where the reference tree is available")
def test_device_cavlc_fuzz_cpu():
    """random single-layer configurations through the whole glue with the device writer's source: bitstream MD5 against the live reference"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "fuzz_bits.py"), "8", "300"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and ", 0 mismatches" in out.stdout, out.stdout[-800:]


@pytest.mark.skipif(not os.path.exists(MULTI_CPU), reason="oracle/_ref/hl_multi_check only exists where the reference tree is available")
@pytest.mark.parametrize("args", [["--size", "176", "144", "--frames", "4", "--qp", "31", "--me-range", "16", "--gen", "g1"],
                                  ["--size", "96", "64", "--frames", "5", "--qp", "26", "--me-range", "32", "--gen", "g2", "--seed", "3"],
                                  ["--size", "96", "80", "--frames", "4", "--qp", "30", "--me-range", "16", "--gen", "g2", "--seed", "8", "--defaults"]])   # deblocking + early termination
def test_multi_stream_batch_cpu(args):
    """five codec instances (coroutines) encoding the same sequence in batch mode, two alternating groups: five identical bitstreams, equal to the reference's"""
    ref = _json([REF] + args)
    got = _json([MULTI_CPU] + args + ["--streams", "5", "--same-content"])
    assert got["all_streams_equal"] is True and (got["bytes"], got["md5"]) == (ref["bytes"], ref["md5"]), (got, ref)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(MULTI_GPU), reason="oracle/_ref/hl_b200_multi not built (needs the reference tree at build time)")
@pytest.mark.parametrize("name,streams", [("g1_cif_10", 6), ("g2_qcif", 9), ("g1_1080p_q31", 3), ("g2_qcif_deblock", 7), ("g1_1080p_defaults", 3)])
def test_multi_stream_drop_in(name, streams):
    """many streams through hl_codec_encode with ONE device launch per picture (encode + device-side CAVLC): every stream's bitstream equals the reference's"""
    g = np.load(os.path.join(GOLD, "encoder_%s.npz" % name))
    w, h, frames, qp, me_range, seed = (int(v) for v in g["config"])
    early, deblock = (int(g["flags"][0]), int(g["flags"][1])) if "flags" in g.files else (0, 0)
    args = ["--size", str(w), str(h), "--frames", str(frames), "--qp", str(qp), "--me-range", str(me_range), "--gen", str(g["gen"]), "--seed", str(seed), "--streams", str(streams),
            "--same-content", "--early-term", str(early), "--deblock", str(deblock)]
    got = _json([MULTI_GPU] + args)
    assert got["all_streams_equal"] is True and got["md5"] == str(g["bitstream_md5"]), (got, str(g["bitstream_md5"]))
