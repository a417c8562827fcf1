"""Generates tests/golden/svc_inter.npz: every enhancement-layer picture (I: I_BL macroblocks predicted from the resampled base layer,
rdo.c:301; P: base-mode inter macroblocks, rdo.c:1273) of three small multi-layer encodes of the UNMODIFIED reference
(oracle/_ref/hl_ref_driver --layers N --trace), per macroblock: inferred partitions / motion vectors, the ChromaAC/DC levels the macroblock
object held before the call, and what hl_codec_264_rdo_mb_guess_best_inter_pred_svc (rdo.c:1273) left: levels, coded-block patterns,
reconstructed samples -- plus the source and reference pictures.  Run in the container that has /root/reference (oracle/build_ref.sh first)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svc_util  # noqa: E402

if __name__ == "__main__":
    named = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, args in svc_util.CONFIGS:
            tr = os.path.join(tmp, name + ".trace")
            svc_util.run_driver_svc(args, tr)
            pics = svc_util.bl_pictures_from_trace(tr) + svc_util.pictures_from_trace(tr)
            named.append((name, pics))
            print(name, [(p["kind"], p["w"], p["h"], p["dqid"], int(p["valid"].sum()), len(p["valid"])) for p in pics])
    svc_util.save_golden(named)
    print(svc_util.GOLDEN, os.path.getsize(svc_util.GOLDEN), "bytes")
