"""Generates tests/golden/svc_bitstream.json: size and MD5 of the multi-layer (SVC, dyadic spatial) bitstreams the UNMODIFIED reference produces
(oracle/_ref/hl_ref_driver --layers N, the counterpart of source/test_encoder.c:150-202) on configurations the drop-in accepts (it refuses enhancement-layer I pictures of fewer than 36 macroblocks and macroblocks that would be coded against scratch memory of an
earlier picture; tests/test_svc_inter.py checks the refusals)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reftrace  # noqa: E402

CONFIGS = {
    "g1_qcif_cif": ["--size", "176", "144", "--layers", "2", "--frames", "3", "--gen", "g1"],                    # lower two layers of BASELINE.json configs[3]
    "g1_qcif_cif_4cif": ["--size", "176", "144", "--layers", "3", "--frames", "2", "--gen", "g1"],              # configs[3]: QCIF -> CIF -> 4CIF
    "g1_3layer_small_q24": ["--size", "64", "48", "--layers", "3", "--frames", "4", "--gen", "g1", "--qp", "24"],
    "g1_2layer_q38": ["--size", "96", "80", "--layers", "2", "--frames", "4", "--gen", "g1", "--qp", "38"],
    "g2_qcif_cif": ["--size", "176", "144", "--layers", "2", "--frames", "4", "--gen", "g2"],                   # has macroblocks with an inherited prediction
    # Intra4x4 macroblocks inside base-layer P pictures: the enhancement layers derive motion from the lower-case vector such a macroblock kept from its last
    # inter commit (host/hlb200_glue.c: glue_apply); both differed through the whole glue before that was reproduced
    "g2_i4_in_p_2layer": ["--size", "64", "64", "--layers", "2", "--frames", "4", "--gen", "g2", "--seed", "8931", "--qp", "29"],
    "g2_3layer_48_q34": ["--size", "48", "48", "--layers", "3", "--frames", "4", "--gen", "g2", "--seed", "22", "--qp", "34"],   # smallest layers the drop-in accepts (36 macroblocks at the dyadic ratio)
    # extended spatial scalability: every layer 1.5 times the one below (ref_driver --scale 3 2) -- the general case of the inter-layer derivation, enhancement
    # macroblocks with sub-macroblock partitions, non-dyadic Intra_Base resampling
    "g2_ess_3layer": ["--size", "128", "192", "--layers", "3", "--frames", "4", "--gen", "g2", "--seed", "6302", "--qp", "35", "--scale", "3", "2"],
    "g1_ess_3layer": ["--size", "256", "128", "--layers", "3", "--frames", "4", "--gen", "g1", "--seed", "6426", "--qp", "25", "--scale", "3", "2"],
    "g2_ess_2layer": ["--size", "128", "64", "--layers", "2", "--frames", "2", "--gen", "g2", "--seed", "6675", "--qp", "23", "--scale", "3", "2"],
}
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svc_bitstream.json")

if __name__ == "__main__":
    res = {}
    for name, args in CONFIGS.items():
        runs = []
        for _ in range(3):   # three runs each: the entry is only kept if the reference agrees with itself
            o = subprocess.run([reftrace.DRIVER] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
            j = json.loads(o.stdout.strip().splitlines()[-1])
            runs.append((j["bytes"], j["md5"]))
        assert len(set(runs)) == 1, (name, runs)
        res[name] = {"args": args, "bytes": runs[0][0], "md5": runs[0][1]}
        print(name, runs[0])
    json.dump(res, open(OUT, "w"), indent=1)
