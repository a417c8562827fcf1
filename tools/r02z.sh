mkdir -p gpurun_out
python -m pytest tests/test_encoder.py -x -q -m gpu -k "g2_qcif or g1_qcif or g3_cif_defaults or g2_small or g2_1080p_q31 or cif_q38" 2>&1 | tail -3
python bench.py --no-all-inter --no-hbm-kernels --no-cpu-baseline 2>gpurun_out/r02z.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'], d['step_ms'])"
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 256 > gpurun_out/r02z_laps_warp256.log 2>&1; tail -33 gpurun_out/r02z_laps_warp256.log | grep -E "pskip:|inter:|intra|pskip chroma|trial run|search ctl"
