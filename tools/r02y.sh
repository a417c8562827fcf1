mkdir -p gpurun_out
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 256 > gpurun_out/r02y_laps_warp256.log 2>&1; echo "laps rc=$?"; tail -33 gpurun_out/r02y_laps_warp256.log | head -30
python bench.py --no-all-inter --no-hbm-kernels --no-cpu-baseline 2>gpurun_out/r02y.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'], d['step_ms'])"
