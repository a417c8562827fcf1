mkdir -p gpurun_out
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 256 > gpurun_out/r02_laps_warp256.log 2>&1; tail -33 gpurun_out/r02_laps_warp256.log | head -31
