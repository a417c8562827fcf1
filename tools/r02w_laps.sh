mkdir -p gpurun_out
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 64 > gpurun_out/r02w_laps_warp64.log 2>&1; echo "laps rc=$?"; cat gpurun_out/r02w_laps_warp64.log
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 256 > gpurun_out/r02w_laps_warp256.log 2>&1; echo "laps rc=$?"; tail -34 gpurun_out/r02w_laps_warp256.log
