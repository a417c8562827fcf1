mkdir -p gpurun_out
python tools/hbm_kernels.py 1 32 128 > gpurun_out/r02a_hbm.log 2>&1; echo "hbm rc=$?"; tail -3 gpurun_out/r02a_hbm.log | cut -c1-1500
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 64 > gpurun_out/r02a_laps_warp64.log 2>&1; echo "laps rc=$?"; cat gpurun_out/r02a_laps_warp64.log
python tools/hbm_kernels.py 32 > gpurun_out/plain_hbm.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_svc_inter_recon|k_svc_resample_intra|k_interp_luma|k_tq_recon" -s 8 -c 8 -f -o gpurun_out/r02a_prof_hbm python tools/hbm_kernels.py 32 > gpurun_out/r02a_ncu_hbm.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
