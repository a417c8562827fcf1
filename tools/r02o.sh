mkdir -p gpurun_out
python tools/hbm_kernels.py 32 > gpurun_out/plain_hbm.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_interp_luma|k_tq_recon|k_interp_chroma" -s 9 -c 3 -f -o gpurun_out/r02o_prof_hbm python tools/hbm_kernels.py 32 > gpurun_out/r02o_ncu_hbm.log 2>&1; echo "ncu rc=$?"
