mkdir -p gpurun_out
HBM_ONLY=interp_luma ncu --set full --clock-control none --import-source on -k regex:k_interp_luma -s 3 -c 1 -f -o gpurun_out/r02v7_k_interp_luma python tools/hbm_kernels.py 32 > gpurun_out/r02v7_ncu_l.log 2>&1; echo "ncu luma rc=$?"
HBM_ONLY=interp_chroma ncu --set full --clock-control none --import-source on -k regex:k_interp_chroma -s 3 -c 1 -f -o gpurun_out/r02v7_k_interp_chroma python tools/hbm_kernels.py 32 > gpurun_out/r02v7_ncu_c.log 2>&1; echo "ncu chroma rc=$?"
