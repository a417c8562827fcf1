mkdir -p gpurun_out
HBM_ONLY=svc_resample_intra ncu --set full --clock-control none --import-source on -k regex:k_svc_resample_intra -s 3 -c 1 -f -o gpurun_out/r02v5_k_svc_resample python tools/hbm_kernels.py 32 > gpurun_out/r02v5_ncu.log 2>&1; echo "ncu rc=$?"
