mkdir -p gpurun_out
for k in k_tq_recon k_interp_chroma k_svc_inter_recon k_svc_resample_intra; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02r_$k python tools/hbm_kernels.py 32 > gpurun_out/r02r_$k.log 2>&1; echo "$k rc=$?"
done
