mkdir -p gpurun_out
python -m pytest tests/test_batch_pictures.py tests/test_codec_264_transf.py tests/test_codec_h264_interpol.py tests/test_codec_h264_pel.py tests/test_fast_prims.py -x -q -m gpu 2>&1 | tail -4
for k in k_interp_luma k_interp_chroma k_tq_recon; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02u_$k python tools/hbm_kernels.py 32 > gpurun_out/r02u_$k.log 2>&1; echo "$k rc=$?"
done
