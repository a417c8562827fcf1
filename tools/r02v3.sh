mkdir -p gpurun_out
: > gpurun_out/r02v3_interp_order.log
for L in default raster default raster; do
  echo "lib=$L" >> gpurun_out/r02v3_interp_order.log
  if [ "$L" = default ]; then HBM_ONLY=interp_luma python tools/hbm_kernels.py 128 >> gpurun_out/r02v3_interp_order.log 2>&1; else HLB200_LIB=$PWD/hartallo_b200/variants/$L.so HBM_ONLY=interp_luma python tools/hbm_kernels.py 128 >> gpurun_out/r02v3_interp_order.log 2>&1; fi
done
grep -o "lib=.*\|\"interp_luma\": {[^}]*}" gpurun_out/r02v3_interp_order.log
python -m pytest tests/test_codec_h264_interpol.py tests/test_batch_pictures.py -q -m gpu 2>&1 | tail -2
