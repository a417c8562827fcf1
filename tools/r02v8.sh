mkdir -p gpurun_out
python -m pytest tests/test_codec_h264_interpol.py tests/test_batch_pictures.py tests/test_fast_prims.py -q -m gpu 2>&1 | tail -2
HBM_ONLY=interp_chroma python tools/hbm_kernels.py 128 > gpurun_out/r02v8_chroma.log 2>&1; grep -o "\"interp_chroma\": {[^}]*}" gpurun_out/r02v8_chroma.log
