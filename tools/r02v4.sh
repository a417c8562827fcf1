mkdir -p gpurun_out
python -m pytest tests/test_svc_bl_resample.py tests/test_svc_inter.py -q -m gpu 2>&1 | tail -2
HBM_ONLY=svc_resample_intra python tools/hbm_kernels.py 32 128 > gpurun_out/r02v4_resample.log 2>&1; grep -o "\"pictures_per_launch\": [0-9]*\|\"svc_resample_intra\": {[^}]*}" gpurun_out/r02v4_resample.log
