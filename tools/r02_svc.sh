python -m pytest tests/test_svc_inter.py tests/test_svc_bl_resample.py -x -q -m gpu 2>&1 | tail -8
python tools/hbm_kernels.py 128 2>&1 | tail -1
