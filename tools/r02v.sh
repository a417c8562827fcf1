python tools/hbm_kernels.py 128 2>&1 | tail -1
python -m pytest tests/test_codec_264_transf.py tests/test_batch_pictures.py -x -q -m gpu 2>&1 | tail -2
