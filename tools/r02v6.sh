mkdir -p gpurun_out
python -m pytest tests/test_codec_264_transf.py tests/test_codec_h264_interpol.py tests/test_batch_pictures.py tests/test_svc_inter.py tests/test_svc_bl_resample.py tests/test_codec_h264_pel.py -q -m gpu 2>&1 | tail -2
python tools/hbm_kernels.py 128 > gpurun_out/r02v6_hbm.jsonl 2>gpurun_out/r02v6_hbm.err
python -c "
import json
d=json.loads(open('gpurun_out/r02v6_hbm.jsonl').read().strip().splitlines()[-1]); print({k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()})"
