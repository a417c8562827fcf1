mkdir -p gpurun_out
python -m pytest tests/test_codec_h264_interpol.py tests/test_codec_264_transf.py tests/test_batch_pictures.py tests/test_codec_h264_pel.py -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/hbm_kernels.py 32 128 > gpurun_out/r02n_hbm.log 2>&1; python - <<'PY'
import json
for ln in open('gpurun_out/r02n_hbm.log'):
    try: d=json.loads(ln)
    except: print(ln[:300]); continue
    print(d['pictures_per_launch'], {k:(v['ms'],v['achieved_gbs'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()})
PY
