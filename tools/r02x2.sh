mkdir -p gpurun_out
python -m pytest tests/test_svc_derive.py tests/test_svc_inter.py tests/test_svc_bl_resample.py tests/test_codec_264_transf.py tests/test_codec_h264_interpol.py tests/test_batch_pictures.py -q -m gpu > gpurun_out/r02x2_tests_svc.log 2>&1; echo "svc+batch tests rc=$?"; tail -6 gpurun_out/r02x2_tests_svc.log
A="--layers 3 --size 256 128 --frames 4 --gen g1 --seed 6426 --qp 25 --scale 3 2"
oracle/_ref/hl_ref_driver $A 2>/dev/null | tail -1 | cut -c1-200
oracle/_ref/hl_b200_encoder $A 2>gpurun_out/r02x2_ess.err | tail -1 | cut -c1-200
