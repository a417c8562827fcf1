mkdir -p gpurun_out
python -m pytest tests/test_svc_derive.py tests/test_svc_inter.py tests/test_svc_bl_resample.py tests/test_abi.py -q -m gpu > gpurun_out/r02x_tests_svc.log 2>&1; echo "svc tests rc=$?"; tail -5 gpurun_out/r02x_tests_svc.log
python tools/hbm_kernels.py 1 32 128 > gpurun_out/r02x_hbm.jsonl 2> gpurun_out/r02x_hbm.err; echo "hbm rc=$?"; tail -c 600 gpurun_out/r02x_hbm.err
A="--layers 3 --size 176 144 --frames 6 --gen g1"
oracle/_ref/hl_ref_driver $A --out /tmp/r.264 2>/dev/null | tail -1 > gpurun_out/r02x_svc_ref.json
for i in 1 2 3; do oracle/_ref/hl_b200_encoder $A --out /tmp/b.264 2>gpurun_out/r02x_svc_b200.err | tail -1 >> gpurun_out/r02x_svc_b200.json; done
cat gpurun_out/r02x_svc_ref.json gpurun_out/r02x_svc_b200.json | cut -c1-300
