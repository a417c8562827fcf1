mkdir -p gpurun_out
python -m pytest tests/test_encoder.py -x -q -m gpu > gpurun_out/r02p_enc.log 2>&1; echo "enc rc=$?"; tail -5 gpurun_out/r02p_enc.log
python -m pytest tests/test_svc_inter.py tests/test_bits.py -x -q -m gpu > gpurun_out/r02p_svc.log 2>&1; echo "svc rc=$?"; tail -3 gpurun_out/r02p_svc.log
for a in "" "--defaults"; do HLB200_DEVICE=0 oracle/_ref/hl_b200_multi --streams 64 --frames 5 --warmup 1 --groups 1 $a 2>&1 | tail -1 | cut -c1-330; done
