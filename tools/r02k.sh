mkdir -p gpurun_out
python -m pytest tests/test_bits.py tests/test_encoder.py tests/test_svc_inter.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02k_pytest.log
for cfg in "64 2" "128 2" "256 2" "256 1" "256 4"; do set -- $cfg; echo "== streams $1 groups $2"; timeout 300 oracle/_ref/hl_b200_multi --streams $1 --frames 7 --warmup 2 --groups $2 2>&1 | tail -1 | cut -c1-420; done
nvidia-smi --query-gpu=memory.used --format=csv | tail -1; free -g | head -2; nproc
