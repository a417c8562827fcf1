mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -x -q -m gpu -k "g2_qcif or g2_small or g2_1080p_q31 or gpu_fuzz or g1_cif_10 or ref4 or large_batch" 2>&1 | tail -2
timeout 600 python bench.py --no-all-inter --no-hbm-kernels --no-cpu-baseline 2>gpurun_out/r02z.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'], d['step_ms'])"
