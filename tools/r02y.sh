mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu > gpurun_out/r02y_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02y_tests_gpu.log
python bench.py --distinct 256 --steps 4 --warmup 3 --no-all-inter --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02y_bench_distinct256.json 2> gpurun_out/r02y_bench_distinct256.err; echo "bench distinct rc=$?"; tail -c 400 gpurun_out/r02y_bench_distinct256.err
python -c "
import json
d=json.loads(open('gpurun_out/r02y_bench_distinct256.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'])"
