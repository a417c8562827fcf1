mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r02g_smoke.log 2>&1; tail -1 gpurun_out/r02g_smoke.log
python -m pytest tests/ -q -m gpu > gpurun_out/r02g_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02g_tests_gpu.log
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02g_bench_ref.json 2>> gpurun_out/r02g_bench.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'])
print({k:v['frac_of_hbm_peak'] for k,v in d['hbm_kernels']['kernels'].items()})
print(d.get('svc_layers'))
r=json.loads(open('gpurun_out/r02g_bench_ref.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline']['cores'])"
