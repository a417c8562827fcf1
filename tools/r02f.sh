mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
python -m pytest tests/ -q -m gpu > gpurun_out/r02f_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_tests_gpu.log
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2>> gpurun_out/r02f_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-all-inter --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1; echo "launch list rc=$?"
HBM_ONLY=tq_recon ncu --set full --clock-control none --import-source on -k regex:k_tq_recon -s 3 -c 1 -f -o gpurun_out/r02f_k_tq_recon python tools/hbm_kernels.py 32 > gpurun_out/r02f_ncu_tq.log 2>&1; echo "ncu tq rc=$?"
HBM_ONLY=svc_derive_motion ncu --set full --clock-control none --import-source on -k regex:k_svc_derive -s 6 -c 2 -f -o gpurun_out/r02f_k_svc_derive python tools/hbm_kernels.py 32 > gpurun_out/r02f_ncu_derive.log 2>&1; echo "ncu derive rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'])
print({k:v['frac_of_hbm_peak'] for k,v in d['hbm_kernels']['kernels'].items()})
print(d.get('svc_layers'))
r=json.loads(open('gpurun_out/r02f_bench_ref.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline']['cores'])"
