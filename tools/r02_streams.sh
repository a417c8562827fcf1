mkdir -p gpurun_out
for s in 384 512; do
python bench.py --streams $s --no-all-inter --no-hbm-kernels --no-cpu-baseline --e2e-groups 1 2>gpurun_out/r02z.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', $s, 'value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked'],'frac',d['roofline']['frac'], d['step_ms'])"
done
