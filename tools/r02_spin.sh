for ns in 32768 65536 262144; do
HLB200_LIB=$PWD/hartallo_b200/libhl_b200_s$ns.so python bench.py --no-all-inter --no-hbm-kernels --no-cpu-baseline --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spin', $ns, 'value',d['value'], [round(x) for x in d['step_ms']])"
done
