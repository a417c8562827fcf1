#!/usr/bin/env bash
# A/B on the GPU box: library variants x slice-kernel variant x stream count (prints value, ms/step)
set -u
IFS=';' read -ra LIST <<< "${SPECS:-libhl_b200.so cta 32;libhl_b200.so cta 256;libhl_b200.so warp 256}"
for spec in "${LIST[@]}"; do
  IFS=' ' read -r a b c <<< "$spec"; set -- $a $b $c
  HLB200_SLICE_KERNEL=$2 HLB200_LIB=$PWD/hartallo_b200/$1 timeout 400 python bench.py --streams $3 --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2 $3', 'value %.0f ms/step %.1f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e: print('$1 $2 $3', 'FAILED', e)
"
done
