#!/usr/bin/env bash
# A/B on the GPU box: encoder parity tests, then bench.py for each slice-kernel variant x stream count (value, ms/step, e2e)
set -u
python -m pytest tests/test_encoder.py -x -q -m gpu 2>&1 | tail -3
for v in ${VARIANTS:-cta warp}; do
  for s in ${STREAMS:-128 256}; do
    HLB200_SLICE_KERNEL=$v timeout 600 python bench.py --streams $s --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', $s, 'value %.0f ms/step %.1f e2e %.0f frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
except Exception as e: print('$v', $s, 'FAILED', e)
"
  done
done
