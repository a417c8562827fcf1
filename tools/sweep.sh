#!/usr/bin/env bash
# tuning sweep on the GPU box: bench.py over library variants x stream counts (prints value, ms/step, e2e)
set -u
python -m pytest tests/test_encoder.py -x -q -m gpu -k "vs_reference or 1080p" 2>&1 | tail -2
for lib in ${LIBS:-libhl_b200.so}; do
  for s in ${STREAMS:-32 128}; do
    HLB200_LIB=$PWD/hartallo_b200/$lib python bench.py --streams $s --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', $s, 'value %.0f ms/step %.1f e2e %.0f frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
except Exception as e: print('$lib', $s, 'FAILED', e)
"
  done
done
