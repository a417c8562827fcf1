for s in 256 384 512 768; do
  timeout 400 python bench.py --streams $s --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $s', 'value %.0f ms/step %.1f e2e %.0f frac %.4f step_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], [round(x) for x in d['step_ms']]))
except Exception as e: print('streams $s FAILED', e)
"
done
