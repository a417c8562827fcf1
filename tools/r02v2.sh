mkdir -p gpurun_out
: > gpurun_out/r02v2_variants.log
for L in default va vb; do
  echo "lib=$L" >> gpurun_out/r02v2_variants.log
  if [ "$L" = default ]; then python tools/hbm_kernels.py 128 >> gpurun_out/r02v2_variants.log 2>&1; else HLB200_LIB=$PWD/hartallo_b200/variants/$L.so python tools/hbm_kernels.py 128 >> gpurun_out/r02v2_variants.log 2>&1; fi
done
python - <<'P'
import json
for l in open('gpurun_out/r02v2_variants.log'):
    if l.startswith('lib='): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print({k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()})
    else: print(l.strip()[:200])
P
