mkdir -p gpurun_out
: > gpurun_out/r02v_tq_variants.log
for L in "" hartallo_b200/variants/tq12.so hartallo_b200/variants/tq16.so ""; do
  echo "lib=$L" >> gpurun_out/r02v_tq_variants.log
  HLB200_LIB=${L:+$PWD/$L} HBM_ONLY=tq_recon python tools/hbm_kernels.py 128 >> gpurun_out/r02v_tq_variants.log 2>&1
done
grep -o "lib=.*\|\"tq_recon\": {[^}]*}" gpurun_out/r02v_tq_variants.log
