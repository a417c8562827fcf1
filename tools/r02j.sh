mkdir -p gpurun_out
B="python bench.py --streams 192 --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_slice_encode" -s 4 -c 1 -f -o gpurun_out/r02j_prof $B > gpurun_out/r02j_ncu.log 2>&1
echo "ncu rc=$?"
