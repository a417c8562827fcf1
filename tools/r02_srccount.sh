mkdir -p gpurun_out
ncu --section SourceCounters --section InstructionStats --clock-control none --import-source on -k regex:k_slice_encode_warp -s 4 -c 1 -f -o gpurun_out/r02x_src python bench.py --steps 2 --warmup 3 --no-all-inter --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02x_src.log 2>&1; echo "rc=$?"
