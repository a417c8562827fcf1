mkdir -p gpurun_out
HLB200_DEVICE=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice_encode -c 40 --csv --log-file gpurun_out/r02_e2e_launches.csv oracle/_ref/hl_b200_multi --streams 512 --groups 2 --frames 12 --warmup 3 --distinct 16 > gpurun_out/r02_e2e_ncu.log 2>&1; echo rc=$?
