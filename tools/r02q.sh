mkdir -p gpurun_out
for a in "--deblock 1" "--early-term 1"; do HLB200_DEVICE=0 oracle/_ref/hl_b200_multi --streams 64 --frames 5 --warmup 1 --groups 1 $a 2>&1 | tail -1 | cut -c1-330; done
HLB200_DEVICE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02q_launches.csv oracle/_ref/hl_b200_multi --streams 64 --frames 4 --warmup 1 --groups 1 --defaults > gpurun_out/r02q_ncu.log 2>&1; echo "ncu rc=$?"
