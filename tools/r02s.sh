mkdir -p gpurun_out
for cfg in "256 1" "512 2" "768 3" "512 1"; do set -- $cfg; HLB200_DEVICE=0 oracle/_ref/hl_b200_multi --streams $1 --groups $2 --frames 7 --warmup 2 --distinct 16 2>&1 | tail -1 | cut -c1-260; done
