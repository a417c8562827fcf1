HLB200_DEVICE=0 oracle/_ref/hl_b200_multi --streams 512 --groups 2 --frames 7 --warmup 2 --distinct 16 2>&1 | tail -1 | cut -c1-600
