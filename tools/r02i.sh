for v in cta warp; do for t in 0 1; do echo "== $v no_tma=$t"; HLB200_SLICE_KERNEL=$v HLB200_NO_TMA=$t timeout 120 python tools/dbg_golden.py g2_1080p_q31 3 2>&1 | tail -4; done; done
