for v in cta warp; do for t in 0 1; do echo "== $v no_tma=$t 352x288 g2"; HLB200_SLICE_KERNEL=$v HLB200_NO_TMA=$t timeout 60 python tools/dbg1080.py 352 288 g2 3 2>&1 | tail -6; done; done
for v in cta warp; do for t in 0 1; do echo "== $v no_tma=$t 1080p g1"; HLB200_SLICE_KERNEL=$v HLB200_NO_TMA=$t timeout 60 python tools/dbg1080.py 1920 1088 g1 3 2>&1 | tail -6; done; done
