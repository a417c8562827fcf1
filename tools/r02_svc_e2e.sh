for a in "--layers 3 --size 176 144 --frames 6 --gen g1" "--layers 3 --size 176 144 --frames 6 --gen g2 --seed 3"; do
echo "== $a"
oracle/_ref/hl_ref_driver $a --out /tmp/r.264 2>/dev/null | tail -1 | cut -c1-260
HLB200_DEVICE=0 oracle/_ref/hl_b200_encoder $a --out /tmp/g.264 2>/dev/null | tail -1 | cut -c1-260
cmp /tmp/r.264 /tmp/g.264 && echo SAME
done
