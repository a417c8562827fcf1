mkdir -p gpurun_out
for k in "cta-g1_1080p_q31" "warp-g1_1080p_q31" "cta-g2_1080p_q31" "warp-g2_1080p_q31" "cta-g1_1080p_ref4" "warp-g1_1080p_ref4"; do
  echo "== $k (TMA)"; timeout 100 python -m pytest tests/test_encoder.py -m gpu -x -q -k "test_slice_encode_vs_reference and $k" 2>&1 | tail -3
done
for k in "cta-g2_1080p_q31" "warp-g2_1080p_q31"; do
  echo "== $k (no TMA)"; HLB200_NO_TMA=1 timeout 100 python -m pytest tests/test_encoder.py -m gpu -x -q -k "test_slice_encode_vs_reference and $k" 2>&1 | tail -3
done
