mkdir -p gpurun_out
echo "== no TMA"; HLB200_NO_TMA=1 python -m pytest tests/test_encoder.py -m gpu -x -q -k "g2_small or g1_qcif" 2>&1 | tail -4
echo "== TMA"; python -m pytest tests/test_encoder.py -m gpu -x -q -k "g2_small" 2>&1 | tail -4
echo "== sanitizer"; HLB200_SLICE_KERNEL=warp compute-sanitizer --tool memcheck python -m pytest tests/test_encoder.py -m gpu -x -q -k "warp-g2_small" > gpurun_out/r02c_sanitize.log 2>&1; grep -E "Illegal|Invalid|at 0x|by thread|Address|ERROR SUMMARY" gpurun_out/r02c_sanitize.log | head -30
