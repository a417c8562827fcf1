mkdir -p gpurun_out
python -m pytest tests/test_fast_prims.py tests/test_encoder.py -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02b_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
HLB200_SLICE_KERNEL=warp HLB200_LIB=$PWD/hartallo_b200/libhl_b200_prof.so python tools/mb_timeline.py 64 > gpurun_out/r02b_laps_warp64.log 2>&1; echo "laps rc=$?"; cat gpurun_out/r02b_laps_warp64.log
