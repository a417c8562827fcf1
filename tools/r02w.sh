mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu > gpurun_out/r02w_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02w_tests_gpu.log
python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_bench_ref.json 2>> gpurun_out/r02w_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 2 --warmup 3 --no-all-inter --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02w_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_slice_encode_warp -s 4 -c 1 -f -o gpurun_out/r02w_prof python bench.py --steps 2 --warmup 3 --no-all-inter --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02w_ncu_full.log 2>&1; echo "ncu full rc=$?"
