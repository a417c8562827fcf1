mkdir -p gpurun_out
python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02t_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02t_bench_ref.json 2>> gpurun_out/r02t_bench.err; echo "ref rc=$?"
