mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02m_bench.json; tail -5 gpurun_out/r02m_bench.err
