mkdir -p gpurun_out
python tools/tma_probe.py 2>&1 | tail -18
bash tools/r02b.sh
