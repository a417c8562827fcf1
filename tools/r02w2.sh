mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r02w_smoke.log 2>&1; tail -1 gpurun_out/r02w_smoke.log
bash tools/r02w.sh
