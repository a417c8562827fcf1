#!/usr/bin/env bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the ncu launch list of the same command and one full capture of the
# slice kernel.  Outputs land in gpurun_out/ (scratch); summaries worth keeping are copied into profiles/ by tools/summarise_profiles.py.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 900 gpurun_out/bench.json
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# full capture: fewer streams (ncu saves and restores every byte the kernel touches between its ~45 replay passes), same kernel variant
HLB200_SLICE_KERNEL=${NCU_VARIANT:-warp} ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-k_slice_encode}" -s ${NCU_SKIP:-4} -c ${NCU_COUNT:-1} -f -o gpurun_out/prof \
  python bench.py --streams ${NCU_STREAMS:-96} --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
