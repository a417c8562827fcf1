#!/usr/bin/env python
"""Key metrics of every kernel of an .ncu-rep as a small markdown table (profiles/*.md): python tools/ncu_summary.py capture.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe"),
    ("sm__icc_request_hit_rate.pct", "SM instruction cache hit rate"), ("sm__icc_requests.sum.pct_of_peak_sustained_elapsed", "SM instruction cache requests (of peak)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping (idle pollers)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (of peak)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print("# %s\n" % (sys.argv[2] if len(sys.argv) > 2 else rep))
print("`ncu --set full --clock-control none --import-source on`, read with `ncu -i %s --page raw --csv` (tools/ncu_summary.py); per-launch values of a cold, serialised replay.\n" % rep.split("/")[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("## %s\n" % d.get("Kernel Name", "?").split("(")[0])
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d and d[k] != "":
            v = d[k]
            try:
                f = float(v)
                v = ("%.0f" % f) if abs(f) >= 1000 else ("%.3g" % f)
            except ValueError:
                pass
            print("| %s (`%s`) | %s %s |" % (label, k, v, units[hdr.index(k)]))
    print()
