#!/usr/bin/env python
"""Summarises gpurun_out/ ncu artefacts into profiles/<tag>_*.{csv,md} (tracked).  Usage: tools/summarise_profiles.py r01a"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)
out = ["# ncu summary %s" % tag, ""]

# ---- launch list: per-kernel totals and shares ----
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    h = rows[0]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    mu = h.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)  # -> us
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 3 --warmup 3)", "",
            "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append("| %s | %d | %.1f | %.2f | %.1f%% |" % (name, n, t, t / n, 100 * t / tot))
            f.write("%s,%d,%.2f,%.3f,%.4f\n" % (name, n, t, t / n, t / tot))
    out.append("")

# ---- full capture: chosen raw metrics per kernel ----
rp = os.path.join(G, "prof.ncu-rep")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum", "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "derived__smsp__sass_thread_inst_executed_op_integer_pred_on.sum", "smsp__sass_thread_inst_executed_op_integer_pred_on.sum"]
if os.path.exists(rp):
    txt = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    cols = [h.index(w) for w in WANT if w in h]
    out += ["## full capture (`ncu --set full --clock-control none --import-source on`)", ""]
    with open(os.path.join(P, tag + "_ncu_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [h[c] + " [" + units[c] + "]" for c in cols])
        for r in rows[2:]:
            name = r[kn].split("(")[0]
            w.writerow([name] + [r[c] for c in cols])
            out.append("### " + name)
            out += ["- %s = %s %s" % (h[c], r[c], units[c]) for c in cols]
            out.append("")
bp = os.path.join(G, "bench.json")
if os.path.exists(bp) and os.path.getsize(bp):
    out += ["## bench.py line of the same build (not under ncu)", "", "```json", open(bp).read().strip().splitlines()[-1], "```", ""]
open(os.path.join(P, tag + "_summary.md"), "w").write("\n".join(out))
print("\n".join(out[:60]))
