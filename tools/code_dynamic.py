#!/usr/bin/env python
"""Executed warp instructions of a kernel attributed to source functions: joins `ncu --page source --csv` (per-instruction counts) with `nvdisasm -g -c`
(line markers) of THE SAME build.  python tools/code_dynamic.py <ncu_source.csv> <dis.txt> <kernel substring> [n_macroblocks]"""
import bisect
import collections
import csv
import os
import re
import sys

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
nmb = float(sys.argv[4]) if len(sys.argv) > 4 else None
rx_fn = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:HLB_HD|HLB_FN|HLB_CAVLC_FN|HLB_INTERP_FN|HLB_FASTPRED_FN|__device__|__global__|static|inline)[^;=]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$')
funcs = {}


def load(path):
    if path in funcs or not os.path.exists(path):
        return
    out = []
    for i, l in enumerate(open(path, errors="ignore"), 1):
        m = rx_fn.match(l)
        if m and not l.strip().startswith("//") and m.group(1) not in ("if", "for", "while", "switch", "return", "defined", "__launch_bounds__"):
            out.append((i, m.group(1)))
    funcs[path] = out


owner = {}   # offset -> (file:function, line)
inside = False
cur = ("?", 0)
for l in open(dis):
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.search(r'/\*([0-9a-f]{4,6})\*/', l)
    if m:
        path, line = cur
        load(path)
        fl = funcs.get(path, [])
        k = bisect.bisect_right([a for a, _ in fl], line) - 1
        owner[int(m.group(1), 16)] = ((os.path.basename(path) + ":" + fl[k][1]) if k >= 0 else os.path.basename(path) + ":?", line)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie, ia = hdr.index("Instructions Executed"), hdr.index("Address")
data = [(int(r[ia], 16), int(r[ie])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
base = data[0][0]
tot = sum(e for _, e in data)
byfn = collections.Counter()
byline = collections.Counter()
for a, e in data:
    fn, line = owner.get(a - base, ("?", 0))
    byfn[fn] += e
    byline[(fn, line)] += e
print("%s: %d static, %.3g executed warp instructions%s" % (kern, len(data), tot, (" = %.0f per macroblock" % (tot / nmb)) if nmb else ""))
for fn, e in byfn.most_common(45):
    print("%6.2f%%  %s%s" % (100.0 * e / tot, fn, ("  (%.0f / MB)" % (e / nmb)) if nmb else ""))
print("-- hottest source lines")
for (fn, line), e in byline.most_common(25):
    print("%6.2f%%  %s:%d" % (100.0 * e / tot, fn, line))
