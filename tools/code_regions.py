#!/usr/bin/env python
"""Out-of-line regions of a kernel (between CALL.REL.NOINC targets) with their size and the source functions that make them up: python tools/code_regions.py <dis.txt> <kernel substring>"""
import bisect
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
dis, kern = sys.argv[1], sys.argv[2]
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


ins = []  # (addr, srcfunc, text)
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
    m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m:
        path, line = cur
        load(path)
        fl = funcs.get(path, [])
        k = bisect.bisect_right([a for a, _ in fl], line) - 1
        ins.append((int(m.group(1), 16), (os.path.basename(path) + ":" + fl[k][1]) if k >= 0 else os.path.basename(path) + ":?", m.group(2)))
targets = set()
labels = {}
for l in open(dis):
    pass
# call targets appear as labels in -c output: `CALL.REL.NOINC `(label)`; resolve label addresses from label lines
lab_addr = {}
inside = False
last_addr = 0
pending = []
for l in open(dis):
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.match(r'^(\.L_x_\d+|\$[^:]+):', l.strip())
    if m:
        pending.append(m.group(1))
        continue
    m = re.search(r'/\*([0-9a-f]{4,6})\*/', l)
    if m:
        a = int(m.group(1), 16)
        for p in pending:
            lab_addr[p] = a
        pending = []
for a, f, t in ins:
    if t.startswith("CALL"):
        m = re.search(r'`\(([^)]+)\)', t)
        if m and m.group(1) in lab_addr:
            targets.add(lab_addr[m.group(1)])
bounds = sorted(targets)
print("%s: %d instructions, %d out-of-line regions" % (kern, len(ins), len(bounds)))
regions = collections.defaultdict(collections.Counter)
for a, f, t in ins:
    k = bisect.bisect_right(bounds, a) - 1
    regions[bounds[k] if k >= 0 else 0][f] += 16
for start in sorted(regions, key=lambda s: -sum(regions[s].values())):
    c = regions[start]
    print("%6x %7d B  %s" % (start, sum(c.values()), ", ".join("%s %d" % (n, b) for n, b in c.most_common(5))))
