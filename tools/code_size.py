#!/usr/bin/env python
"""Code bytes of a kernel attributed to the source functions they were compiled from (nvdisasm -g line markers): python tools/code_size.py <dis.txt> <kernel substring>
Used to find what makes up the instruction footprint of the slice kernel (instruction-fetch bound)."""
import bisect
import collections
import os
import re
import sys

dis, kern = sys.argv[1], sys.argv[2]
funcs = {}  # file -> sorted [(line, name)]


def load(path):
    if path in funcs or not os.path.exists(path):
        return
    out = []
    rx = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:HLB_HD|HLB_FN|HLB_CAVLC_FN|HLB_INTERP_FN|HLB_FASTPRED_FN|__device__|__global__|static|inline)[^;=]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$')
    for i, l in enumerate(open(path, errors="ignore"), 1):
        m = rx.match(l)
        if m and not l.strip().startswith("//") and m.group(1) not in ("if", "for", "while", "switch", "return", "defined", "__launch_bounds__"):
            out.append((i, m.group(1)))
    funcs[path] = out


size = collections.Counter()
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
    if re.search(r'/\*[0-9a-f]{4,6}\*/', l):
        path, line = cur
        load(path)
        fl = funcs.get(path, [])
        k = bisect.bisect_right([a for a, _ in fl], line) - 1
        name = fl[k][1] if k >= 0 else "?"
        size[(os.path.basename(path), name)] += 16
tot = sum(size.values())
print("%s: %d bytes" % (kern, tot))
for (f, n), b in size.most_common(60):
    print("%7d  %5.1f%%  %s:%s" % (b, 100.0 * b / tot, f, n))
