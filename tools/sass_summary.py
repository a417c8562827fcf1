#!/usr/bin/env python
"""Per-kernel static instruction counts and the mnemonics that prove the packed-byte / TMA formulation, from `cuobjdump -sass` on stdin (profiles/r02_sass.txt)."""
import collections
import re
import subprocess
import sys

PATS = ['UTMALDG', 'SYNCS', 'VABSDIFF4', 'IDP.4A', 'I2IP', 'VIADDMNMX', 'PRMT', 'SHF.R.W', 'LDG.E.U8', 'LDS', 'STS', 'LDG', 'STG', 'DADD', 'DMUL', 'DSETP', 'SHFL', 'VOTE', 'REDUX', 'BAR', 'NANOSLEEP']
for f in re.split(r'\n\s*Function : ', sys.stdin.read())[1:]:
    name = f.split('\n', 1)[0].strip()
    lines = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4,6}\*/', l)]
    cnt = collections.Counter()
    for l in lines:
        for p in PATS:
            if re.search(r'\b' + re.escape(p), l):
                cnt[p] += 1
    short = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip().split('(')[0]
    print("%-40s %6d instructions (%5.1f KB)  " % (short[:40], len(lines), len(lines) * 16 / 1024) + "  ".join("%s %d" % (k, cnt[k]) for k in PATS if cnt[k]))
