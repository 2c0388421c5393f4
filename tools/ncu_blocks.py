#!/usr/bin/env python3
"""tools/ncu_blocks.py <report.ncu-rep> <tiles> — executed instructions and stall samples per basic block (SASS level)."""
import csv, subprocess, sys
from collections import Counter
rep, tiles = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ismp]) for r in data)
print(f"total inst {tot}  ({tot / tiles:.0f} per tile)   samples {tots}")
def op(r):
    t = r[isrc].split()
    return t[1] if t[0].startswith('@') else t[0]
blocks, cur = [], None
for i, r in enumerate(data):
    n = int(r[ia])
    if cur and abs(n - cur['n']) <= 0.02 * max(n, cur['n'], 1):
        cur['rows'].append(r)
    else:
        cur = {'start': i, 'n': n, 'rows': [r]}
        blocks.append(cur)
for b in blocks:
    s = sum(int(r[ia]) for r in b['rows']); smp = sum(int(r[ismp]) for r in b['rows'])
    if s > 0.004 * tot or smp > 0.01 * tots:
        c = Counter(op(r).split('.')[0] for r in b['rows'] if any(k in op(r) for k in ('ATOM', 'LDG', 'STG', 'LDS', 'STS', 'VOTE', 'SHFL', 'NANOSLEEP', 'CALL', 'RET', 'LD.', 'ST.', 'RED', 'POPC', 'WARPSYNC')))
        st = Counter()
        for r in b['rows']:
            for i, h in stall_cols:
                st[h] += int(r[i] or 0)
        top = ", ".join(f"{h[6:]}:{100 * v / max(smp, 1):.0f}%" for h, v in st.most_common(3))
        print(f"@{b['start']:5d} len {len(b['rows']):4d} exec/tile {b['n'] / tiles:8.2f} inst% {100 * s / tot:5.1f} smp% {100 * smp / tots:5.1f} [{top}] {dict(c)}")
