#!/usr/bin/env python3
"""tools/ncu_opmix.py <source-page.csv> [ntiles] — dynamic instruction mix of a kernel from `ncu --page source --csv`:
executed warp instructions per opcode (and per tile when ntiles is given), with the stall samples they collected."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
ntiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    if not m: continue
    op = m.group(2)
    try: n = int(r[iex]); s = int(r[ismp])
    except ValueError: continue
    ops[op] += n; smp[op] += s; tot += n
tsmp = sum(smp.values())
print(f"total executed warp instructions {tot}" + (f" = {tot / ntiles:.1f} per tile" if ntiles else ""))
for op, n in ops.most_common(40):
    line = f"{op:12s} {n:14d} {100.0 * n / tot:6.2f}%"
    if ntiles: line += f" {n / ntiles:9.1f}/tile"
    line += f"   samples {100.0 * smp[op] / max(tsmp, 1):5.1f}%"
    print(line)
