"""Aggregates the SASS source page of one kernel of an .ncu-rep per opcode: stall samples, executed warp instructions, shared
memory wavefronts (real vs ideal).  usage: ncu_sass_ops.py report.ncu-rep kernel_regex"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', f'regex:{sys.argv[2]}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Address')
iS, iI, iW, iWi = (hdr.index(k) for k in ('# Samples', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal'))
seen, data = set(), []
for r in rows:
    if len(r) != len(hdr) or r[0] == 'Address' or r[0] in seen:
        continue                      # second launch of the same kernel repeats the addresses
    seen.add(r[0])
    data.append(r)
tot = sum(int(r[iS]) for r in data)
g = defaultdict(lambda: [0, 0, 0, 0])
for r in data:
    w = r[1].split()
    op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
    a = g[op]
    a[0] += int(r[iS]); a[1] += int(r[iI]); a[2] += int(r[iW]); a[3] += int(r[iWi])
print(f'# {sys.argv[2]}: {len(data)} SASS instructions, {tot} stall samples, {sum(a[1] for a in g.values())} warp instructions executed')
for op, a in sorted(g.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f'{op:10s} samples {100 * a[0] / max(tot, 1):5.1f}%  executed {a[1]:9d}  smem wavefronts {a[2]:9d} (ideal {a[3]:9d})')
