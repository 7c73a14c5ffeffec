#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
share of stall samples, instructions per 32-coordinate batch, dominant stall reasons."""
import csv
import sys

path, batches = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 221184.0
rows = list(csv.reader(open(path)))
hdr = rows[2]
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
names = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_not_selected', 'stall_barrier',
         'stall_dispatch', 'stall_selected', 'stall_lg', 'stall_branch_resolving', 'stall_no_inst', 'stall_mio']
cols = {k: hdr.index(k) for k in names}


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = {k: 0 for k in cols}
src = []
for r in rows[3:]:
    if len(r) <= iex or r[0] == '':
        continue
    d = {k: num(r[c]) for k, c in cols.items()}
    for k in d:
        tot[k] += d[k]
    src.append((num(r[0]), r[1][:100], num(r[iex]), num(r[ismp]), d))
S = sum(x[3] for x in src)
I = sum(x[2] for x in src)
print('stall samples by reason:', {k: round(100 * v / max(S, 1), 1) for k, v in tot.items()})
print('samples', S, 'warp instructions', I, 'per batch %.1f' % (I / batches))
for ln, text, n, smp, d in sorted(src, key=lambda x: -x[3])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print('%5.1f%% smp %6.1f inst/batch  L%-4d %-72s' % (100 * smp / S, n / batches, ln, text[:72]),
          {k[6:]: v for k, v in d.items() if v > 0.2 * smp})
