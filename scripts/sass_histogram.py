#!/usr/bin/env python
"""Opcode histogram of one kernel launch from an .ncu-rep (ncu --set full --import-source on): executed warp instructions
per 32 coordinates (= thread instructions per coordinate) and static count per SASS opcode.
  python scripts/sass_histogram.py gpurun_out/r2_quantize.ncu-rep [coordinates per launch]"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
coords = float(sys.argv[2]) if len(sys.argv) > 2 else 7077888.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Instructions Executed" in r)
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
ex, st = defaultdict(float), defaultdict(int)
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= i_ex:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[i_src])
    if not m:
        continue
    op = m.group(1)
    try:
        n = float(r[i_ex])
    except ValueError:
        n = 0.0
    ex[op] += n
    st[op] += 1
tot = sum(ex.values())
print("Total executed: %.1f thread instructions per coordinate (%d warp instructions); %d static instructions" % (tot * 32 / coords, tot, sum(st.values())))
print("%-12s %14s %8s" % ("opcode", "exec/coord", "static"))
for op in sorted(ex, key=lambda k: -ex[k]):
    print("%-12s %14.2f %8d" % (op, ex[op] * 32 / coords, st[op]))
