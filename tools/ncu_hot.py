#!/usr/bin/env python
"""Group the SASS page of an .ncu-rep into runs of equal execution count: python tools/ncu_hot.py x.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, iex, isamp, ithr = (hdr.index(x) for x in ('Address', 'Source', 'Instructions Executed', '# Samples', 'Avg. Threads Executed'))
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iex]), int(r[isamp]), r[ia], r[isrc], r[ithr]))
    except Exception:
        pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print('total warp-inst', tot, 'samples', tots, 'sass lines', len(data))
runs = []
for d in data:
    if runs and abs(runs[-1][0] - d[0]) <= max(1, 0.02 * d[0]):
        runs[-1][1] += 1; runs[-1][2] += d[1]; runs[-1][3] += d[0]
    else:
        runs.append([d[0], 1, d[1], d[0], d[2][-5:], d[3][:50], d[4]])
for r in runs:
    if r[3] > tot * 0.01 or r[2] > tots * 0.015:
        print(f"exec={r[0]:>9} n={r[1]:>4} samples={r[2]:>6} ({100*r[2]/tots:4.1f}%) inst={r[3]:>10} ({100*r[3]/tot:4.1f}%) thr={r[6]:>5} @{r[4]} {r[5]}")
