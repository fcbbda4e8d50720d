#!/usr/bin/env python
"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py to the per-kernel shares of the last
product:  python tools/launch_shares.py gpurun_out/launches.csv profiles/rN_launch_shares_<workload>.txt"""
import collections, csv, sys

src, dst = sys.argv[1:3]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
data = [(int(r[iid]), r[ik], float(r[iv].replace(',', ''))) for r in rows[1:]]
idx = [i for i, d in enumerate(data) if 'k_prepare' in d[1]]
call = [d for d in data[idx[-2]:] if 'planes_to_aos' not in d[1]]
tot = sum(d[2] for d in call)
agg = collections.OrderedDict()
for _, k, v in call:
    name = k.split('(')[0].replace('void ', '')
    if name.startswith('cub::'):
        name = name.split('<')[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
with open(dst, 'w') as fh:
    fh.write(f"# the last product of {src.split('/')[-1]} (pk_mul of bench.py's e2e loop; the upload kernels before it are excluded)\n")
    fh.write("# ncu gpu__time_duration.sum per launch (cold-cache, serialised): compare shares, not absolutes\n")
    for name, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        fh.write(f"{v / 1e3:9.1f} us  {100 * v / tot:5.1f} %  x{n}  {name}\n")
    fh.write(f"{tot / 1e3:9.1f} us  total\n")
print(open(dst).read())
