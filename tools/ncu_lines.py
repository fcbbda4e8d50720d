#!/usr/bin/env python
"""Source-line view of an ncu report: instructions and stall samples per CUDA source line (top N by samples and by instructions).
usage: python tools/ncu_lines.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep, N = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
al = []
for hi in hdr_idx:
    hdr = rows[hi]
    ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    fname = rows[hi - 2][1].split("/")[-1]
    for r in rows[hi + 1:]:
        if len(r) < len(hdr): break
        try:
            st = {h: int(r[j] or 0) for j, h in stall_cols}
            al.append((fname, int(r[0]), r[1], int(r[ci] or 0), int(r[si] or 0), int(r[ti] or 0), st))
        except ValueError:
            pass
ti = sum(d[3] for d in al); ts = sum(d[4] for d in al)
print(f"total warp instructions {ti}, samples {ts}")
print("== top by samples")
for d in sorted(al, key=lambda d: -d[4])[:N]:
    top = sorted(d[6].items(), key=lambda kv: -kv[1])[:2]
    print(f"{d[0][:13]:13s}{d[1]:5d} samp {d[4]/max(ts,1)*100:5.1f}% inst {d[3]/max(ti,1)*100:5.1f}% lanes {d[5]/max(d[3],1):4.1f} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]} | {d[2][:90]}")
print("== top by instructions")
for d in sorted(al, key=lambda d: -d[3])[:N]:
    print(f"{d[0][:13]:13s}{d[1]:5d} inst {d[3]/max(ti,1)*100:5.1f}% samp {d[4]/max(ts,1)*100:5.1f}% lanes {d[5]/max(d[3],1):4.1f} | {d[2][:100]}")
