#!/usr/bin/env python
"""Summarise a PK_ZONE_TRACE dump: python tools/trace_report.py trace.bin"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
t = t[t[:, 1] > 0]  # chunks that ran
start, end = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)
claimed, smid = (t[:, 2] >> np.uint64(32)).astype(np.int64), (t[:, 2] & np.uint64(0xFFFFFFFF)).astype(np.int64)
w = t[:, 3].astype(np.int64)
t0 = start.min()
dur = (end - start) / 1e3
print(f"chunks {len(t)}  kernel span {(end.max()-t0)/1e3:.1f} us  sum of chunk time {dur.sum()/1e3:.2f} ms  "
      f"mean {dur.mean():.1f} us  max {dur.max():.1f} us  p50 {np.median(dur):.1f} us")
per_sm = {}
for s_, d, e in zip(smid, dur, end):
    a = per_sm.setdefault(int(s_), [0.0, 0, 0])
    a[0] += d; a[1] += 1; a[2] = max(a[2], e - t0)
busy = np.array([v[0] for v in per_sm.values()]); fin = np.array([v[2] for v in per_sm.values()]) / 1e3
print(f"SMs {len(per_sm)}  busy/SM mean {busy.mean():.1f} max {busy.max():.1f} min {busy.min():.1f} us;  finish time mean {fin.mean():.1f} max {fin.max():.1f} min {fin.min():.1f} us")
order = np.argsort(-dur)[:8]
print("slowest chunks (id, claimed#, dur us, start us, rel work):")
for i in order:
    print(f"   {i:5d} {claimed[i]:5d} {dur[i]:8.1f} {(start[i]-t0)/1e3:8.1f} {w[i]}")
# correlation between thread-32 product count and duration
if w.sum() > 0:
    print("corr(duration, products) =", float(np.corrcoef(dur, w)[0, 1]))
    # linear fit dur = a + b*w
    A = np.vstack([np.ones_like(w, dtype=float), w.astype(float)]).T
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    print(f"fit: dur_us = {coef[0]:.1f} + {coef[1]:.4f} * products_of_one_thread")
last = np.argsort(-end)[:5]
print("last finishing chunks:", [(int(i), int(claimed[i]), round(float(dur[i]),1)) for i in last])
