#!/usr/bin/env python
"""pk_mul through a multi-device context: wall time per call for N devices (debug aid; PK_MULTI_TRACE=1 prints the phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import piranha_b200 as pb
from piranha_b200 import abi
from bench import build_operands

w = sys.argv[1] if len(sys.argv) > 1 else "pearce1"
a, b, nv, L = build_operands(w)
for n in [int(x) for x in (sys.argv[2:] or ["1", "2", "4", "8"])]:
    m = pb.Context.multi(list(range(n)))
    pa, ka = m._poly_struct(a, nv); pb_, kb = m._poly_struct(b, nv)
    o, ko = abi.make_opts(); r = abi.PkResult()
    for _ in range(3):
        m.mul_raw(pa, pb_, o, r); m.result_free(r)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); m.mul_raw(pa, pb_, o, r); ts.append(time.perf_counter() - t0); m.result_free(r)
    print(f"{w} devices={n}: pk_mul min {min(ts)*1e3:.3f} ms  med {sorted(ts)[len(ts)//2]*1e3:.3f} ms  terms {int(r.n)}", flush=True)
    m.close()
