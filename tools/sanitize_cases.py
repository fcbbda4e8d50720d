#!/usr/bin/env python
"""Small products through every accumulation strategy, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piranha_b200 as pb
from oracle import pyoracle as po

def main():
    ctx = pb.Context(0)
    cases = [("fateman5", po.fateman_operands(5), 4), ("pearce3", po.pearce_operands(3), 5), ("monagan5", po.monagan_operands(5), 5)]
    for name, (f, g), nv in cases:
        if name == "monagan5":
            f = {k: v for i, (k, v) in enumerate(f.items()) if i % 7 == 0}
            g = {k: -v if i % 2 else v for i, (k, v) in enumerate(g.items()) if i % 9 == 0}
        a, b = po.poly_to_soa(f, nv, shuffle_seed=1), po.poly_to_soa(g, nv, shuffle_seed=2)
        da, db = ctx.upload(a, nv), ctx.upload(b, nv)
        ref = None
        runs = [("dense", {}), ("hash", {}), ("zone", dict(zone_mode=1)), ("zone", dict(zone_mode=2, zone_cap=37)),
                ("block", dict(block_mode=1, block_min_tile=0)), ("block", dict(block_mode=2, block_min_tile=0)),
                ("block", dict(block_mode=2, block_hz=3, zone_cap=29, block_min_tile=0))]
        for algo, tuning in runs:
            c2 = pb.Context(0)
            c2.set_tuning(**tuning) if tuning else None
            xa, xb = c2.upload(a, nv), c2.upload(b, nv)
            r = c2.mul_dev(xa, xb, algo={v: k for k, v in pb.ALGO_NAMES.items()}[algo])
            d = r.digest()
            ref = ref if ref is not None else d
            assert d == ref, (name, algo, tuning)
            rt = c2.mul_dev(xa, xb, algo={v: k for k, v in pb.ALGO_NAMES.items()}[algo], trunc_mode=1, max_degree=6)
            parts = [c2.mul_dev(xa, xb, part_index=p, part_count=3).digest() for p in range(3)]
            assert sum(parts) & ((1 << 64) - 1) == ref, (name, "parts")
            print(name, algo, tuning, "ok", len(r), len(rt), flush=True)
            del r, rt, xa, xb
            c2.close()
    ctx.close()
    print("all cases ok")

if __name__ == "__main__":
    main()
