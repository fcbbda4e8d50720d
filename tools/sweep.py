#!/usr/bin/env python
"""Tuning sweep of the zone kernel on one GPU: python tools/sweep.py [workload ...]
Prints median pairing-kernel ms and whole-call device ms per tuning combination (results are always verified by digest
against the first run)."""
import itertools, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piranha_b200 as pb
from bench import build_operands

def main():
    wl = sys.argv[1:] or ["fateman1", "pearce1"]
    grid = {"zone_tpb": (256, 512, 1024), "zone_chunk": (2, 4, 8), "zone_target": (1184, 2368, 4736)}
    if os.environ.get("SWEEP_GRID"):  # e.g. "zone_tpb=1024;zone_chunk=1,2,3;zone_target=2368,3552"
        grid = {k: tuple(int(x) for x in v.split(",")) for k, v in (kv.split("=") for kv in os.environ["SWEEP_GRID"].split(";"))}
    combos = [dict(zip(grid, vals)) for vals in itertools.product(*grid.values())]
    extra = os.environ.get("SWEEP_EXTRA")
    for w in wl:
        a, b, nv, L = build_operands(w)
        ctx = pb.Context(0)
        da, db = ctx.upload(a, nv), ctx.upload(b, nv)
        ref = None
        rows = []
        for c in combos:
            ctx.set_tuning(**c)
            if extra:
                ctx.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in extra.split(","))})
            ks, ds = [], []
            for it in range(6):
                r = ctx.mul_dev(da, db)
                st = ctx.stats()
                if it:
                    ks.append(st["mul_kernel_ms"]); ds.append(st["device_ms"])
            d = r.digest()
            if ref is None: ref = d
            assert d == ref, "digest changed with tuning!"
            rows.append((statistics.median(ks), statistics.median(ds), c, st["retries"] & 1, st["table_slots"]))
        rows.sort(key=lambda x: x[0])
        print(f"== {w}: best 8 of {len(rows)}")
        for k, d, c, mode, cap in rows[:8]:
            print(f"  kernel {k:.3f} ms  device {d:.3f} ms  {c}  mode={'bitmap' if mode else 'dense'} cap={cap}")
        print("  worst:", f"{rows[-1][0]:.3f}", rows[-1][2])
        ctx.close()

if __name__ == "__main__":
    main()
