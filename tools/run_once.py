#!/usr/bin/env python
"""Run the device-resident product of one or more workloads a few times and print the stats of the last call.
Short command for ncu:  python tools/run_once.py fateman1 [--iters 3] [--algo block] [--part i/n] [key=value tuning ...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piranha_b200 as pb
from bench import build_operands

def main():
    args = sys.argv[1:]
    iters, algo, tuning, wl, trunc, part = 3, pb.PK_ALGO_AUTO, {}, [], None, None
    i = 0
    while i < len(args):
        a = args[i]
        if a == "--iters": iters = int(args[i + 1]); i += 2; continue
        if a == "--trunc": trunc = int(args[i + 1]); i += 2; continue
        if a == "--part": part = tuple(int(x) for x in args[i + 1].split("/")); i += 2; continue
        if a == "--algo": algo = {v: k for k, v in pb.ALGO_NAMES.items()}[args[i + 1]]; i += 2; continue
        if "=" in a: k, v = a.split("="); tuning[k] = int(v)
        else: wl.append(a)
        i += 1
    for w in wl or ["fateman1"]:
        a, b, nv, L = build_operands(w)
        ctx = pb.Context(0)
        if tuning: ctx.set_tuning(**tuning)
        da, db = ctx.upload(a, nv), ctx.upload(b, nv)
        ks, ds = [], []
        for it in range(iters):
            opt = dict(trunc_mode=1, max_degree=trunc) if trunc is not None else {}
            if part: opt.update(part_index=part[0], part_count=part[1])
            r = ctx.mul_dev(da, db, algo=algo, **opt)
            st = ctx.stats()
            ks.append(st["mul_kernel_ms"]); ds.append(st["device_ms"])
        W = st["term_products"]
        print(f"{w}: algo={pb.ALGO_NAMES[st['algo']]} mode={'bitmap' if st['retries'] & 1 else 'dense'} tpb={st['retries'] >> 8} "
              f"NW={st['acc_lanes']} cap={st['table_slots']} R={st['result_terms']} launches={st['kernel_launches']} "
              f"kernel_ms min {min(ks):.4f} med {statistics.median(ks):.4f}  device_ms min {min(ds):.4f} med {statistics.median(ds):.4f}  "
              f"-> {W / min(ks) / 1e6:.1f} G products/s (kernel)  digest {r.digest():016x}", flush=True)
        if os.environ.get("PK_ONCE_ALL"):
            print("   device_ms:", " ".join(f"{x:.3f}" for x in ds), flush=True)
            print("   kernel_ms:", " ".join(f"{x:.3f}" for x in ks), flush=True)
        ctx.close()

if __name__ == "__main__":
    main()
