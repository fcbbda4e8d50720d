#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels in the built library (cuobjdump -sass): the evidence for which instructions
the pairing kernels issue (ATOMS = shared-memory atomics, UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, ...).
usage: python tools/sass_hist.py [kernel-name-substring ...] > profiles/r2_sass_hist.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "piranha_b200", "lib", "libpiranha_b200.so")
want = sys.argv[1:] or ["k_block_accILi3ELi3ELb0ELi1E", "k_block_denseILi3ELi3ELb0ELi1E", "k_block_denseILi5ELi4ELb0ELi1E", "k_block_mark",
                        "k_block_plan", "k_block_zones", "k_block_tilesILi1E", "k_mul_globalILi1ELi1ELi1ELb0E"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
for w in want:
    for name, ops in funcs.items():
        if w not in name:
            continue
        base = collections.Counter(o.split(".")[0] for o in ops)
        full = collections.Counter(ops)
        print(f"== {demangle(name)}\n   {len(ops)} SASS instructions")
        print("   by mnemonic: " + ", ".join(f"{k} {v}" for k, v in base.most_common(24)))
        keys = [k for k in full if k.split(".")[0] in ("ATOMS", "ATOMG", "RED", "REDG", "UBLKCP", "SYNCS", "LDG", "LDS", "STS", "STG", "BAR", "IMAD", "SHFL", "POPC")]
        print("   bulk copy / mbarrier: " + (", ".join(f"{k} {full[k]}" for k in sorted(full) if k.split(".")[0] in ("UBLKCP", "SYNCS", "UTMALDG")) or "none"))
        print("   memory / sync / multiply variants: " + ", ".join(f"{k} {full[k]}" for k in sorted(keys, key=lambda k: -full[k])[:28]))
        break
