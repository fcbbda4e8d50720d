"""Shared helpers of the test-suite (test scaffolding, not product code)."""
from __future__ import annotations

import numpy as np

from oracle import cpu_oracle, pyoracle as po


def terms(p, nv, seed=0):
    return po.poly_to_soa(p, nv, shuffle_seed=seed)


def as_pairs(key, mag, sign):
    """[(code, coefficient)] in the order given."""
    return list(zip(np.asarray(key).tolist(), po.soa_coefficients(np.asarray(mag).reshape(len(key), -1), sign)))


def assert_terms_equal(got, ref, what=""):
    """Term-by-term equality after canonical sort: same key set, same integer coefficients."""
    gk, gm, gs = got
    rk, rm, rs = ref
    assert len(gk) == len(rk), f"{what}: {len(gk)} terms, expected {len(rk)}"
    go, ro = np.argsort(gk, kind="stable"), np.argsort(rk, kind="stable")
    assert np.array_equal(np.asarray(gk)[go], np.asarray(rk)[ro]), f"{what}: key sets differ"
    if len(gk) == 0:
        return
    gm = np.asarray(gm).reshape(len(gk), -1)[go]
    rm = np.asarray(rm).reshape(len(rk), -1)[ro]
    L = max(gm.shape[1], rm.shape[1])
    gm = np.pad(gm, ((0, 0), (0, L - gm.shape[1])))
    rm = np.pad(rm, ((0, 0), (0, L - rm.shape[1])))
    assert np.array_equal(gm, rm), f"{what}: coefficient magnitudes differ"
    assert np.array_equal(np.asarray(gs)[go], np.asarray(rs)[ro]), f"{what}: signs differ"


def assert_sorted_unique(key):
    key = np.asarray(key)
    assert np.all(key[1:] > key[:-1]), "result keys must be strictly ascending (canonical order, unique)"


def oracle_terms(p, nv):
    """Exact Python-oracle polynomial -> canonical SoA."""
    items = sorted((po.kron_encode(e), c) for e, c in p.items())
    L = max(1, max((abs(c).bit_length() + 63) // 64 for _, c in items)) if items else 1
    key = np.array([k for k, _ in items], dtype=np.int64)
    mag = np.zeros((len(items), L), dtype=np.uint64)
    sign = np.zeros(len(items), dtype=np.int8)
    for i, (_, c) in enumerate(items):
        a = abs(c)
        for l in range(L):
            mag[i, l] = (a >> (64 * l)) & ((1 << 64) - 1)
        sign[i] = 1 if c > 0 else -1
    return key, mag, sign


def cpu_ref(a, b, nv, **kw):
    return cpu_oracle.cpu_mul(a, b, nv, **kw).canonical()


P31 = 2147483629  # prime < 2^31


def eval_mod_p(key, mag, sign, nv, point, p=P31):
    """sum_i cf_i * prod_v point_v^e_iv  (mod p), vectorised; exponents may be negative (modular inverse)."""
    key = np.asarray(key, dtype=np.int64)
    n = len(key)
    if n == 0:
        return 0
    M, hmin, _, _ = po.kron_limits(nv) if nv else ((), 0, 0, 0)
    code = (key - hmin).astype(np.uint64)
    acc = np.ones(n, dtype=np.uint64)
    for v in range(nv):
        r = np.uint64(2 * M[v] + 1)
        e = (code % r).astype(np.int64) - M[v]
        code = code // r
        lo, hi = int(e.min()), int(e.max())
        table = np.array([pow(point[v], ee, p) for ee in range(lo, hi + 1)], dtype=np.uint64)
        acc = acc * table[e - lo] % np.uint64(p)
    mag = np.asarray(mag, dtype=np.uint64).reshape(n, -1)
    cf = np.zeros(n, dtype=np.uint64)
    w = 1
    for l in range(mag.shape[1]):
        cf = (cf + (mag[:, l] % np.uint64(p)) * np.uint64(w)) % np.uint64(p)
        w = w * pow(2, 64, p) % p
    sg = np.where(np.asarray(sign) < 0, np.uint64(p - 1), np.uint64(1))
    tot = acc * cf % np.uint64(p) * sg % np.uint64(p)
    return int(tot.sum(dtype=np.uint64) % np.uint64(p)) if n < (1 << 32) else int(sum(int(x) for x in tot) % p)
