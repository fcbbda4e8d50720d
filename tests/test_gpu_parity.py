"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracles, bit-exact.

The cases follow the reference's own tests for this path:
  tests/polynomial_multiplier_01.cpp:120-174 (bounds), :187-238 (empty, reduced Fateman 10626 / 5786)
  tests/base_series_multiplier.cpp:517-605     (reduced Fateman and Pearce: 591235 / 591184)
  tests/polynomial_truncation.cpp:69-177       (total / partial degree truncation)
  tests/kronecker_monomial_01.cpp:404-461      (term multiply: 2*3 = 6, 2*(-4) = -8 with exponent sums)
  benchmarks/*.cpp                             (result sizes of fateman1/2, pearce1, monagan1-5, truncated fateman1)
"""
import os

import numpy as np
import pytest

import piranha_b200 as pb
from oracle import pyoracle as po
from polyhelper import (P31, assert_sorted_unique, assert_terms_equal, cpu_ref, eval_mod_p, oracle_terms, terms)

pytestmark = pytest.mark.gpu

ALGOS = [pb.PK_ALGO_AUTO, pb.PK_ALGO_DENSE, pb.PK_ALGO_HASH, pb.PK_ALGO_ZONE, pb.PK_ALGO_BLOCK]


def gpu_mul(ctx, f, g, nv, **opt):
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    out = ctx.mul(a, b, nv, **opt)
    assert_sorted_unique(out[0])
    return a, b, out


# ---------------------------------------------------------------------------------------------- tiny exact cases


def test_term_multiply_fixtures(ctx):
    # kronecker_monomial_01.cpp:404-461: 2*x * 3*x^2 = 6*x^3 ; 2*x*y^-1 * -4*x^2 = -8*x^3*y^-1
    r = ctx.mul(terms({(1,): 2}, 1), terms({(2,): 3}, 1), 1)
    assert_terms_equal(r, oracle_terms({(3,): 6}, 1))
    r = ctx.mul(terms({(1, -1): 2}, 2), terms({(2, 0): -4}, 2), 2)
    assert_terms_equal(r, oracle_terms({(3, -1): -8}, 2))


def test_empty_and_constants(ctx):
    # polynomial_multiplier_01.cpp:193-202: empty * x = 0 with the symbol set preserved; pt{2} * pt{3} = 6
    empty = (np.zeros(0, np.int64), np.zeros((0, 1), np.uint64), np.zeros(0, np.int8))
    x = terms({(1,): 1}, 1)
    for a, b in ((empty, x), (x, empty), (empty, empty)):
        k, m, s = ctx.mul(a, b, 1)
        assert len(k) == 0
    r = ctx.mul(terms({(): 2}, 0), terms({(): 3}, 0), 0)
    assert_terms_equal(r, oracle_terms({(): 6}, 0))
    r = ctx.mul(terms({(0, 0, 0): 2}, 3), terms({(0, 0, 0): 3}, 3), 3)
    assert_terms_equal(r, oracle_terms({(0, 0, 0): 6}, 3))


def test_square_of_linear(ctx):
    # polynomial_truncation.cpp:75-77
    f = {(1, 0, 0, 0): 1, (0, 1, 0, 0): 1, (0, 0, 1, 0): 1, (0, 0, 0, 1): 1}
    for algo in ALGOS:
        _, _, r = gpu_mul(ctx, f, f, 4, algo=algo)
        assert_terms_equal(r, oracle_terms(po.p_mul(f, f), 4))


@pytest.mark.parametrize("algo", ALGOS)
def test_truncation_exact_cases(ctx, algo):
    # polynomial_truncation.cpp:79-107 (total degree), :141-170 (partial degree); variables t,x,y,z -> indices 0..3
    nv = 4
    one = (0, 0, 0, 0)
    X, Y = (0, 1, 0, 0), (0, 0, 1, 0)
    f = {one: 1, X: 1, Y: 1}
    for D in (1, 2, 3, 0, -1):
        _, _, r = gpu_mul(ctx, f, f, nv, trunc_mode=1, max_degree=D, algo=algo)
        assert_terms_equal(r, oracle_terms(po.p_mul(f, f, max_degree=D), nv), f"total D={D}")
    lin = {(1, 0, 0, 0): 1, X: 1, Y: 1, (0, 0, 0, 1): 1}
    _, _, r = gpu_mul(ctx, lin, lin, nv, trunc_mode=1, max_degree=1, algo=algo)
    assert len(r[0]) == 0
    for mask in ([0, 1, 0, 0], [0, 0, 0, 1], [0, 1, 1, 0]):
        _, _, r = gpu_mul(ctx, f, f, nv, trunc_mode=2, max_degree=1, degree_mask=mask, algo=algo)
        assert_terms_equal(r, oracle_terms(po.p_mul(f, f, max_degree=1, mask=mask), nv), f"partial {mask}")
    _, _, r = gpu_mul(ctx, lin, lin, nv, trunc_mode=2, max_degree=1, degree_mask=[0, 1, 0, 0], algo=algo)
    assert_terms_equal(r, oracle_terms(po.p_mul(lin, lin, max_degree=1, mask=[0, 1, 0, 0]), nv))


def test_truncation_degree_overflow(ctx):
    # degree_sub -> safe_int_sub overflow -> std::overflow_error (polynomial.hpp:1243-1252)
    f = {(-1,): 1, (0,): 1}
    with pytest.raises(OverflowError):
        ctx.mul(terms(f, 1), terms({(1,): 1}, 1), 1, trunc_mode=1, max_degree=(1 << 63) - 1)


def test_bounds_fixtures(ctx):
    # polynomial_multiplier_01.cpp:120-174, 3 variables
    M = po.kron_limits(3)[0]
    lim, _, _ = ctx.limits(3)
    assert tuple(int(x) for x in lim) == M
    top = {(M[0], M[1], M[2]): 1}
    for v in range(3):
        e = [0, 0, 0]
        e[v] = 1
        with pytest.raises(OverflowError):
            ctx.mul(terms(top, 3), terms({tuple(e): 1}, 3), 3)
        with pytest.raises(OverflowError):  # (x^M0 + 1)(y^M1 + 1)(z^M2 + 1) * (v + 1)
            big = po.p_mul(po.p_mul({(M[0], 0, 0): 1, (0, 0, 0): 1}, {(0, M[1], 0): 1, (0, 0, 0): 1}),
                           {(0, 0, M[2]): 1, (0, 0, 0): 1})
            ctx.mul(terms(big, 3), terms({tuple(e): 1, (0, 0, 0): 1}, 3), 3)
        neg = [M[0], M[1], M[2]]
        neg[v] = -neg[v]
        e[v] = -1
        with pytest.raises(OverflowError):
            ctx.mul(terms({tuple(neg): 1}, 3), terms({tuple(e): 1}, 3), 3)
        # one below the limit succeeds and lands exactly on the limit
        below = [M[0], M[1], M[2]]
        below[v] -= 1
        e[v] = 1
        r = ctx.mul(terms({tuple(below): 1}, 3), terms({tuple(e): 1}, 3), 3)
        assert_terms_equal(r, oracle_terms(top, 3))
    allneg = {(-M[0] + 1, -M[1], -M[2]): 1}
    r = ctx.mul(terms(allneg, 3), terms({(-1, 0, 0): 1}, 3), 3)
    assert_terms_equal(r, oracle_terms({(-M[0], -M[1], -M[2]): 1}, 3))


def test_argument_errors(ctx):
    a = terms({(1,): 1}, 1)
    da, db = ctx.upload(a, 1), ctx.upload(terms({(1, 0): 1}, 2), 2)
    with pytest.raises(ValueError):  # "incompatible arguments sets", base_series_multiplier.hpp:365-367
        ctx.mul_dev(da, db)
    with pytest.raises(ValueError):
        ctx.mul(a, a, 1, trunc_mode=7)
    with pytest.raises(ValueError):
        ctx.mul(a, a, 1, part_index=3, part_count=2)


# ---------------------------------------------------------------------------------------------- known answers


@pytest.mark.parametrize("algo", ALGOS)
def test_reduced_fateman(ctx, algo):
    f, g = po.fateman_operands(10)
    a, b, r = gpu_mul(ctx, f, g, 4, algo=algo)
    assert len(r[0]) == 10626  # polynomial_multiplier_01.cpp:214
    assert_terms_equal(r, cpu_ref(a, b, 4))
    assert_terms_equal(r, oracle_terms(po.fateman_product_closed_form(10), 4), "closed form")


@pytest.mark.parametrize("algo", ALGOS)
def test_reduced_fateman_cancellations(ctx, algo):
    f, h = po.fateman_cancel_operands(10)
    a, b, r = gpu_mul(ctx, f, h, 4, algo=algo)
    assert len(r[0]) == 5786  # polynomial_multiplier_01.cpp:230
    assert_terms_equal(r, cpu_ref(a, b, 4))


@pytest.mark.parametrize("minus_u,size", [(False, 591235), (True, 591184)])
@pytest.mark.parametrize("algo", ALGOS)
def test_reduced_pearce(ctx, algo, minus_u, size):
    f, g = po.pearce_operands(8, minus_u)
    a, b, r = gpu_mul(ctx, f, g, 5, algo=algo)
    assert len(r[0]) == size  # base_series_multiplier.cpp:586, :604
    assert_terms_equal(r, cpu_ref(a, b, 5, n_threads=4))


def test_fateman1_full(ctx):
    f, g = po.fateman_operands(20)
    a, b, r = gpu_mul(ctx, f, g, 4)
    assert len(r[0]) == 135751  # benchmarks/fateman1.cpp:55
    assert_terms_equal(r, oracle_terms(po.fateman_product_closed_form(20), 4), "closed form")
    st = ctx.stats()
    assert st["term_products"] == 10626 * 10626 and st["kernel_launches"] > 0


def test_pearce1_full(ctx):
    f, g = po.pearce_operands(12)
    a, b, r = gpu_mul(ctx, f, g, 5)
    assert len(r[0]) == 5821335  # benchmarks/pearce1.cpp:56
    assert_terms_equal(r, cpu_ref(a, b, 5, n_threads=8))
    # size-independent property: evaluation homomorphism at a random point mod a prime
    pt = [3, 5, 7, 11, 13]
    ea, eb = eval_mod_p(*a, 5, pt), eval_mod_p(*b, 5, pt)
    assert eval_mod_p(*r, 5, pt) == ea * eb % P31


@pytest.mark.parametrize("parts", [2, 8])
def test_pearce1_full_partitions(ctx, parts):
    """BASELINE-size pearce1 cut into output partitions (the per-rank work of a multi-GPU product; bitmap zones cut at
    equal work): ordered, disjoint, and together the whole product -- checksum of checksums and term counts."""
    f, g = po.pearce_operands(12)
    a, b = terms(f, 5, 1), terms(g, 5, 2)
    da, db = ctx.upload(a, 5), ctx.upload(b, 5)
    whole = ctx.mul_dev(da, db)
    digest, n, pairs, last_key = 0, 0, 0, None
    for p in range(parts):
        part = ctx.mul_dev(da, db, part_index=p, part_count=parts)
        st = ctx.stats()
        pairs += st["term_products"]
        digest = (digest + part.digest()) & ((1 << 64) - 1)
        n += len(part)
        assert len(part) > 0
        k = part.download()[0]
        assert_sorted_unique(k)
        assert last_key is None or k[0] > last_key
        last_key = k[-1]
    assert n == len(whole) == 5821335
    assert digest == whole.digest()
    assert pairs == 6188 * 6188


@pytest.mark.parametrize("which,size", [(1, 12341), (2, 12341), (3, 39711), (4, 135751), (5, 417311)])
def test_monagan(ctx, which, size):
    f, g = po.monagan_operands(which)
    nv = len(next(iter(f)))
    a, b, r = gpu_mul(ctx, f, g, nv)
    assert len(r[0]) == size  # benchmarks/monagan1-5.cpp
    assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=8))


def test_monagan5_all_strategies(ctx):
    f, g = po.monagan_operands(5)
    a, b = terms(f, 5, 1), terms(g, 5, 2)
    ref = cpu_ref(a, b, 5, n_threads=8)
    for algo in (pb.PK_ALGO_DENSE, pb.PK_ALGO_HASH, pb.PK_ALGO_ZONE, pb.PK_ALGO_BLOCK):
        assert_terms_equal(ctx.mul(a, b, 5, algo=algo), ref)


@pytest.mark.parametrize("tuning", [
    dict(zone_mode=1),                               # dense zones
    dict(zone_mode=2),                               # bitmap zones (perfect hash by rank)
    dict(zone_mode=2, zone_cap=37),                  # bitmap zones far over capacity: many passes per zone
    dict(zone_mode=2, zone_span_log2=10, zone_cap=5),
    dict(zone_mode=1, zone_cap=64),                  # tiny dense zones: long look-back chains
    dict(zone_mode=2, zone_cursor_bytes=0),          # no row cursors: one binary search per row per zone
    dict(zone_mode=1, zone_cursor_bytes=0, zone_target=7),
    dict(zone_mode=2, zone_chunk=1, zone_target=100000),
    dict(zone_mode=1, zone_chunk=64),
    dict(zone_mode=1, zone_tpb=256),
    dict(zone_mode=2, zone_uniform=1),
    dict(zone_mode=1, zone_uniform=1, zone_chunk=3),
    dict(zone_mode=2, zone_tpb=256, zone_cap=100),
    dict(zone_mode=2, zone_tpb=512),
    dict(zone_mode=1, zone_tpb=1024, zone_target=3),
    dict(zone_mode=2, zone_tpb=1024, zone_span_log2=20),
])
@pytest.mark.parametrize("case", ["pearce", "cancel", "signed"])
def test_zone_variants(tuning, case):
    """Every way the zone kernel can be configured gives the same bits (tuning never changes results)."""
    if case == "pearce":
        (f, g), nv = po.pearce_operands(5), 5
    elif case == "cancel":
        (f, g), nv = po.fateman_cancel_operands(7), 4
    else:
        (f, g), nv = po.monagan_operands(5), 5
        f = {k: v for i, (k, v) in enumerate(f.items()) if i % 3 == 0}
        g = {k: -v if i % 2 else v for i, (k, v) in enumerate(g.items()) if i % 5 == 0}
    c2 = pb.Context(0)
    try:
        c2.set_tuning(**tuning)
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        r = c2.mul(a, b, nv, algo=pb.PK_ALGO_ZONE)
        st = c2.stats()
        assert st["algo"] == pb.PK_ALGO_ZONE
        assert_sorted_unique(r[0])
        assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=4))
        assert st["term_products"] == len(a[0]) * len(b[0])
    finally:
        c2.close()


def _variant_case(case):
    if case == "pearce":
        (f, g), nv = po.pearce_operands(5), 5
    elif case == "cancel":
        (f, g), nv = po.fateman_cancel_operands(7), 4
    else:
        (f, g), nv = po.monagan_operands(5), 5
        f = {k: v for i, (k, v) in enumerate(f.items()) if i % 3 == 0}
        g = {k: -v if i % 2 else v for i, (k, v) in enumerate(g.items()) if i % 5 == 0}
    return f, g, nv


@pytest.mark.parametrize("tuning", [
    dict(block_mode=1),                                   # dense blocks
    dict(block_mode=2),                                   # bitmap blocks
    dict(block_mode=1, block_hz=1),
    dict(block_mode=1, block_hz=7, block_natural=1),
    dict(block_mode=2, block_hz=3, zone_cap=37),          # bitmap zones far over capacity: many passes per zone
    dict(block_mode=2, block_hz=1000000, zone_cap=5),
    dict(block_mode=1, block_target=32),                  # many tiny tiles
    dict(block_mode=2, block_target=100000),              # whole group pairs as single tiles
    dict(block_mode=1, block_hz=2, block_target=64),
    dict(block_mode=2, block_natural=1, block_hz=5),
    dict(block_mode=1, block_zone_work=1),
    dict(block_mode=2, block_zone_work=100000000, zone_span_log2=12),
    dict(block_mode=2, zone_span_log2=20),
    dict(block_mode=1, block_permute=1),
    dict(block_mode=2, block_permute=1, block_hz=2),
    dict(block_mode=2, block_hash=0),                     # two-walk bitmap scheme only
    dict(block_mode=2, block_hash=0, block_hz=3, zone_cap=37),
    dict(block_mode=2, block_hash=1, zone_cap=64),         # hash attempts that run out of probes and fall back
    dict(block_mode=2, block_hash=1, block_hz=1000000),
    dict(block_mode=2, block_ctas=2),                      # two accumulating blocks of 512 threads per SM
    dict(block_mode=2, block_ctas=2, block_hz=3, zone_cap=37),
    dict(block_mode=2, block_merge=0),                     # one tile per pair of groups (no row ranges across groups)
    dict(block_mode=1, block_merge=0, block_hz=2),
    dict(block_mode=2, block_unit_codes=64, block_zone_products=100),   # tiny zone units, tiny zones
    dict(block_mode=2, block_unit_codes=1000000, block_zone_products=1000000000),
])
@pytest.mark.parametrize("case", ["pearce", "cancel", "signed"])
def test_block_variants(tuning, case):
    """Every way the block kernel can be configured gives the same bits."""
    f, g, nv = _variant_case(case)
    c2 = pb.Context(0)
    try:
        c2.set_tuning(block_min_tile=0, **tuning)
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        r = c2.mul(a, b, nv, algo=pb.PK_ALGO_BLOCK)
        st = c2.stats()
        assert st["algo"] == pb.PK_ALGO_BLOCK
        assert_sorted_unique(r[0])
        assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=4))
        assert st["term_products"] == len(a[0]) * len(b[0])
    finally:
        c2.close()


@pytest.mark.parametrize("tuning", [
    dict(block_mode=1),
    dict(block_mode=2),
    dict(block_mode=2, block_hz=4, zone_cap=41),     # multi-pass bitmap zones with the degree test
    dict(block_mode=1, block_target=32, block_wide=0),
])
@pytest.mark.parametrize("case", ["total", "partial", "laurent", "none_pass", "all_pass"])
def test_block_truncation(tuning, case):
    """Degree truncation inside the block kernel (tile-level pruning + per-product test on straddling tiles)."""
    rs = np.random.RandomState(5)
    nv = 4
    if case == "laurent":
        f = {(int(a) - 6, int(b), int(c) - 3, int(d)): int(v) - 50 or 1 for a, b, c, d, v in rs.randint(0, 12, size=(900, 5)) * [1, 1, 1, 1, 9]}
        g = {(int(a), int(b) - 5, int(c), int(d) - 2): int(v) + 1 for a, b, c, d, v in rs.randint(0, 12, size=(700, 5))}
        opt = dict(trunc_mode=1, max_degree=7)
    else:
        f, g = po.fateman_operands(7)
        g = {k: (-v if sum(k) % 3 == 0 else v) for k, v in g.items()}
        opt = {"total": dict(trunc_mode=1, max_degree=9), "partial": dict(trunc_mode=2, max_degree=5, degree_mask=[1, 0, 1, 0]),
               "none_pass": dict(trunc_mode=1, max_degree=-1), "all_pass": dict(trunc_mode=1, max_degree=1000)}[case]
    c2 = pb.Context(0)
    try:
        c2.set_tuning(block_min_tile=0, **tuning)
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        r = c2.mul(a, b, nv, algo=pb.PK_ALGO_BLOCK, **opt)
        st = c2.stats()
        assert st["algo"] == pb.PK_ALGO_BLOCK
        assert_sorted_unique(r[0])
        kw = dict(trunc_mode=opt["trunc_mode"], max_degree=opt["max_degree"])
        if "degree_mask" in opt:
            kw["deg_mask"] = opt["degree_mask"]
        assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=4, **kw))
        # the pairs actually multiplied are exactly those within the degree limit
        mask = opt.get("degree_mask", [1] * nv)
        da = np.array([sum(e for e, m in zip(k, mask) if m) for k in f])
        db = np.array([sum(e for e, m in zip(k, mask) if m) for k in g])
        assert st["term_products"] == int((da[:, None] + db[None, :] <= opt["max_degree"]).sum())
    finally:
        c2.close()


def test_block_truncation_degree_overflow(ctx):
    # the checked subtraction max_degree - d1 (polynomial.hpp:1243-1252) also guards the block path
    f, g = po.fateman_operands(6)
    f = {(k[0] - 3, k[1], k[2], k[3]): v for k, v in f.items()}
    with pytest.raises(OverflowError):
        ctx.mul(terms(f, 4, 1), terms(g, 4, 2), 4, trunc_mode=1, max_degree=2 ** 63 - 1, algo=pb.PK_ALGO_BLOCK)


@pytest.mark.parametrize("small", [0, 1])
@pytest.mark.parametrize("parts", [2, 4, 7])
def test_pipelined_host_product(parts, small):
    """pk_mul of a sparse product computes output ranges one after the other and copies each to the host while the next
    is computed; the result is the same bits as the one-shot product (small=1 undersizes the host arrays on purpose so
    that the exact re-copy path runs)."""
    (f, g), nv = po.pearce_operands(6), 5
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    c2 = pb.Context(0)
    try:
        c2.set_tuning(pipe_parts=0)
        ref = c2.mul(a, b, nv)
        st0 = c2.stats()
        c2.set_tuning(pipe_parts=parts, pipe_min_products=1, pipe_test_small=small)
        r = c2.mul(a, b, nv)
        st = c2.stats()
        assert_sorted_unique(r[0])
        assert_terms_equal(r, ref)
        assert st["term_products"] == st0["term_products"] == len(a[0]) * len(b[0])
        assert st["result_terms"] == len(ref[0])
        assert st["kernel_launches"] > st0["kernel_launches"]  # it really ran in parts
        assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=4))
    finally:
        c2.close()


def test_block_declines_without_a_digit_split(ctx):
    # one variable: there is no hi digit, the generic zone path takes over
    f = {(i,): i + 1 for i in range(0, 3000, 3)}
    g = {(i,): 2 * i + 1 for i in range(0, 2000, 2)}
    a, b = terms(f, 1, 1), terms(g, 1, 2)
    r = ctx.mul(a, b, 1, algo=pb.PK_ALGO_BLOCK)
    assert ctx.stats()["algo"] == pb.PK_ALGO_ZONE
    assert_terms_equal(r, cpu_ref(a, b, 1))


def test_block_is_default_for_benchmark_configs(ctx):
    for (f, g), nv in ((po.fateman_operands(12), 4), (po.pearce_operands(7), 5)):
        ctx.mul(terms(f, nv), terms(g, nv, 1), nv)
        assert ctx.stats()["algo"] == pb.PK_ALGO_BLOCK


def test_fateman2_full(ctx):
    # stresses coefficient width: largest result coefficient is 2^127.96
    f, g = po.fateman_operands(30)
    a, b, r = gpu_mul(ctx, f, g, 4)
    assert len(r[0]) == 635376  # benchmarks/fateman2.cpp:55
    assert_terms_equal(r, oracle_terms(po.fateman_product_closed_form(30), 4), "closed form")
    mx = max(abs(c) for c in po.soa_coefficients(r[1], r[2]))
    assert 127 < mx.bit_length() <= 128


@pytest.mark.parametrize("D,size", [(20, 10626), (30, 46376)])
def test_fateman1_truncated(ctx, D, size):
    f, g = po.fateman_operands(20)
    a, b, r = gpu_mul(ctx, f, g, 4, trunc_mode=1, max_degree=D)
    assert len(r[0]) == size  # benchmarks/fateman1_unpacked_truncation.cpp:57, :59
    assert_terms_equal(r, cpu_ref(a, b, 4, n_threads=8, trunc_mode=1, max_degree=D))
    assert ctx.stats()["term_products"] == {20: 3108105, 30: 38970998}[D]  # SURVEY.md section 8 size table


@pytest.mark.parametrize("which", ["fateman1_dynamic", "pearce_dynamic_reduced"])
def test_dynamic_benchmarks(ctx, which):
    """benchmarks/fateman1_dynamic.cpp, pearce1_dynamic.cpp: the first operand scaled by 2^64 - 1, so operands need two
    limbs and the result four (mp++ would leave its static storage; here the width bound selects wider accumulators)."""
    factor = 2 ** 64 - 1
    if which == "fateman1_dynamic":
        (f, g), nv, size = po.fateman_operands(20), 4, 135751
        f = {k: v * factor for k, v in f.items()}
        g = dict(f)
        one = (0,) * nv
        g[one] = g.get(one, 0) + 1  # f * (f + 1) with the scaled f (benchmarks/fateman1.hpp:49-55)
    else:
        (f, g), nv, size = po.pearce_operands(6), 5, None
        f = {k: v * factor for k, v in f.items()}
        g = {k: v * factor for k, v in g.items()}
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    assert a[1].shape[1] == 2
    r = ctx.mul(a, b, nv)
    assert_sorted_unique(r[0])
    if size is not None:
        assert len(r[0]) == size
    assert r[1].shape[1] >= 3
    assert_terms_equal(r, cpu_ref(a, b, nv, n_threads=8))


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("case", ["both_wide", "a_wide", "b_wide", "wide_trunc", "max_width"])
def test_block_two_limb_operands(ctx, mode, case):
    """The block kernel with operands of up to two limbs (products of up to 256 bits, accumulators of up to 9 words)."""
    (f, g), nv = po.fateman_cancel_operands(7), 4
    big, huge = 2 ** 64 - 1, 2 ** 100 + 12345
    opt, kw = {}, {}
    if case in ("both_wide", "wide_trunc"):
        f = {k: v * big for k, v in f.items()}
        g = {k: v * huge for k, v in g.items()}
    elif case == "a_wide":
        f = {k: v * huge for k, v in f.items()}
    elif case == "b_wide":
        g = {k: v * big * 3 for k, v in g.items()}
    else:  # magnitudes just below 2^128 on both sides
        f = {k: (2 ** 128 - 1 - i) * (1 if v > 0 else -1) for i, (k, v) in enumerate(f.items())}
        g = {k: (2 ** 128 - 159 - 7 * i) * (1 if v > 0 else -1) for i, (k, v) in enumerate(g.items())}
    if case == "wide_trunc":
        opt, kw = dict(trunc_mode=1, max_degree=9), dict(max_degree=9)
    ctx.set_tuning(block_mode=mode, block_min_tile=0)
    try:
        _, _, r = gpu_mul(ctx, f, g, nv, algo=pb.PK_ALGO_BLOCK, **opt)
        assert ctx.stats()["algo"] == pb.PK_ALGO_BLOCK
        assert_terms_equal(r, oracle_terms(po.p_mul(f, g, **kw), nv), f"{case} mode {mode}")
    finally:
        ctx.set_tuning(block_mode=0, block_min_tile=8)


def test_pearce2_full(ctx):
    """benchmarks/pearce2.cpp:57 -- the largest configuration with a known answer that one GPU call and the CPU oracle
    both finish in seconds: 20349^2 = 414 M products, 28 398 035 result terms (710 MB of result, ~10 GB of staging)."""
    (f, g), nv = po.pearce_operands(16), 5
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    da, db = ctx.upload(a, nv), ctx.upload(b, nv)
    r = ctx.mul_dev(da, db)
    st = ctx.stats()
    assert len(r) == 28398035
    assert st["algo"] == pb.PK_ALGO_BLOCK and st["term_products"] == 20349 ** 2
    # the parts of a 4-way partition add up to the same digest (checksum of checksums)
    parts = sum(ctx.mul_dev(da, db, part_index=p, part_count=4).digest() for p in range(4)) & ((1 << 64) - 1)
    assert parts == r.digest()
    key, mag, sign = r.download()
    del r
    assert_sorted_unique(key)
    ref = cpu_ref(a, b, nv, n_threads=os.cpu_count() or 8)
    assert_terms_equal((key, mag, sign), ref)
    assert pb.digest_terms(*ref) == parts


# ---------------------------------------------------------------------------------------------- wider coverage


def _random_poly(rs, nv, nterms, emin, emax, bits):
    p = {}
    while len(p) < nterms:
        e = tuple(int(x) for x in rs.randint(emin, emax + 1, size=nv))
        c = int.from_bytes(rs.bytes((bits + 7) // 8), "little") >> ((8 - bits % 8) % 8)
        c = c or 1
        p[e] = -c if rs.randint(2) else c
    return p


@pytest.mark.parametrize("bits_a,bits_b", [(64, 64), (100, 40), (40, 128), (128, 128)])
@pytest.mark.parametrize("algo", ALGOS)
def test_multi_limb_random(ctx, algo, bits_a, bits_b):
    rs = np.random.RandomState(bits_a * 1000 + bits_b)
    f = _random_poly(rs, 3, 300, -4, 9, bits_a)
    g = _random_poly(rs, 3, 200, -7, 5, bits_b)
    _, _, r = gpu_mul(ctx, f, g, 3, algo=algo)
    assert_terms_equal(r, oracle_terms(po.p_mul(f, g), 3))


@pytest.mark.parametrize("seed", range(24))
def test_random_products_every_strategy(ctx, seed):
    """Random shapes against the exact Python oracle through every accumulation strategy: variable counts 2..6,
    exponent boxes that put 31/32/33/64/65 terms in a group, lattice steps, negative exponents, mixed signs, widths up to
    64 bits, truncation on some seeds."""
    rs = np.random.RandomState(1000 + seed)
    nv = int(rs.randint(2, 7))
    lo = [int(x) for x in rs.randint(-3, 2, size=nv)]
    span = [int(x) for x in rs.choice([1, 2, 3, 5, 8, 31, 32, 33, 64, 65], size=nv)]
    span[0] = int(rs.choice([31, 32, 33, 64, 65, 96, 7]))  # digit 0 decides the tile shapes
    step = [int(x) for x in rs.choice([1, 1, 1, 2, 3], size=nv)]
    bits = int(rs.choice([3, 20, 40, 64]))

    def poly(nterms, sgn):
        p = {}
        for _ in range(nterms):
            e = tuple(lo[v] + step[v] * int(rs.randint(0, span[v])) for v in range(nv))
            c = int(rs.randint(1, 2 ** 31)) ** 3 % (2 ** bits - 1) + 1
            p[e] = -c if (sgn and rs.rand() < 0.4) else c
        return p

    f, g = poly(int(rs.randint(40, 1500)), seed % 2), poly(int(rs.randint(40, 1500)), seed % 3 == 0)
    opt, kw = {}, {}
    if seed % 4 == 3:
        D = int(np.median([sum(k) for k in f]) + np.median([sum(k) for k in g]))
        opt, kw = dict(trunc_mode=1, max_degree=D), dict(max_degree=D)
    expect = oracle_terms(po.p_mul(f, g, **kw), nv)
    ctx.set_tuning(block_min_tile=0)
    try:
        for algo in ALGOS:
            try:
                _, _, r = gpu_mul(ctx, f, g, nv, algo=algo, **opt)
            except MemoryError:
                assert algo == pb.PK_ALGO_DENSE  # a forced dense accumulator over a sparse 2^40 range is refused, not faked
                continue
            assert_terms_equal(r, expect, f"seed {seed} algo {algo}")
    finally:
        ctx.set_tuning(block_min_tile=8)


@pytest.mark.parametrize("w0,w1", [(31, 3), (32, 2), (33, 4), (64, 2), (65, 3), (96, 2), (1, 70), (130, 1)])
def test_full_boxes_every_strategy(ctx, w0, w1):
    """Operands that fill an exponent box: groups of exactly 31 / 32 / 33 / 64 / 65 / 96 terms hit every tile shape of the
    block kernel (64-term, 32-term and ragged chunks) with every lane occupied."""
    rs = np.random.RandomState(w0 * 100 + w1)
    f = {(a, b, c): int(rs.randint(1, 2 ** 40)) * (1 if (a + b) % 3 else -1) for a in range(w0) for b in range(w1) for c in range(5)}
    g = {(a, b, c): int(rs.randint(1, 2 ** 38)) for a in range(w0) for b in range(w1) for c in range(0, 9, 2)}
    expect = oracle_terms(po.p_mul(f, g), 3)
    ctx.set_tuning(block_min_tile=0)
    try:
        for algo in ALGOS:
            _, _, r = gpu_mul(ctx, f, g, 3, algo=algo)
            assert_terms_equal(r, expect, f"box {w0}x{w1} algo {algo}")
        for mode in (1, 2):
            ctx.set_tuning(block_mode=mode, block_target=64, block_permute=mode - 1)
            _, _, r = gpu_mul(ctx, f, g, 3, algo=pb.PK_ALGO_BLOCK)
            assert ctx.stats()["algo"] == pb.PK_ALGO_BLOCK
            assert_terms_equal(r, expect, f"box {w0}x{w1} block mode {mode}")
    finally:
        ctx.set_tuning(block_min_tile=8, block_mode=0, block_target=1024, block_permute=0)


def test_laurent_and_gcd_compaction(ctx):
    # exponents on a coarse lattice (steps 3 and 5) with negative powers: exercises the re-encoding with gcd steps
    rs = np.random.RandomState(7)
    f = {(3 * int(a), -5 * int(b)): int(c) + 1 for a, b, c in rs.randint(0, 30, size=(400, 3))}
    g = {(-3 * int(a), 5 * int(b) + 2): -(int(c) + 1) for a, b, c in rs.randint(0, 30, size=(400, 3))}
    for algo in ALGOS:
        _, _, r = gpu_mul(ctx, f, g, 2, algo=algo)
        assert_terms_equal(r, oracle_terms(po.p_mul(f, g), 2))


def test_many_variables(ctx):
    rs = np.random.RandomState(11)
    nv = 12
    f = _random_poly(rs, nv, 150, 0, 3, 30)
    g = _random_poly(rs, nv, 150, 0, 3, 30)
    _, _, r = gpu_mul(ctx, f, g, nv)
    assert_terms_equal(r, oracle_terms(po.p_mul(f, g), nv))


def test_commutative_and_operand_order(ctx):
    f, g = po.pearce_operands(5)
    a, b = terms(f, 5, 1), terms(g, 5, 2)
    assert_terms_equal(ctx.mul(a, b, 5), ctx.mul(b, a, 5))
    small = terms({(0, 0, 0, 0, 0): 7, (1, 0, 0, 0, 0): -7}, 5)
    assert_terms_equal(ctx.mul(a, small, 5), ctx.mul(small, a, 5))


@pytest.mark.parametrize("parts", [2, 3, 8])
@pytest.mark.parametrize("case", ["fateman", "pearce", "trunc"])
def test_output_partitions(ctx, parts, case):
    """Disjoint output ranges: the concatenation of the parts (in order) is the whole product."""
    opt = {}
    if case == "fateman":
        (f, g), nv = po.fateman_operands(8), 4
    elif case == "pearce":
        (f, g), nv = po.pearce_operands(6), 5
    else:
        (f, g), nv = po.fateman_operands(8), 4
        opt = dict(trunc_mode=1, max_degree=9)
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    da, db = ctx.upload(a, nv), ctx.upload(b, nv)
    whole = ctx.mul_dev(da, db, **opt)
    wk, wm, ws = whole.download()
    keys, mags, signs, digest, pairs = [], [], [], 0, 0
    for p in range(parts):
        part = ctx.mul_dev(da, db, part_index=p, part_count=parts, **opt)
        pairs += ctx.stats()["term_products"]
        k, m, s = part.download()
        digest = (digest + part.digest()) & ((1 << 64) - 1)
        keys.append(k)
        mags.append(np.pad(m, ((0, 0), (0, wm.shape[1] - m.shape[1]))) if len(k) else np.zeros((0, wm.shape[1]), np.uint64))
        signs.append(s)
    allk = np.concatenate(keys)
    assert_sorted_unique(allk)  # ordered, disjoint ranges
    assert_terms_equal((allk, np.concatenate(mags), np.concatenate(signs)), (wk, wm, ws))
    assert digest == whole.digest() == pb.digest_terms(wk, wm, ws)
    ctx.mul_dev(da, db, **opt)
    assert pairs == ctx.stats()["term_products"]  # every term product is done by exactly one partition
    if case != "trunc":
        sizes = [len(k) for k in keys]
        assert min(sizes) > 0


@pytest.mark.parametrize("parts", [2, 5])
def test_peer_exchange_kernels_one_gpu(ctx, parts):
    """pk_dpoly_publish_size / pk_dpoly_export_peers with every 'peer' on this GPU: each part pushes its terms into the
    result arrays of all peers at the offset derived from the size table; every peer ends up with the whole product."""
    import torch

    (f, g), nv = po.pearce_operands(6), 5
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    da, db = ctx.upload(a, nv), ctx.upload(b, nv)
    wk, wm, ws = ctx.mul_dev(da, db).download()
    n, L = len(wk), wm.shape[1]
    dev = torch.device("cuda", 0)
    tables = [torch.zeros(parts * 2, dtype=torch.int64, device=dev) for _ in range(parts)]
    keys = [torch.full((n + 8,), -1, dtype=torch.int64, device=dev) for _ in range(parts)]
    mags = [torch.full(((n + 8) * (L + 1),), -1, dtype=torch.int64, device=dev) for _ in range(parts)]
    signs = [torch.zeros(n + 8, dtype=torch.int8, device=dev) for _ in range(parts)]
    torch.cuda.synchronize()
    ps = [ctx.mul_dev(da, db, part_index=r, part_count=parts) for r in range(parts)]
    for r, part in enumerate(ps):
        part.publish_size([t.data_ptr() + 16 * r for t in tables])
    for cap_terms, cap_limbs, written in ((n - 1, L + 1, False), (n + 8, L - 1 if L > 1 else 0, False), (n + 8, L + 1, True)):
        for r, part in enumerate(ps):
            part.export_peers(r, tables[r].data_ptr(), [k.data_ptr() for k in keys], [m.data_ptr() for m in mags],
                              [x.data_ptr() for x in signs], cap_terms, cap_limbs)
        ctx.download(ps[0])  # synchronises the context's stream
        for j in range(parts):
            if not written:  # a result that does not fit is not written at all
                assert int(keys[j].max()) == -1 and int(signs[j].abs().sum()) == 0
                continue
            assert tables[j].cpu().view(parts, 2)[:, 0].tolist() == [len(x) for x in ps]
            got = (keys[j][:n].cpu().numpy(), mags[j][:n * L].cpu().numpy().view(np.uint64).reshape(n, L), signs[j][:n].cpu().numpy())
            assert_terms_equal(got, (wk, wm, ws))
            assert int(keys[j][n:].max()) == -1


def test_peer_exchange_two_gpus():
    """The peer-memory exchange between two processes / two GPUs (torch symmetric memory + this library's push kernel)
    against the whole product and against the NCCL exchange.  Needs two GPUs."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "dist_peer_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("PEER-EXCHANGE-OK") == 2, out.stdout[-3000:]


def test_device_resident_chain(ctx):
    # benchmarks/fateman1.hpp:46-48: f *= tmp, nine times, operands never leave HBM
    nv = 4
    base = {(0, 0, 0, 0): 1, (1, 0, 0, 0): 1, (0, 1, 0, 0): 1, (0, 0, 1, 0): 1, (0, 0, 0, 1): 1}
    tmp = ctx.upload(terms(base, nv), nv)
    f = ctx.upload(terms(base, nv), nv)
    for _ in range(9):
        f = ctx.mul_dev(f, tmp)
    ref, _ = po.fateman_operands(10)
    assert_terms_equal(f.download(), oracle_terms(ref, nv))
    sq = ctx.mul_dev(f, f)
    assert_terms_equal(sq.download(), oracle_terms(po.p_pow_multinomial(po._linear(4, [1, 1, 1, 1]), 20), nv))


def test_trim_releases_and_products_still_work():
    import torch

    c2 = pb.Context(0)
    try:
        (f, g), nv = po.pearce_operands(6), 5
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        r1 = c2.mul(a, b, nv)
        free_before = torch.cuda.mem_get_info(0)[0]
        c2.trim()
        free_after = torch.cuda.mem_get_info(0)[0]
        assert free_after >= free_before  # the staging arrays went back to the driver
        assert_terms_equal(c2.mul(a, b, nv), r1)
    finally:
        c2.close()


def test_custom_limits_table(ctx):
    """pk_ctx_set_limits: a caller whose standard library produced a different Kronecker table."""
    c2 = pb.Context(0)
    try:
        M = [100, 200, 300]
        c2.set_limits(3, M)
        lim, hmin, hmax = c2.limits(3)
        assert list(lim) == M

        def enc(e):
            code, c = 0, 1
            for x, m in zip(e, M):
                code += x * c
                c *= 2 * m + 1
            return code

        f = {(1, 2, 3): 5, (0, -7, 4): -2}
        g = {(10, 20, 30): 3, (-1, 1, 0): 11}
        mk = lambda p: (np.array([enc(e) for e in p], np.int64), np.array([[abs(c)] for c in p.values()], np.uint64),
                        np.array([1 if c > 0 else -1 for c in p.values()], np.int8))
        k, m, s = c2.mul(mk(f), mk(g), 3)
        ref = po.p_mul(f, g)
        got = {int(kk): (int(mm[0]) if ss > 0 else -int(mm[0])) for kk, mm, ss in zip(k, m, s)}
        assert got == {enc(e): c for e, c in ref.items()}
        with pytest.raises(OverflowError):
            c2.mul(mk({(100, 0, 0): 1}), mk({(1, 0, 0): 1}), 3)
    finally:
        c2.close()


# ---------------------------------------------------------------------------------------------- multi-device context


def _multi_devices(n):
    """n sub-contexts: real GPUs when the box has them, else all on GPU 0 (same code path, parts share the device)."""
    import torch

    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n_dev", [1, 2, 3, 8])
def test_multi_ctx_host_product(ctx, n_dev):
    """pk_ctx_create_multi + pk_mul: N ordered parts land in ONE host result that equals the single-device product and the
    oracle, term by term (reduced Pearce 591235, reduced Fateman with cancellation 5786, truncated Fateman)."""
    m = pb.Context.multi(_multi_devices(n_dev))
    try:
        assert m.device_count == n_dev
        (f, g), nv = po.pearce_operands(8), 5
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        out = m.mul(a, b, nv)
        assert_sorted_unique(out[0])
        assert len(out[0]) == 591235  # tests/base_series_multiplier.cpp:571-605
        assert_terms_equal(out, cpu_ref(a, b, nv))
        st = m.stats()
        assert st["term_products"] == len(a[0]) * len(b[0]) and st["result_terms"] == 591235
        # cancellation: (1+x+y+z+t)^10 * ((1-x+y+z+t)^10) has 5786 terms (polynomial_multiplier_01.cpp:218-231)
        f2, g2 = po.fateman_cancel_operands(10)
        a2, b2 = terms(f2, 4, 3), terms(g2, 4, 4)
        out2 = m.mul(a2, b2, 4)
        assert_terms_equal(out2, ctx.mul(a2, b2, 4))
        assert_terms_equal(out2, oracle_terms(po.p_mul(f2, g2), 4))
        # truncation through the partitioned path
        ff, gg = po.fateman_operands(10)
        a3, b3 = terms(ff, 4, 5), terms(gg, 4, 6)
        out3 = m.mul(a3, b3, 4, trunc_mode=1, max_degree=12)
        assert_terms_equal(out3, cpu_ref(a3, b3, 4, trunc_mode=1, max_degree=12))
        # empty operand: empty product
        empty = (np.zeros(0, np.int64), np.zeros((0, 1), np.uint64), np.zeros(0, np.int8))
        assert len(m.mul(empty, a3, 4)[0]) == 0
    finally:
        m.close()


def test_multi_ctx_device_resident_chain(ctx):
    """Products of a multi-device context stay partitioned in HBM and are replicated peer-to-peer when they become
    operands: f *= tmp nine times (benchmarks/fateman1.hpp:46-48), then a square, against the closed forms."""
    m = pb.Context.multi(_multi_devices(3))
    try:
        nv = 4
        base = {(0, 0, 0, 0): 1, (1, 0, 0, 0): 1, (0, 1, 0, 0): 1, (0, 0, 1, 0): 1, (0, 0, 0, 1): 1}
        tmp = m.upload(terms(base, nv), nv)
        f = m.upload(terms(base, nv), nv)
        for _ in range(9):
            f = m.mul_dev(f, tmp)
        ref, _ = po.fateman_operands(10)
        assert len(f) == len(ref)
        assert_terms_equal(f.download(), oracle_terms(ref, nv))
        sq = m.mul_dev(f, f)
        want = oracle_terms(po.p_pow_multinomial(po._linear(4, [1, 1, 1, 1]), 20), nv)
        got = sq.download()
        assert_sorted_unique(got[0])
        assert_terms_equal(got, want)
        assert sq.digest() == pb.digest_terms(*want)
        # the same product on the single-device context has the same digest
        sf = ctx.upload(oracle_terms(ref, nv), nv)
        assert ctx.mul_dev(sf, sf).digest() == sq.digest()
    finally:
        m.close()


def test_multi_ctx_argument_errors(ctx):
    m = pb.Context.multi(_multi_devices(2))
    try:
        (f, g), nv = po.pearce_operands(4), 5
        a, b = terms(f, nv, 1), terms(g, nv, 2)
        with pytest.raises(ValueError):
            m.mul(a, b, nv, part_index=0, part_count=2)  # the context partitions the product itself
        da_single = ctx.upload(a, nv)
        da_multi, db_multi = m.upload(a, nv), m.upload(b, nv)
        with pytest.raises(ValueError):
            m.mul_dev(da_single, db_multi)  # polynomials belong to the context that created them
        with pytest.raises(ValueError):
            ctx.mul_dev(da_multi, da_single)
        with pytest.raises(OverflowError):  # check_bounds through the multi-device handle (polynomial_multiplier_01.cpp:120-174)
            lim, _, _ = m.limits(1)
            big = terms({(int(lim[0]),): 1}, 1)
            m.mul(big, terms({(1,): 1}, 1), 1)
        assert_terms_equal(m.mul_dev(da_multi, db_multi).download(), ctx.mul(a, b, nv))
    finally:
        m.close()


def test_upload_rejects_invalid_terms(ctx):
    """pk_upload checks what the reference's containers guarantee: keys inside the codification's range (decode throws
    std::invalid_argument otherwise, kronecker_array.hpp:329-365), signs in {-1, 0, +1}, every key once (series.hpp:1312-1320)."""
    lim, hmin, hmax = ctx.limits(2)
    ok = terms({(1, 2): 3, (0, 1): -5}, 2)
    for bad_key in (hmax + 1, hmin - 1):
        k = ok[0].copy()
        k[0] = bad_key
        with pytest.raises(ValueError):
            ctx.upload((k, ok[1], ok[2]), 2)
    s = ok[2].copy()
    s[1] = 2
    with pytest.raises(ValueError):
        ctx.upload((ok[0], ok[1], s), 2)
    k = ok[0].copy()
    k[1] = k[0]
    with pytest.raises(ValueError):
        ctx.upload((k, ok[1], ok[2]), 2)
    with pytest.raises(ValueError):
        ctx.mul((k, ok[1], ok[2]), ok, 2)
    # a zero coefficient is accepted and contributes nothing
    z, zm = ok[2].copy(), ok[1].copy()
    z[0] = 0
    zm[0] = 0
    r = ctx.mul((ok[0], zm, z), ok, 2)
    full = {(1, 2): 3, (0, 1): -5}
    keep = {e: c for e, c in full.items() if po.kron_encode(e) != int(ok[0][0])}
    assert_terms_equal(r, oracle_terms(po.p_mul(keep, full), 2))


def _random_wide_poly(rs, n, bits, nv=3, signed=True):
    p = {}
    while len(p) < n:
        e = tuple(int(x) for x in rs.randint(0, 7, size=nv))
        c = int.from_bytes(rs.bytes((bits + 7) // 8), "little") >> ((8 - bits % 8) % 8) or 1
        p[e] = -c if (signed and rs.randint(2)) else c
    return p


@pytest.mark.parametrize("bits_a,bits_b", [(200, 60), (256, 256), (130, 129), (129, 1), (192, 190)])
def test_wide_operands(ctx, bits_a, bits_b):
    """Operands of three and four limbs (the width protocol of the C ABI: 128-bit halves through the pairing kernels, summed
    with shifts on the device), against the exact Python oracle; also as an output partition and truncated."""
    rs = np.random.RandomState(bits_a * 1000 + bits_b)
    nv = 3
    f, g = _random_wide_poly(rs, 300, bits_a), _random_wide_poly(rs, 250, bits_b)
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    assert a[1].shape[1] == (bits_a + 63) // 64
    want = oracle_terms(po.p_mul(f, g), nv)
    got = ctx.mul(a, b, nv)
    assert_sorted_unique(got[0])
    assert_terms_equal(got, want)
    parts = [ctx.mul(a, b, nv, part_index=r, part_count=3) for r in range(3)]
    L = max(p[1].shape[1] for p in parts)
    cat = (np.concatenate([p[0] for p in parts]), np.concatenate([np.pad(p[1], ((0, 0), (0, L - p[1].shape[1]))) for p in parts]),
           np.concatenate([p[2] for p in parts]))
    assert_sorted_unique(cat[0])
    assert_terms_equal(cat, want)
    got_t = ctx.mul(a, b, nv, trunc_mode=1, max_degree=9)
    assert_terms_equal(got_t, oracle_terms(po.p_mul(f, g, max_degree=9), nv))


def test_wide_chain_and_limit(ctx):
    """benchmarks/fateman1_dynamic.cpp: operands scaled by 2^64 - 1; the product (4 limbs) is multiplied on, device resident;
    beyond four operand limbs the library reports PK_E_CF_WIDTH (std::overflow_error), it never wraps."""
    nv = 4
    f, _ = po.fateman_operands(4)
    k = 2 ** 64 - 1
    f = {e: c * k for e, c in f.items()}
    g = dict(f)
    g[(0, 0, 0, 0)] = g.get((0, 0, 0, 0), 0) + 1
    df, dg = ctx.upload(terms(f, nv, 1), nv), ctx.upload(terms(g, nv, 2), nv)
    p1 = ctx.mul_dev(df, dg)           # two-limb operands: the pairing kernels
    p2 = ctx.mul_dev(p1, df)           # three-limb operand x two-limb operand: halves
    ref1 = po.p_mul(f, g)
    ref2 = po.p_mul(ref1, f)
    assert_terms_equal(p1.download(), oracle_terms(ref1, nv))
    assert_terms_equal(p2.download(), oracle_terms(ref2, nv))
    p3 = ctx.mul_dev(p2, p1)           # four-limb x three-limb operands: still inside the protocol (result of six limbs)
    assert_terms_equal(p3.download(), oracle_terms(po.p_mul(ref2, ref1), nv))
    assert p1.limbs == 3 and p2.limbs == 4 and p3.limbs > 4
    with pytest.raises(OverflowError):
        ctx.mul_dev(p3, p3)            # operands of more than four limbs: reported, never wrapped
    m = pb.Context.multi(_multi_devices(2))
    try:
        rs = np.random.RandomState(3)
        fa, fb = _random_wide_poly(rs, 200, 250), _random_wide_poly(rs, 200, 250)
        assert_terms_equal(m.mul(terms(fa, 3, 1), terms(fb, 3, 2), 3), oracle_terms(po.p_mul(fa, fb), 3))
    finally:
        m.close()


def test_unsupported_sizes_have_their_own_status(ctx):
    """A product the reference would compute but whose re-encoded code range reaches 2^62 is reported as PK_E_UNSUPPORTED (a
    runtime error), not as the overflow_error of the bounds check: nothing overflowed."""
    big = 12 * 10 ** 17  # the 1-variable Kronecker limit is 3.0e18: sums of two such exponents stay inside it
    f = {(-big,): 1, (0,): 1, (1,): 1, (big,): 1}
    with pytest.raises(pb.PkError) as ei:
        ctx.mul(terms(f, 1), terms(f, 1), 1)  # code range 4.8e18 + 1 >= 2^62
    assert ei.value.code == 7
    # ... while exponents that leave the Kronecker limits still raise the reference's overflow_error
    lim, _, _ = ctx.limits(1)
    with pytest.raises(OverflowError):
        ctx.mul(terms({(int(lim[0]),): 1, (0,): 1}, 1), terms(f, 1), 1)


# ---------------------------------------------------------------------------------------------- state kept across products


def _boxed_sparse_poly(seed, nv, n, side, negate=False):
    """n distinct terms of the box [0, side)^nv, both corners included (every seed has the same exponent bounds, so the
    same code range, unit size and term count: what the context's size hint compares)."""
    rs = np.random.RandomState(seed)
    codes = set(int(c) for c in rs.choice(side ** nv - 2, size=n - 2, replace=False) + 1) | {0, side ** nv - 1}
    d = {}
    for c in sorted(codes):
        e = []
        for _ in range(nv):
            e.append(c % side)
            c //= side
        v = int(rs.randint(1, 1000))
        d[tuple(e)] = -v if negate and (e[0] & 1) else v
    return d


@pytest.mark.parametrize("negate", [False, True])
def test_repeated_interleaved_and_recycled_operands(negate):
    """State a context keeps from one product to the next -- the size hint (a product of the same operand objects as the
    last one sizes its result from it; the device verifies the count and the host repeats the long way on a miss), the
    bitmap whose bits the accumulating kernel clears itself, the cache of idle device blocks -- never changes the bits:
    repeated products, products of other shapes in between, and new operands that take the place (device blocks and
    usually the handle address) of freed ones of the same shape."""
    nv = 4
    c2 = pb.Context(0)
    try:
        c2.set_tuning(block_min_tile=0)
        b = terms(_boxed_sparse_poly(100, nv, 2200, 40, negate), nv, 2)
        db = c2.upload(b, nv)
        (pf, pg), pnv = po.pearce_operands(5), 5
        pa, pb_ = terms(pf, pnv, 1), terms(pg, pnv, 2)
        other = None
        refs = {}
        modes = set()
        for round_ in range(3):
            for seed in (1, 2, 3):
                a = terms(_boxed_sparse_poly(seed, nv, 2500, 40, negate), nv, 1)
                if seed not in refs:
                    refs[seed] = cpu_ref(a, b, nv, n_threads=4)
                da = c2.upload(a, nv)
                for rep in range(3 if round_ == 0 else 2):  # the second and third time the hint holds
                    r = c2.mul_dev(da, db, algo=pb.PK_ALGO_BLOCK)
                    modes.add(c2.stats()["retries"] & 1)
                    assert_terms_equal(r.download(), refs[seed])
                    r.free()
                if round_ == 1:  # a product of another shape (other code range, other bitmap extent) in between
                    out = c2.mul(pa, pb_, pnv, algo=pb.PK_ALGO_BLOCK)
                    other = other or cpu_ref(pa, pb_, pnv, n_threads=4)
                    assert_terms_equal(out, other)
                da.free()  # the next operand of the same shape takes its place
        assert 1 in modes  # (the sparse, bitmap form of the block kernel is what ran)
        # the miss path, forced: every product is sized from the one before it, whatever its operands were
        c2.set_tuning(block_size_hint=2)
        (ff, fg), fnv = po.fateman_operands(6), 4
        fa, fb = terms(ff, fnv, 1), terms(fg, fnv, 2)
        dense_ref = cpu_ref(fa, fb, fnv, n_threads=4)
        for seed in (1, 2, 3, 1):
            a = terms(_boxed_sparse_poly(seed, nv, 2500, 40, negate), nv, 1)
            assert_terms_equal(c2.mul(a, b, nv, algo=pb.PK_ALGO_BLOCK), refs[seed])
            assert_terms_equal(c2.mul(fa, fb, fnv, algo=pb.PK_ALGO_BLOCK), dense_ref)  # (dense form: the gather is sized by the hint)
            assert_terms_equal(c2.mul(pa, pb_, pnv, algo=pb.PK_ALGO_BLOCK), other)
    finally:
        c2.close()


@pytest.mark.parametrize("shift,bits", [(0, 1), (12, 8), (3, 16), (60, 4)])
def test_download_grouped(ctx, shift, bits):
    """pk_download_grouped: the same terms, ordered by ((uint64) key >> shift) & (2^bits - 1) and by key inside a group
    (what a power-of-two hash table with the identity hash of kronecker_monomial wants to be filled from)."""
    (f, g), nv = po.pearce_operands(5), 5
    g = {k: (-v if sum(k) % 3 == 0 else v) for k, v in g.items()}
    a, b = terms(f, nv, 1), terms(g, nv, 2)
    r = ctx.mul_dev(ctx.upload(a, nv), ctx.upload(b, nv))
    plain = r.download()
    key, mag, sign = ctx.download_grouped(r, shift, bits)
    group = (key.astype(np.uint64) >> np.uint64(shift)) & np.uint64((1 << bits) - 1)
    order = np.lexsort((key, group))
    assert np.array_equal(order, np.arange(len(key)))  # already in (group, key) order
    back = np.argsort(key, kind="stable")
    assert_terms_equal((key[back], mag[back], sign[back]), plain)
    # the same in pieces, with the progress callback after each
    seen = []
    offsets = np.full((1 << bits) + 1, 12345, dtype=np.uint64)
    k2, m2, s2 = ctx.download_grouped(r, shift, bits, n_chunks=7, progress=seen.append, offsets=offsets)
    assert seen == [len(key) * (c + 1) // 7 for c in range(7)]
    assert np.array_equal(k2, key) and np.array_equal(m2, mag) and np.array_equal(s2, sign)
    assert np.array_equal(offsets, np.searchsorted(group, np.arange((1 << bits) + 1, dtype=np.uint64), side="left"))  # first term of every group
    empty = ctx.mul_dev(ctx.upload(terms({}, nv), nv), ctx.upload(b, nv))
    assert len(ctx.download_grouped(empty, shift, bits)[0]) == 0
    assert len(ctx.download_grouped(empty, shift, bits, n_chunks=3, progress=seen.append)[0]) == 0 and len(seen) == 7
    with pytest.raises(ValueError):
        ctx.download_grouped(r, 0, 0)
    with pytest.raises(ValueError):
        ctx.download_grouped(r, 64, 8)
