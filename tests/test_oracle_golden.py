"""Pins the CPU oracles (oracle/pyoracle.py, oracle/piranha_cpu.c) against every known answer the reference's own
tests and benchmarks hold for the multiplier path.  Runs without a GPU."""
import json
import os

import numpy as np
import pytest

from oracle import cpu_oracle, pyoracle as po
from polyhelper import as_pairs, assert_terms_equal, cpu_ref, oracle_terms, terms

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# ---------------------------------------------------------------------------------------------- kronecker_array


def test_limits_structure():
    # tests/kronecker_array.cpp:52-83 checks structure only (sizes, signs); the values are whatever libstdc++ yields
    tab = po.kron_limits_table()
    assert len(tab) > 1 and tab[0] == ((), 0, 0, 0)
    for m in range(1, len(tab)):
        M, hmin, hmax, diff = tab[m]
        assert len(M) == m and all(x > 0 for x in M)
        assert hmin < 0 < hmax and diff == hmax - hmin and diff + 1 < 2**63
        # h_max is the code of the all-max vector
        c, acc = 1, 0
        for x in M:
            acc += x * c
            c *= 2 * x + 1
        assert acc == hmax == -hmin


def test_limits_recorded_values():
    # values recorded in SURVEY.md section 8c for this image's libstdc++ 13 (not pinned by the reference)
    tab = po.kron_limits_table()
    assert tab[4][0] == (19319, 19851, 21635, 19661)
    assert tab[5][0] == (2493, 2504, 2753, 2430, 2609)
    with open(os.path.join(GOLDEN, "kronecker_limits.json")) as fh:
        gold = json.load(fh)
    assert [list(t[0]) for t in tab] == gold["limits"]
    assert [t[1] for t in tab] == gold["h_min"]


def test_encode_decode_roundtrip():
    # tests/kronecker_array.cpp:86-160: min / max / random vectors round-trip; out-of-range throws
    rs = np.random.RandomState(0)
    for m in (1, 2, 3, 4, 5, 8, 12):
        M = po.kron_limits(m)[0]
        vecs = [tuple(-x for x in M), tuple(M), (0,) * m]
        vecs += [tuple(int(rs.randint(-x, x + 1)) for x in M) for _ in range(50)]
        for v in vecs:
            assert po.kron_decode(po.kron_encode(v), m) == v
        assert po.kron_encode(tuple(-x for x in M)) == po.kron_limits(m)[1]
        assert po.kron_encode(tuple(M)) == po.kron_limits(m)[2]
        bad = list(M)
        bad[0] += 1
        with pytest.raises(ValueError):
            po.kron_encode(bad)
    assert po.kron_encode(()) == 0 and po.kron_decode(0, 0) == ()
    with pytest.raises(ValueError):
        po.kron_decode(1, 0)


def test_key_multiply_is_code_addition():
    # kronecker_monomial::multiply (kronecker_monomial.hpp:427-436): key = key1 + key2, cf = cf1 * cf2;
    # fixtures of tests/kronecker_monomial_01.cpp:404-461
    assert po.kron_encode((1,)) + po.kron_encode((2,)) == po.kron_encode((3,))
    assert po.kron_encode((1, -1)) + po.kron_encode((2, 0)) == po.kron_encode((3, -1))
    r = cpu_ref(terms({(1, -1): 2}, 2), terms({(2, 0): -4}, 2), 2)
    assert as_pairs(*r) == [(po.kron_encode((3, -1)), -8)]


# ---------------------------------------------------------------------------------------------- known-answer sizes


def test_reduced_fateman_sizes():
    f, g = po.fateman_operands(10)
    prod = po.p_mul(f, g)
    assert len(f) == 1001 and len(prod) == 10626  # tests/polynomial_multiplier_01.cpp:214
    assert prod == po.fateman_product_closed_form(10)
    f, h = po.fateman_cancel_operands(10)
    assert len(po.p_mul(f, h)) == 5786  # :230


@pytest.mark.parametrize("minus_u,size", [(False, 591235), (True, 591184)])
def test_reduced_pearce_sizes(minus_u, size):
    f, g = po.pearce_operands(8, minus_u)
    a, b = terms(f, 5), terms(g, 5, 1)
    r = cpu_ref(a, b, 5, n_threads=4)
    assert len(r[0]) == size  # tests/base_series_multiplier.cpp:586, :604
    if not minus_u:
        assert_terms_equal(r, oracle_terms(po.p_mul(f, g), 5), "C oracle vs Python oracle")


@pytest.mark.parametrize("name,builder,nv,size", [
    ("fateman1", lambda: po.fateman_operands(20), 4, 135751),  # benchmarks/fateman1.cpp:55
    ("pearce1", lambda: po.pearce_operands(12), 5, 5821335),  # benchmarks/pearce1.cpp:56
    ("monagan1", lambda: po.monagan_operands(1), 3, 12341),  # benchmarks/monagan1.cpp:51
    ("monagan2", lambda: po.monagan_operands(2), 3, 12341),
    ("monagan3", lambda: po.monagan_operands(3), 3, 39711),
    ("monagan4", lambda: po.monagan_operands(4), 4, 135751),
    ("monagan5", lambda: po.monagan_operands(5), 5, 417311),  # benchmarks/monagan5.cpp:49
])
def test_benchmark_sizes(name, builder, nv, size):
    f, g = builder()
    r = cpu_oracle.cpu_mul(terms(f, nv), terms(g, nv, 1), nv, n_threads=os.cpu_count() or 1)
    assert len(r.key) == size
    if name == "fateman1":
        assert_terms_equal(r.canonical(), oracle_terms(po.fateman_product_closed_form(20), 4), "closed form")


def test_fateman2_size_and_width():
    f, g = po.fateman_operands(30)
    assert len(f) == 46376
    r = cpu_oracle.cpu_mul(terms(f, 4), terms(g, 4, 1), 4, n_threads=os.cpu_count() or 1)
    assert len(r.key) == 635376  # benchmarks/fateman2.cpp:55
    k, m, s = r.canonical()
    assert_terms_equal((k, m, s), oracle_terms(po.fateman_product_closed_form(30), 4), "closed form")
    assert max(abs(c) for c in po.soa_coefficients(m, s)).bit_length() == 128  # 2^127.96, SURVEY.md section 7.3


@pytest.mark.parametrize("D,size,pairs", [(20, 10626, 3108105), (30, 46376, 38970998)])
def test_truncated_fateman1(D, size, pairs):
    f, g = po.fateman_operands(20)
    r = cpu_oracle.cpu_mul(terms(f, 4), terms(g, 4, 1), 4, n_threads=os.cpu_count() or 1, trunc_mode=1, max_degree=D)
    assert len(r.key) == size  # benchmarks/fateman1_unpacked_truncation.cpp:57, :59
    da = np.array([sum(e) for e in f])
    db = np.sort(np.array([sum(e) for e in g]))
    assert int(np.searchsorted(db, D - da, side="right").sum()) == pairs  # SURVEY.md section 8 size table


# ---------------------------------------------------------------------------------------------- behaviour


def test_thread_count_sweep():
    # tests/polynomial_multiplier_02.cpp:55-96: same product with 1..4 threads, min_work_per_thread = 1
    f, g = po.pearce_operands(4)
    a, b = terms(f, 5), terms(g, 5, 1)
    ref = cpu_ref(a, b, 5, n_threads=1, min_work=1)
    assert_terms_equal(ref, oracle_terms(po.p_mul(f, g), 5))
    for nt in (2, 3, 4):
        assert_terms_equal(cpu_ref(a, b, 5, n_threads=nt, min_work=1), ref, f"{nt} threads")


def test_truncation_exact_cases():
    # tests/polynomial_truncation.cpp:75-170 (variables t,x,y,z -> positions 0..3)
    one, X, Y, Z, T = (0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (1, 0, 0, 0)

    def add(*es):
        return tuple(sum(c) for c in zip(*es))

    f = {one: 1, X: 1, Y: 1}
    expect = {
        1: {X: 2, Y: 2, one: 1},
        2: {X: 2, Y: 2, one: 1, add(X, Y): 2, add(X, X): 1, add(Y, Y): 1},
        3: {X: 2, Y: 2, one: 1, add(X, Y): 2, add(X, X): 1, add(Y, Y): 1},
        0: {one: 1},
        -1: {},
    }
    for D, want in expect.items():
        assert po.p_mul(f, f, max_degree=D) == want
        for nt in (1, 4):
            r = cpu_ref(terms(f, 4), terms(f, 4, 1), 4, n_threads=nt, min_work=1, trunc_mode=1, max_degree=D)
            assert_terms_equal(r, oracle_terms(want, 4), f"D={D} nt={nt}")
    lin = {T: 1, X: 1, Y: 1, Z: 1}
    assert po.p_mul(lin, lin, max_degree=1) == {}
    # partial degree in x only: (x+y+1)^2 truncated at 1 -> 2xy + 2x + y^2 + 2y + 1
    want = {add(X, Y): 2, X: 2, add(Y, Y): 1, Y: 2, one: 1}
    assert po.p_mul(f, f, max_degree=1, mask=[0, 1, 0, 0]) == want
    r = cpu_ref(terms(f, 4), terms(f, 4, 1), 4, trunc_mode=2, max_degree=1, deg_mask=[0, 1, 0, 0])
    assert_terms_equal(r, oracle_terms(want, 4))
    # partial degree in z: nothing is cut
    assert po.p_mul(f, f, max_degree=1, mask=[0, 0, 0, 1]) == po.p_mul(f, f)
    # partial degree in {x, y}
    assert po.p_mul(f, f, max_degree=1, mask=[0, 1, 1, 0]) == {X: 2, Y: 2, one: 1}
    # (x+y+z+t)^2 with partial degree 1 in x loses only x^2
    want = po.p_mul(lin, lin)
    del want[add(X, X)]
    assert po.p_mul(lin, lin, max_degree=1, mask=[0, 1, 0, 0]) == want


def test_bounds_fixtures():
    # tests/polynomial_multiplier_01.cpp:120-174
    M = po.kron_limits(3)[0]
    top = {tuple(M): 1}
    for v in range(3):
        e = [0, 0, 0]
        e[v] = 1
        with pytest.raises(OverflowError):
            po.check_bounds(top, {tuple(e): 1}, 3)
        with pytest.raises(OverflowError):
            cpu_oracle.cpu_mul(terms(top, 3), terms({tuple(e): 1}, 3), 3)
        below = list(M)
        below[v] -= 1
        r = cpu_ref(terms({tuple(below): 1}, 3), terms({tuple(e): 1}, 3), 3)
        assert as_pairs(*r) == [(po.kron_encode(M), 1)]
    r = cpu_ref(terms({(0, 0, 0): 2}, 3), terms({(0, 0, 0): 3}, 3), 3)  # pt{2} * pt{3} == 6
    assert as_pairs(*r) == [(0, 6)]


def test_empty_operands():
    empty = (np.zeros(0, np.int64), np.zeros((0, 1), np.uint64), np.zeros(0, np.int8))
    x = terms({(1,): 1}, 1)
    for a, b in ((empty, x), (x, empty), (empty, empty)):
        assert len(cpu_oracle.cpu_mul(a, b, 1).key) == 0


def test_two_limb_operands_and_signs():
    rs = np.random.RandomState(5)
    f = {(int(a), int(b)): (int(c) - 50) * (1 << 70) + 12345 for a, b, c in rs.randint(0, 100, size=(150, 3))}
    g = {(int(a), int(b)): (int(c) - 50) * (1 << 90) - 999 for a, b, c in rs.randint(0, 100, size=(120, 3))}
    f = {k: v for k, v in f.items() if v}
    g = {k: v for k, v in g.items() if v}
    r = cpu_ref(terms(f, 2), terms(g, 2, 1), 2)
    assert_terms_equal(r, oracle_terms(po.p_mul(f, g), 2))


def test_golden_products():
    """Committed fixtures (tests/golden/make_golden.py): small products with every coefficient pinned."""
    with open(os.path.join(GOLDEN, "products.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        nv = case["n_vars"]
        f = {tuple(e): int(c) for e, c in case["a"]}
        g = {tuple(e): int(c) for e, c in case["b"]}
        want = {tuple(e): int(c) for e, c in case["product"]}
        kw = {}
        if case.get("max_degree") is not None:
            kw = dict(max_degree=case["max_degree"], mask=case.get("mask"))
        assert po.p_mul(f, g, **kw) == want, case["name"]
        ckw = {}
        if kw:
            ckw = dict(trunc_mode=2 if case.get("mask") else 1, max_degree=case["max_degree"], deg_mask=case.get("mask"))
        assert_terms_equal(cpu_ref(terms(f, nv), terms(g, nv, 1), nv, **ckw), oracle_terms(want, nv), case["name"])
