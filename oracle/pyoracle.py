"""
TEST INFRASTRUCTURE ONLY -- exact, slow, independent CPU oracle (pure Python big integers).

Nothing under ``piranha_b200/`` may import this module.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker.

What it restates (citations are relative to /root/reference/include/piranha/):

* ``kronecker_array<std::int64_t>``: the limits table (``kronecker_array.hpp:128-229``), ``encode``
  (``:274-307``) and ``decode`` (``:329-365``).  The limits come out of a perturbed-prime search driven by
  ``std::mt19937(m)`` and ``std::uniform_int_distribution<int>(-5, 5)`` (``:131-137``); the distribution is
  restated here as libstdc++ >= 11 implements it (Lemire's nearly-divisionless method on a 32-bit
  generator), because the reference's numeric table is whatever the local standard library produces.
* term-by-term multiplication: ``kronecker_monomial::multiply`` (``kronecker_monomial.hpp:427-436``) is
  ``cf = cf1 * cf2`` and ``key = key1 + key2``; accumulation and removal of zero coefficients follow
  ``series_multiplier<polynomial>::task_consume`` and ``sanitise_series``
  (``polynomial.hpp:1723-1759``, ``base_series_multiplier.hpp:855-949``).
* truncation: degree by unpacking (``kronecker_monomial.hpp:1083-1119``), a product term is kept iff
  ``d1 + d2 <= max_degree`` (``polynomial.hpp:1402-1449``, ``:1477-1531``).
* the benchmark operands of ``benchmarks/fateman1.hpp``, ``fateman2.hpp``, ``pearce1.hpp``,
  ``monagan.hpp`` and the reduced ones used by ``tests/polynomial_multiplier_01.cpp:203-231`` and
  ``tests/base_series_multiplier.cpp:571-605``.

Parity pinning: the reference cannot be built in this image (no Boost / mp++ / gmp.h), so this oracle is
pinned against every known answer the reference's own tests and benchmarks hold for the path (term counts,
tiny exact products, overflow fixtures) -- see ``tests/test_oracle_golden.py``.  Individual coefficients of
the big products and the numeric limits table are NOT pinned by the reference; they are pinned here by
closed forms (multinomials) and by cross-checking this file against the independent C restatement in
``oracle/piranha_cpu.c`` and the C++ host mirror's ``kronecker_array``.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

Poly = Dict[Tuple[int, ...], int]

# --------------------------------------------------------------------------------------------------
# std::mt19937 + libstdc++ uniform_int_distribution<int>
# --------------------------------------------------------------------------------------------------


class MT19937:
    """32-bit Mersenne twister with std::mt19937's seeding (init_genrand)."""

    def __init__(self, seed: int):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        self.idx = 0

    def __call__(self) -> int:
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def libstdcxx_uniform_int(engine: MT19937, a: int, b: int) -> int:
    """uniform_int_distribution<int>(a, b)(engine) as libstdc++ (GCC >= 11) evaluates it for a 32-bit URNG."""
    rng = b - a + 1
    assert 0 < rng < (1 << 32)
    product = engine() * rng
    low = product & 0xFFFFFFFF
    if low < rng:
        threshold = ((1 << 32) - rng) % rng
        while low < threshold:
            product = engine() * rng
            low = product & 0xFFFFFFFF
    return (product >> 32) + a


# --------------------------------------------------------------------------------------------------
# nextprime (what mppp::integer::nextprime / mpz_nextprime returns for these magnitudes)
# --------------------------------------------------------------------------------------------------

_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:  # deterministic below 3.3e24
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def nextprime(n: int) -> int:
    n += 1
    while not _is_prime(n):
        n += 1
    return n


# --------------------------------------------------------------------------------------------------
# kronecker_array<int64_t> limits / encode / decode
# --------------------------------------------------------------------------------------------------

_I64_MIN, _I64_MAX = -(1 << 63), (1 << 63) - 1


def _fits_i64(x: int) -> bool:
    return _I64_MIN <= x <= _I64_MAX


def _cdiv(a: int, b: int) -> int:
    """C++ / GMP truncating division (operands here are positive, kept explicit for safety)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _determine_limit(m: int):
    """kronecker_array.hpp:128-215 restated.  Returns (M_vec, h_min, h_max, h_max - h_min) or empty M_vec."""
    engine = MT19937(m)

    def perturb(arg: int) -> int:
        arg += _cdiv(libstdcxx_uniform_int(engine, -5, 5) * arg, 100)
        return nextprime(arg)

    c_vec = [1]
    m_vec = [-1]
    M_vec = [1]
    for _ in range(1, m):
        m_vec.append(-1)
        M_vec.append(1)
        c_vec.append(c_vec[-1] * 3)
    prev = None
    while True:
        h_min = sum(c * v for c, v in zip(c_vec, m_vec))
        h_max = sum(c * v for c, v in zip(c_vec, M_vec))
        diff = h_max - h_min
        if not (_fits_i64(h_min) and _fits_i64(h_max) and _fits_i64(diff + 1)):
            if prev is not None:
                pc, pm, pM = prev
                h_min = sum(c * v for c, v in zip(pc, pm))
                h_max = sum(c * v for c, v in zip(pc, pM))
                return list(pM), h_min, h_max, h_max - h_min
            return [], 0, 0, 0
        prev = (list(c_vec), list(m_vec), list(M_vec))
        pc = prev[0]
        for i in range(1, len(c_vec)):
            d = _cdiv(c_vec[i], pc[i - 1])  # recover original delta
            d *= 2
            d = perturb(d)
            c_vec[i] = d * c_vec[i - 1]
        for i in range(len(M_vec) - 1):
            M_vec[i] = _cdiv(_cdiv(c_vec[i + 1], c_vec[i]) - 1, 2)
            m_vec[i] = -M_vec[i]
        M_vec[-1] = perturb(_cdiv(4 * M_vec[-1] + 1, 2))
        m_vec[-1] = -M_vec[-1]


@lru_cache(maxsize=None)
def kron_limits_table() -> Tuple[Tuple[Tuple[int, ...], int, int, int], ...]:
    """kronecker_array<int64_t>::get_limits() (kronecker_array.hpp:216-229): index = vector dimension."""
    out = [((), 0, 0, 0)]
    m = 1
    while True:
        M, hmin, hmax, diff = _determine_limit(m)
        if not M:
            break
        out.append((tuple(M), hmin, hmax, diff))
        m += 1
    return tuple(out)


def kron_limits(m: int) -> Tuple[Tuple[int, ...], int, int, int]:
    tab = kron_limits_table()
    if m >= len(tab):
        raise ValueError("size of vector to be encoded is too large")
    return tab[m]


def kron_encode(v: Sequence[int]) -> int:
    """kronecker_array.hpp:274-307."""
    m = len(v)
    if m == 0:
        return 0
    M, hmin, _, _ = kron_limits(m)
    for e, lim in zip(v, M):
        if e < -lim or e > lim:
            raise ValueError("a component of the vector to be encoded is out of bounds")
    ret = v[0] + M[0]
    cur = 2 * M[0] + 1
    for i in range(1, m):
        ret += (v[i] + M[i]) * cur
        cur *= 2 * M[i] + 1
    return ret + hmin


def kron_decode(n: int, m: int) -> Tuple[int, ...]:
    """kronecker_array.hpp:329-365."""
    if m == 0:
        if n != 0:
            raise ValueError("a vector of size 0 must always be encoded as 0")
        return ()
    M, hmin, hmax, _ = kron_limits(m)
    if n < hmin or n > hmax:
        raise ValueError("the integer to be decoded is out of bounds")
    code = n - hmin
    out = []
    mod = 1
    for i in range(m):
        r = 2 * M[i] + 1
        out.append((code % (mod * r)) // mod - M[i])
        mod *= r
    return tuple(out)


# --------------------------------------------------------------------------------------------------
# Exact polynomial arithmetic on {exponent tuple: int}
# --------------------------------------------------------------------------------------------------


def p_const(c: int, nvars: int) -> Poly:
    return {(0,) * nvars: c} if c else {}


def p_var(i: int, nvars: int, power: int = 1, cf: int = 1) -> Poly:
    e = [0] * nvars
    e[i] = power
    return {tuple(e): cf}


def p_add(a: Poly, b: Poly) -> Poly:
    r = dict(a)
    for k, c in b.items():
        s = r.get(k, 0) + c
        if s:
            r[k] = s
        else:
            r.pop(k, None)
    return r


def p_scale(a: Poly, c: int) -> Poly:
    return {k: v * c for k, v in a.items()} if c else {}


def p_degree(e: Sequence[int], mask: Sequence[int] | None = None) -> int:
    if mask is None:
        return sum(e)
    return sum(x for x, m in zip(e, mask) if m)


def p_mul(a: Poly, b: Poly, max_degree: int | None = None, mask: Sequence[int] | None = None) -> Poly:
    """Exact product; with ``max_degree`` a pair is kept iff deg(t1) + deg(t2) <= max_degree
    (polynomial.hpp:1495-1509: the first skipped term of the second series is the first with d2 > max - d1)."""
    r: Poly = {}
    bl = [(kb, cb, p_degree(kb, mask)) for kb, cb in b.items()]
    for ka, ca in a.items():
        da = p_degree(ka, mask)
        for kb, cb, db in bl:
            if max_degree is not None and da + db > max_degree:
                continue
            k = tuple(x + y for x, y in zip(ka, kb))
            r[k] = r.get(k, 0) + ca * cb
    return {k: v for k, v in r.items() if v != 0}


def p_pow(a: Poly, n: int) -> Poly:
    """Repeated multiplication, the way benchmarks/fateman1.hpp:46-48 builds its operands."""
    assert n >= 1
    r = dict(a)
    for _ in range(n - 1):
        r = p_mul(r, a)
    return r


def _compositions(total: int, parts: int):
    if parts == 1:
        yield (total,)
        return
    for i in range(total + 1):
        for rest in _compositions(total - i, parts - 1):
            yield (i,) + rest


def p_pow_multinomial(terms: Sequence[Tuple[Tuple[int, ...], int]], n: int) -> Poly:
    """(sum_k c_k x^{e_k})^n by the multinomial theorem -- closed form, independent of p_mul."""
    nvars = len(terms[0][0])
    r: Poly = {}
    fact = [math.factorial(i) for i in range(n + 1)]
    for comp in _compositions(n, len(terms)):
        cf = fact[n]
        for i in comp:
            cf //= fact[i]
        e = [0] * nvars
        for i, (ek, ck) in zip(comp, terms):
            if i:
                cf *= ck**i
                for v in range(nvars):
                    e[v] += ek[v] * i
        k = tuple(e)
        r[k] = r.get(k, 0) + cf
    return {k: v for k, v in r.items() if v != 0}


# --------------------------------------------------------------------------------------------------
# Benchmark / test operands.  Variables are ordered alphabetically by name, as symbol_fset does.
# --------------------------------------------------------------------------------------------------


def _linear(nvars: int, coeffs: Sequence[int], powers: Sequence[int] | None = None, const: int = 1):
    terms = []
    if const:
        terms.append(((0,) * nvars, const))
    for i, c in enumerate(coeffs):
        if c:
            e = [0] * nvars
            e[i] = 1 if powers is None else powers[i]
            terms.append((tuple(e), c))
    return terms


def fateman_operands(power: int = 20, nvars: int = 4) -> Tuple[Poly, Poly]:
    """benchmarks/fateman1.hpp:43-55 (power 20), fateman2.hpp (power 30): f, f + 1."""
    f = p_pow_multinomial(_linear(nvars, [1] * nvars), power)
    return f, p_add(f, p_const(1, nvars))


def fateman_cancel_operands(power: int = 10) -> Tuple[Poly, Poly]:
    """tests/polynomial_multiplier_01.cpp:218-231: (1+x+y+z+t)^p * (1-x+y+z+t)^p; order t,x,y,z."""
    f = p_pow_multinomial(_linear(4, [1, 1, 1, 1]), power)
    h = p_pow_multinomial(_linear(4, [1, -1, 1, 1]), power)
    return f, h


def pearce_operands(power: int = 12, minus_u: bool = False) -> Tuple[Poly, Poly]:
    """benchmarks/pearce1.hpp:40-61; variable order t,u,x,y,z.
    f = (1 + x + y + 2z^2 + 3t^3 + 5u^5)^p, g = (1 + u + t + 2z^2 + 3y^3 + 5x^5)^p.
    ``minus_u``: tests/base_series_multiplier.cpp:589-605 flips the sign of the linear u term in g."""
    t, u, x, y, z = range(5)

    def mono(**kw):
        e = [0] * 5
        for name, p in kw.items():
            e[{"t": t, "u": u, "x": x, "y": y, "z": z}[name]] = p
        return tuple(e)

    f_terms = [(mono(), 1), (mono(x=1), 1), (mono(y=1), 1), (mono(z=2), 2), (mono(t=3), 3), (mono(u=5), 5)]
    g_terms = [(mono(), 1), (mono(u=1), -1 if minus_u else 1), (mono(t=1), 1), (mono(z=2), 2), (mono(y=3), 3),
               (mono(x=5), 5)]
    return p_pow_multinomial(f_terms, power), p_pow_multinomial(g_terms, power)


def monagan_operands(which: int) -> Tuple[Poly, Poly]:
    """benchmarks/monagan.hpp:42-104."""
    if which == 1:
        f = p_add(p_pow_multinomial(_linear(3, [1, 1, 1]), 20), p_const(1, 3))
    elif which == 2:
        f = p_add(p_pow_multinomial(_linear(3, [1, 1, 1], powers=[2, 2, 2]), 20), p_const(1, 3))
    elif which == 3:
        f = p_add(p_pow_multinomial(_linear(3, [1, 1, 1]), 30), p_const(1, 3))
    elif which == 4:
        f = p_add(p_pow_multinomial(_linear(4, [1, 1, 1, 1]), 20), p_const(1, 4))
    elif which == 5:
        # symbols u, v, w, x, y (alphabetical)
        f = p_add(p_pow_multinomial(_linear(5, [1, 1, 1, 1, -1], powers=[2, 1, 2, 1, 1]), 10), p_const(1, 5))
        g = p_add(p_pow_multinomial(_linear(5, [1, 1, 1, 1, 1], powers=[1, 2, 1, 2, 1]), 10), p_const(1, 5))
        return f, g
    else:
        raise ValueError(which)
    return f, p_add(f, p_const(1, len(next(iter(f)))))


def fateman_product_closed_form(power: int, nvars: int = 4) -> Poly:
    """f*(f+1) = f^2 + f with f = (1 + x_1 + ... + x_v)^n: every coefficient is a sum of two multinomials."""
    sq = p_pow_multinomial(_linear(nvars, [1] * nvars), 2 * power)
    f = p_pow_multinomial(_linear(nvars, [1] * nvars), power)
    return p_add(sq, f)


# --------------------------------------------------------------------------------------------------
# SoA marshalling: {exponents: int}  <->  (int64 key[], uint64 mag[n, limbs], int8 sign[])
# --------------------------------------------------------------------------------------------------


def limbs_needed(p: Poly) -> int:
    mx = max((abs(c) for c in p.values()), default=0)
    return max(1, (mx.bit_length() + 63) // 64)


def poly_to_soa(p: Poly, nvars: int, limbs: int | None = None, shuffle_seed: int | None = 0):
    """Encode with the piranha Kronecker codification.  Term order is scrambled (hash_set order is unspecified)."""
    items = list(p.items())
    if shuffle_seed is not None and len(items) > 1:
        rs = np.random.RandomState(shuffle_seed)
        items = [items[i] for i in rs.permutation(len(items))]
    n = len(items)
    L = limbs or limbs_needed(p)
    key = np.zeros(n, dtype=np.int64)
    mag = np.zeros((n, L), dtype=np.uint64)
    sign = np.zeros(n, dtype=np.int8)
    mask = (1 << 64) - 1
    for i, (e, c) in enumerate(items):
        assert len(e) == nvars
        key[i] = kron_encode(e)
        a = abs(c)
        assert a.bit_length() <= 64 * L
        for l in range(L):
            mag[i, l] = (a >> (64 * l)) & mask
        sign[i] = 1 if c > 0 else (-1 if c < 0 else 0)
    return key, mag, sign


def soa_to_poly(key, mag, sign, nvars: int) -> Poly:
    r: Poly = {}
    mag = np.asarray(mag, dtype=np.uint64).reshape(len(key), -1)
    for i in range(len(key)):
        a = 0
        for l in range(mag.shape[1]):
            a |= int(mag[i, l]) << (64 * l)
        e = kron_decode(int(key[i]), nvars)
        assert e not in r, "duplicate key in result"
        r[e] = a if sign[i] >= 0 else -a
    return r


def soa_coefficients(mag, sign) -> List[int]:
    mag = np.asarray(mag, dtype=np.uint64)
    n = mag.shape[0]
    out = [0] * n
    cols = [mag[:, l].tolist() for l in range(mag.shape[1])]
    sg = np.asarray(sign).tolist()
    for i in range(n):
        a = 0
        for l, col in enumerate(cols):
            a |= col[i] << (64 * l)
        out[i] = a if sg[i] >= 0 else -a
    return out


def canonical(p: Poly, nvars: int):
    """Canonical (sorted by Kronecker code) list of (code, coefficient)."""
    return sorted((kron_encode(e), c) for e, c in p.items())


def check_bounds(a: Poly, b: Poly, nvars: int) -> None:
    """polynomial.hpp:1118-1194: per-variable min/max of both operands, summed, must stay inside +-limits."""
    if not a or not b or nvars == 0:
        return
    M = kron_limits(nvars)[0]
    for v in range(nvars):
        lo = min(e[v] for e in a) + min(e[v] for e in b)
        hi = max(e[v] for e in a) + max(e[v] for e in b)
        if lo < -M[v] or hi > M[v]:
            raise OverflowError("Kronecker monomial components are out of bounds")
