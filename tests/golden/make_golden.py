#!/usr/bin/env python
"""Generates the committed fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference cannot be built or imported in this image (no Boost / mp++ / gmp.h, pyranha not installed), so the
fixtures come from implementations that are independent of the oracles under test:
  * products.json         -- small products expanded with SymPy (exact rational arithmetic, its own multiplication
                             algorithm), including the literal expectations written in the reference's
                             tests/polynomial_truncation.cpp:75-170;
  * kronecker_limits.json -- the Kronecker limits table printed by a C++ program that drives the real libstdc++
                             std::mt19937 / std::uniform_int_distribution (piranha_b200/host/kronecker_array.hpp),
                             i.e. what kronecker_array.hpp:128-229 produces with this image's standard library.
"""
import json
import os
import subprocess
import tempfile

import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def poly_terms(expr, syms):
    """[(exponents, coefficient)] of a (Laurent) polynomial expression."""
    expr = sp.expand(expr)
    out = []
    for term in sp.Add.make_args(expr):
        c, rest = term.as_coeff_Mul()
        powers = rest.as_powers_dict()
        e = [int(powers.get(s, 0)) for s in syms]
        assert sp.simplify(rest - sp.Mul(*[s**k for s, k in zip(syms, e)])) == 0
        if c != 0:
            out.append((e, str(int(c))))
    return sorted(out)


def truncate(terms_a, terms_b, syms, D, mask):
    """Pairwise truncated product: keep (i, j) iff deg_i + deg_j <= D (polynomial.hpp:1495-1509)."""
    acc = {}
    for ea, ca in terms_a:
        for eb, cb in terms_b:
            da = sum(x for x, m in zip(ea, mask) if m)
            db = sum(x for x, m in zip(eb, mask) if m)
            if da + db > D:
                continue
            k = tuple(x + y for x, y in zip(ea, eb))
            acc[k] = acc.get(k, 0) + int(ca) * int(cb)
    return sorted((list(k), str(v)) for k, v in acc.items() if v)


def main():
    t, u, x, y, z = sp.symbols("t u x y z")
    cases = []

    def add_case(name, syms, a, b, D=None, mask=None):
        ta, tb = poly_terms(a, syms), poly_terms(b, syms)
        if D is None:
            prod = poly_terms(sp.expand(a * b), syms)
        else:
            prod = truncate(ta, tb, syms, D, mask or [1] * len(syms))
        cases.append({"name": name, "n_vars": len(syms), "a": ta, "b": tb, "product": prod, "max_degree": D,
                      "mask": mask})

    s4 = [t, x, y, z]
    f = (1 + x + y + z + t) ** 4
    add_case("fateman4", s4, f, f + 1)
    add_case("fateman3_cancel", s4, (1 + x + y + z + t) ** 3, (1 - x + y + z + t) ** 3)
    s5 = [t, u, x, y, z]
    pf = (1 + x + y + 2 * z**2 + 3 * t**3 + 5 * u**5) ** 2
    pg = (1 + u + t + 2 * z**2 + 3 * y**3 + 5 * x**5) ** 2
    add_case("pearce2", s5, pf, pg)
    add_case("pearce2_minus_u", s5, pf, (1 - u + t + 2 * z**2 + 3 * y**3 + 5 * x**5) ** 2)
    add_case("laurent", [x, y], (x**-3 + 2 * y**2 - 7 * x * y**-1) ** 3, (x**2 - y**-2 + 3) ** 3)
    add_case("difference_of_squares", [x, y], x + y, x - y)  # tests/polynomial_multiplier_03.cpp:75-127 family
    big = 2**70 + 12345
    add_case("wide_coefficients", [x, y], (big * x + (big + 7) * y - big) * (x - 3 * y**2 + 1),
             ((2**100 + 3) * x - y + 3) * (x - 2 * y + 5))  # operands of 73 and 103 bits, products of 176
    add_case("monagan5_lite", [u, sp.Symbol("v"), sp.Symbol("w"), x, y],
             (1 + u**2 + sp.Symbol("v") + sp.Symbol("w") ** 2 + x - y) ** 3 + 1,
             (1 + u + sp.Symbol("v") ** 2 + sp.Symbol("w") + x**2 + y) ** 3 + 1)
    # literal truncation expectations of tests/polynomial_truncation.cpp:79-107 and :141-170
    g = x + y + 1
    for D in (1, 2, 3, 0, -1):
        add_case(f"trunc_total_{D}", s4, g, g, D=D)
    add_case("trunc_total_linear", s4, x + y + z + t, x + y + z + t, D=1)
    add_case("trunc_partial_x", s4, g, g, D=1, mask=[0, 1, 0, 0])
    add_case("trunc_partial_z", s4, g, g, D=1, mask=[0, 0, 0, 1])
    add_case("trunc_partial_xy", s4, g, g, D=1, mask=[0, 1, 1, 0])
    add_case("trunc_partial_x_linear", s4, x + y + z + t, x + y + z + t, D=1, mask=[0, 1, 0, 0])
    add_case("trunc_fateman4_D5", s4, f, f + 1, D=5)
    with open(os.path.join(HERE, "products.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py (SymPy %s)" % sp.__version__, "cases": cases}, fh)

    src = r'''
#include "%s/piranha_b200/host/kronecker_array.hpp"
#include <cstdio>
int main() {
    const auto &l = pb200::kronecker_array<std::int64_t>::get_limits();
    std::printf("{\"limits\": [");
    for (std::size_t m = 0; m < l.size(); ++m) {
        std::printf("%%s[", m ? ", " : "");
        const auto &v = std::get<0>(l[m]);
        for (std::size_t i = 0; i < v.size(); ++i) std::printf("%%s%%lld", i ? ", " : "", (long long)v[i]);
        std::printf("]");
    }
    std::printf("], \"h_min\": [");
    for (std::size_t m = 0; m < l.size(); ++m) std::printf("%%s%%lld", m ? ", " : "", (long long)std::get<1>(l[m]));
    std::printf("]}\n");
}
''' % ROOT
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "lim.cpp")
        with open(cpp, "w") as fh:
            fh.write(src)
        exe = os.path.join(td, "lim")
        subprocess.check_call(["g++", "-O2", "-std=c++17", cpp, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    tab = json.loads(out)
    tab["generator"] = "piranha_b200/host/kronecker_array.hpp compiled with g++ (libstdc++ std::mt19937 / uniform_int_distribution)"
    with open(os.path.join(HERE, "kronecker_limits.json"), "w") as fh:
        json.dump(tab, fh)
    print(len(cases), "product cases;", len(tab["limits"]), "limit rows")


if __name__ == "__main__":
    main()
