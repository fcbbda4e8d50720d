"""The C++ host mirror of the reference's operator interface (piranha_b200/host/): polynomial, kronecker_monomial,
hash_set, integer and the series_multiplier specialisation that calls the C ABI.

  host_logic_test   CPU only: arithmetic, codification, container, series addition (no product is computed)
  multiplier_test   GPU: the reference's own multiplier tests restated against operator* (tests/cpp/multiplier_test.cpp)
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "bin")


def _build():
    subprocess.check_call(["make", "-s", "-C", ROOT, "host"], stdout=subprocess.DEVNULL)


def _run(name, timeout):
    _build()
    p = subprocess.run([os.path.join(BIN, name)], capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    assert " 0 failures" in p.stdout
    return p.stdout


def test_host_logic():
    out = _run("host_logic_test", 300)
    for t in ("integer_test", "kronecker_array_test", "kronecker_monomial_test", "hash_set_test", "series_test"):
        assert f"[  OK  ] {t}" in out


def test_host_kronecker_limits_match_the_golden_table():
    """The C++ kronecker_array the library codifies with (piranha_b200/host/kronecker_array.hpp: the perturbed-prime
    search of the reference's kronecker_array.hpp:128-229) yields the table recorded in tests/golden/."""
    import json

    _build()
    out = subprocess.check_output([os.path.join(BIN, "host_logic_test"), "--limits"], text=True)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "kronecker_limits.json")))
    rows = [list(map(int, line.split())) for line in out.splitlines()]
    assert len(rows) == len(golden["limits"])
    for m, hmin, *M in rows:
        assert M == golden["limits"][m], m
        assert hmin == golden["h_min"][m], m


def test_host_integer_and_rational_against_python():
    """The mirror's exact arithmetic (integer +, -, *, truncating / and %, gcd; rational +, -, *, / in canonical form) on
    random operands of up to ~180 bits against Python's int and fractions.Fraction (C semantics for / and %)."""
    import random
    from fractions import Fraction

    _build()
    rng = random.Random(20240607)

    def big(bits, nonzero=False):
        while True:
            v = rng.getrandbits(rng.randrange(1, bits)) * rng.choice((-1, 1))
            if v or not nonzero:
                return v

    lines, want = [], []
    for _ in range(1500):
        op = rng.choice(("add", "sub", "mul", "div", "mod", "gcd"))
        a, b = big(400), big(400, nonzero=op in ("div", "mod"))
        if op == "div":
            r = abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
        elif op == "mod":
            r = abs(a) % abs(b) * (-1 if a < 0 else 1)
        elif op == "gcd":
            import math
            r = math.gcd(a, b)
        else:
            r = {"add": a + b, "sub": a - b, "mul": a * b}[op]
        lines.append(f"{op} {a} {b}")
        want.append(str(r))
    for _ in range(1500):
        op = rng.choice(("qadd", "qsub", "qmul", "qdiv"))
        an, ad, bn, bd = big(90), big(90, True), big(90, nonzero=op == "qdiv"), big(90, True)
        fa, fb = Fraction(an, ad), Fraction(bn, bd)
        r = {"qadd": fa + fb, "qsub": fa - fb, "qmul": fa * fb, "qdiv": (fa / fb) if bn else None}[op]
        lines.append(f"{op} {an} {ad} {bn} {bd}")
        want.append(f"{r.numerator} {r.denominator}")
    lines.append("mul " + str((1 << 200) + 12345) + " " + str(-(1 << 200) + 1))  # promoted beyond the static storage
    want.append(str(((1 << 200) + 12345) * (-(1 << 200) + 1)))
    lines += ["div 5 0", "qdiv 1 2 0 3"]
    p = subprocess.run([os.path.join(BIN, "host_logic_test"), "--calc"], input="\n".join(lines) + "\n", capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    got = p.stdout.splitlines()
    assert len(got) == len(lines)
    bad = [(l, g, w) for l, g, w in zip(lines, got, want) if g != w]
    assert not bad, bad[:5]
    assert all(g.startswith("error") for g in got[len(want):])  # division by zero: reported


@pytest.mark.gpu
def test_reference_multiplier_tests_through_operator_star():
    out = _run("multiplier_test", 900)
    for t in ("residency_test", "rational_test", "unpacked_test", "bounds_test", "multiplication_test", "pearce_test", "truncation_test", "threads_test",
              "shared_operands_test", "multi_gpu_test"):
        assert f"[  OK  ] {t}" in out


@pytest.mark.gpu
@pytest.mark.parametrize("prog,size", [("fateman1", 135751), ("pearce1", 5821335), ("fateman2", 635376),
                                       ("fateman1_rational", 135751), ("pearce1_rational", 5821335), ("fateman1_dynamic", 135751)])
def test_cpp_benchmarks(prog, size):
    """benchmarks/*.cpp: operands built by repeated operator*=, the final product timed, the result size checked
    against the reference's known answer (benchmarks/fateman1.cpp:55, pearce1.cpp:56, fateman2.cpp:55,
    fateman1_rational.cpp:55, pearce1_rational.cpp:56)."""
    _build()
    p = subprocess.run([os.path.join(BIN, prog), "2"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert f"-> {size} terms" in p.stdout
