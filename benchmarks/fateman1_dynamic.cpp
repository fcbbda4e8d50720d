// Fateman's test number 1 with operands scaled by the largest limb value (reference: benchmarks/fateman1_dynamic.cpp:50-60,
// fateman1.hpp:49-51: f *= 2^64 - 1 before the timed product), so that operand coefficients leave one limb and the
// product's leave two: the width protocol of the C ABI (two-limb operands on the pairing kernels).  Expect 135751 terms;
// then the product is multiplied on (three-limb operand: 128-bit halves), which the reference does by promoting.
// Usage: fateman1_dynamic [repetitions]
#include <cstdlib>

#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    pt g;
    pt f = fateman(20, &g);
    f = f * 18446744073709551615ull;
    g = f + 1;
    const int rc = run("fateman1_dynamic", f, g, 135751u, argc > 1 ? std::atoi(argv[1]) : 5);
    if (rc) return rc;
    // chained: (f * g) * f, operands of three limbs and two limbs
    const pt p = f * g;
    simple_timer timer("fateman1_dynamic chained (f * g) * f");
    const pt q = p * f;
    std::printf("chained product: %zu terms\n", q.size());
    return q.size() == 635376u ? 0 : 1; // every monomial of degree <= 60 in four variables: C(64, 4)
}
