// Fateman's test number 1 with rational coefficients: f = (1+x+y+z+t)^20 over Q, time f * (f + 1), expect 135751 terms
// (reference: benchmarks/fateman1_rational.cpp:48-56).  The rational multiplier brings both operands to a common
// denominator, runs the integer product on the GPU and divides back (piranha_b200/host/rational.hpp).
// Usage: fateman1_rational [repetitions]
#include <cstdlib>

#include "../piranha_b200/host/rational.hpp"
#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    using qt = pb200::polynomial<pb200::rational, pb200::kronecker_monomial<std::int64_t>>;
    qt x{"x"}, y{"y"}, z{"z"}, t{"t"};
    qt f = x + y + z + t + 1;
    const qt tmp(f);
    for (int i = 1; i < 20; ++i) f *= tmp;
    const qt g = f + 1;
    const int reps = argc > 1 ? std::atoi(argv[1]) : 3;
    std::size_t size = 0;
    double best = 1e300;
    for (int r = 0; r < reps; ++r) {
        simple_timer timer("fateman1_rational");
        const qt res = f * g;
        size = res.size();
        best = std::min(best, timer.elapsed_ms());
    }
    const pk_stats st = pb200::series_multiplier<pt>::last_stats();
    std::printf("fateman1_rational: %zu x %zu terms -> %zu terms; operator* %.3f ms including the lcm pass, the GPU product "
                "(device %.3f ms) and the canonicalisation of every result coefficient on the host\n",
                f.size(), g.size(), size, best, st.device_ms);
    if (size != 135751u) {
        std::printf("FAILED: expected 135751 terms\n");
        return 1;
    }
    return 0;
}
