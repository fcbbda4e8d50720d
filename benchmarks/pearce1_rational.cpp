// Pearce's test number 1 with rational coefficients: f = (1+x+y+2z^2+3t^3+5u^5)^12, g = (1+u+t+2z^2+3y^3+5x^5)^12 over Q,
// time f * g, expect 5821335 terms (reference: benchmarks/pearce1_rational.cpp:48-57, benchmarks/pearce1.hpp:40-61).
// Common-denominator pass, integer product on the GPU, rational result container filled by the host threads
// (piranha_b200/host/rational.hpp).  Usage: pearce1_rational [repetitions]
#include <cstdlib>

#include "../piranha_b200/host/rational.hpp"
#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    using qt = pb200::polynomial<pb200::rational, pb200::kronecker_monomial<std::int64_t>>;
    qt x{"x"}, y{"y"}, z{"z"}, t{"t"}, u{"u"};
    qt f = x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1;
    qt g = u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1;
    const qt tmp_f(f), tmp_g(g);
    for (int i = 1; i < 12; ++i) {
        f *= tmp_f;
        g *= tmp_g;
    }
    const int reps = argc > 1 ? std::atoi(argv[1]) : 3;
    std::size_t size = 0;
    double best = 1e300;
    for (int r = 0; r < reps; ++r) {
        simple_timer timer("pearce1_rational");
        const qt res = f * g;
        size = res.size();
        best = std::min(best, timer.elapsed_ms());
    }
    const pk_stats st = pb200::series_multiplier<pt>::last_stats();
    std::printf("pearce1_rational: %zu x %zu terms -> %zu terms; operator* %.3f ms including the lcm pass, the GPU product "
                "(device %.3f ms) and the fill of the rational result container on %u host threads\n",
                f.size(), g.size(), size, best, st.device_ms, pb200::settings::get_n_threads());
    if (size != 5821335u) {
        std::printf("FAILED: expected 5821335 terms\n");
        return 1;
    }
    return 0;
}
