// Shared helpers of the C++ benchmark programs (the counterparts of the reference's benchmarks/fateman1.hpp,
// pearce1.hpp and simple_timer.hpp, written against the host mirror in piranha_b200/host/).
#ifndef PB200_BENCH_COMMON_HPP
#define PB200_BENCH_COMMON_HPP

#include <chrono>
#include <cstdio>
#include <string>

#include "../piranha_b200/host/polynomial.hpp"

namespace pb200_bench
{

using pt = pb200::polynomial<pb200::integer, pb200::kronecker_monomial<std::int64_t>>;

// RAII wall-clock timer printing "Elapsed time: N ms", like benchmarks/simple_timer.hpp:40-57
struct simple_timer {
    std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
    const char *what;
    explicit simple_timer(const char *w = "") : what(w) {}
    double elapsed_ms() const
    {
        return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    }
    ~simple_timer() { std::printf("Elapsed time%s%s: %.3f ms\n", *what ? " " : "", what, elapsed_ms()); }
};

// f = (1 + x + y + z + t)^n built by repeated f *= tmp (benchmarks/fateman1.hpp:43-48); the timed product is f * (f + 1)
inline pt fateman(int n, pt *g_out)
{
    pt x{"x"}, y{"y"}, z{"z"}, t{"t"};
    pt f = x + y + z + t + 1;
    const pt tmp(f);
    for (int i = 1; i < n; ++i) f *= tmp;
    *g_out = f + 1;
    return f;
}

// f = (1+x+y+2z^2+3t^3+5u^5)^n, g = (1+u+t+2z^2+3y^3+5x^5)^n (benchmarks/pearce1.hpp:40-61)
inline pt pearce(int n, pt *g_out)
{
    pt x{"x"}, y{"y"}, z{"z"}, t{"t"}, u{"u"};
    pt f = x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1;
    pt g = u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1;
    const pt tmp_f(f), tmp_g(g);
    for (int i = 1; i < n; ++i) {
        f *= tmp_f;
        g *= tmp_g;
    }
    *g_out = g;
    return f;
}

// time `reps` products a * b through operator*.  The product is left where it is computed (HBM); reading its terms
// brings it into the host hash_set once.  Both times are reported: the reference's benchmark times operator* alone, and
// its result is a host container, so the comparable figure is the sum.
inline int run(const char *name, const pt &a, const pt &b, std::size_t expected_size, int reps)
{
    std::size_t size = 0;
    double best_mul = 1e300, best_host = 1e300;
    for (int r = 0; r < reps; ++r) {
        simple_timer timer(name);
        const pt res = a * b;
        size = res.size();
        const double t_mul = timer.elapsed_ms();
        res.materialise();
        const double t_all = timer.elapsed_ms();
        best_mul = std::min(best_mul, t_mul);
        best_host = std::min(best_host, t_all - t_mul);
    }
    const pk_stats st = pb200::series_multiplier<pt>::last_stats();
    const double W = static_cast<double>(a.size()) * static_cast<double>(b.size());
    std::printf("%s: %zu x %zu terms -> %zu terms; operator* %.3f ms with the result resident in HBM (device %.3f ms, pairing "
                "kernel %.3f ms), + %.3f ms to fill the host hash_set; %.1f / %.2f G term-products/s\n",
                name, a.size(), b.size(), size, best_mul, st.device_ms, st.mul_kernel_ms, best_host, W / best_mul / 1e6,
                W / (best_mul + best_host) / 1e6);
    if (size != expected_size) {
        std::printf("FAILED: expected %zu terms\n", expected_size);
        return 1;
    }
    return 0;
}

} // namespace pb200_bench

#endif
