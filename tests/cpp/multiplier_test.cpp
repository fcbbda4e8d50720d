// GPU tests of the drop-in through the reference's own operator interface: every product below goes
// polynomial::operator* -> series_multiplier<polynomial<integer, kronecker_monomial<>>> -> pk_mul (CUDA).
// The cases restate the reference's tests for this path:
//   tests/polynomial_multiplier_01.cpp:120-174  bounds (std::overflow_error from the constructor; one below the limit is fine)
//   tests/polynomial_multiplier_01.cpp:187-238  empty series, reduced Fateman 10626, cancellations 5786
//   tests/base_series_multiplier.cpp:517-605    reduced Pearce 591235 / 591184
//   tests/polynomial_truncation.cpp:69-177      total and partial degree auto-truncation
//   tests/polynomial_multiplier_03.cpp:75-127   untruncated_multiplication / truncated_multiplication statics
//   tests/polynomial_01.cpp:321-422             differential test against an independent multiplier (here: a plain
//                                               host double loop used only as the checker inside this test)
#include <cstdint>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../piranha_b200/host/monomial.hpp"
#include "../../piranha_b200/host/polynomial.hpp"
#include "../../piranha_b200/host/rational.hpp"
#include "check.hpp"

using namespace pb200;
using k_type = kronecker_monomial<std::int64_t>;
using pt = polynomial<integer, k_type>;

// checker only: schoolbook product on the host through series::insert (the role of polynomial_alt's plain multiplier)
static pt plain_product(const pt &a, const pt &b)
{
    symbol_fset ss(a.get_symbol_set());
    ss.insert(b.get_symbol_set().begin(), b.get_symbol_set().end());
    const pt a2 = a.extend_symbol_set(ss), b2 = b.extend_symbol_set(ss);
    pt r;
    r.set_symbol_set(ss);
    for (const auto &t1 : a2._container())
        for (const auto &t2 : b2._container()) {
            pt::term_type t;
            k_type::multiply(t, t1, t2, ss);
            r.insert(t);
        }
    return r;
}

// (x + y + 1)^3 * (x + 1) truncated to partial degree 3 in x: the untruncated product with the terms of higher x-degree dropped
template <typename S>
static S pt_partial_reference()
{
    S x{"x"}, y{"y"};
    S::unset_auto_truncate_degree();
    const S full = (x + y + 1).pow(3) * (x + 1);
    S r;
    r.set_symbol_set(full.get_symbol_set());
    for (const auto &t : full._container())
        if (t.m_key.unpack(full.get_symbol_set())[0] <= 3) r.insert(t);
    S::set_auto_truncate_degree(3, {"x"});
    return r;
}

static void bounds_test()
{
    using ka = kronecker_array<std::int64_t>;
    const auto &limits = std::get<0>(ka::get_limits()[3]);
    pt x{"x"}, y{"y"}, z{"z"};
    CHECK_THROW(x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]) * x, std::overflow_error);
    CHECK_THROW((x.pow(limits[0]) + 1) * (y.pow(limits[1]) + 1) * (z.pow(limits[2]) + 1) * (x + 1), std::overflow_error);
    CHECK_THROW(x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]) * y, std::overflow_error);
    CHECK_THROW(x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]) * z, std::overflow_error);
    CHECK_THROW(x.pow(-limits[0]) * y.pow(limits[1]) * z.pow(limits[2]) * x.pow(-1), std::overflow_error);
    CHECK_THROW((x.pow(limits[0]) + 1) * (y.pow(-limits[1]) + 1) * (z.pow(limits[2]) + 1) * (y.pow(-1) + 1), std::overflow_error);
    CHECK_THROW(x.pow(limits[0]) * y.pow(limits[1]) * z.pow(-limits[2]) * z.pow(-1), std::overflow_error);
    CHECK_EQUAL(x.pow(limits[0] - 1) * y.pow(limits[1]) * z.pow(limits[2]) * x, x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]));
    CHECK_EQUAL(x.pow(limits[0]) * y.pow(limits[1] - 1) * z.pow(limits[2]) * y, x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]));
    CHECK_EQUAL(x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2] - 1) * z, x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]));
    CHECK_EQUAL(x.pow(-limits[0] + 1) * y.pow(-limits[1]) * z.pow(-limits[2]) * x.pow(-1),
                x.pow(-limits[0]) * y.pow(-limits[1]) * z.pow(-limits[2]));
    CHECK_EQUAL(pt{2} * pt{3}, 6);
    // the exception comes from the constructor of the multiplier, before anything is computed
    const pt big = x.pow(limits[0]) * y.pow(limits[1]) * z.pow(limits[2]);
    const pt xx = x.extend_symbol_set(big.get_symbol_set());
    CHECK_THROW((series_multiplier<pt>(big, xx)), std::overflow_error);
    // mismatching symbol sets reach the multiplier only through direct use
    CHECK_THROW((series_multiplier<pt>(x, y)), std::invalid_argument);
}

static void multiplication_test()
{
    pt e1, e2;
    CHECK_EQUAL(e1 * e2, 0);
    CHECK_EQUAL((e1 * e2).get_symbol_set().size(), 0u);
    pt x{"x"};
    CHECK_EQUAL(e1 * x, 0);
    CHECK_EQUAL(x * e1, 0);
    CHECK(((x * e1).get_symbol_set() == symbol_fset{"x"}));
    CHECK(((e1 * x).get_symbol_set() == symbol_fset{"x"}));
    // a reduced Fateman benchmark
    pt y{"y"}, z{"z"}, t{"t"};
    auto f = 1 + x + y + z + t;
    auto tmp2 = f;
    for (int i = 1; i < 10; ++i) f *= tmp2;
    auto g = f + 1;
    auto retval = f * g;
    CHECK_EQUAL(retval.size(), 10626u);
    // every coefficient of f*(f+1) has the closed form multinomial(20; a,b,c,d,.) + [deg <= 10] multinomial(10; ...)
    CHECK_EQUAL(retval.find_cf({0, 0, 0, 0}), integer(2));
    CHECK_EQUAL(retval.find_cf({20, 0, 0, 0}), integer(1));
    CHECK_EQUAL(retval.find_cf({1, 0, 0, 0}), integer(30));
    CHECK_EQUAL(retval.find_cf({5, 5, 5, 5}), integer("11732745024"));
    // with cancellations
    auto h = 1 - x + y + z + t;
    f = 1 + x + y + z + t;
    tmp2 = h;
    auto tmp3 = f;
    for (int i = 1; i < 10; ++i) {
        h *= tmp2;
        f *= tmp3;
    }
    retval = f * h;
    CHECK_EQUAL(retval.size(), 5786u);
    // differential check on a smaller pair against the host schoolbook product
    auto f5 = (1 + x + y + z + t).pow(5), h5 = (1 - x + y + z + t).pow(5);
    CHECK(f5 * h5 == plain_product(f5, h5));
    CHECK(f5 * h5 == h5 * f5);
}

static void pearce_test()
{
    pt x{"x"}, y{"y"}, z{"z"}, t{"t"}, u{"u"};
    auto f = (x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1);
    auto g = (u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1);
    const auto tmp_f(f), tmp_g(g);
    for (int i = 1; i < 8; ++i) {
        f *= tmp_f;
        g *= tmp_g;
    }
    CHECK_EQUAL((f * g).size(), 591235u);
    auto f2 = (x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1);
    auto g2 = (-1 * u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1);
    const auto tmp_f2(f2), tmp_g2(g2);
    for (int i = 1; i < 8; ++i) {
        f2 *= tmp_f2;
        g2 *= tmp_g2;
    }
    CHECK_EQUAL((f2 * g2).size(), 591184u);
}

static void truncation_test()
{
    pt x{"x"}, y{"y"}, z{"z"}, t{"t"};
    auto tup = pt::get_auto_truncate_degree();
    CHECK_EQUAL(std::get<0>(tup), 0);
    CHECK_EQUAL((x + y + z + t) * (x + y + z + t),
                t * t + 2 * t * x + 2 * t * y + 2 * t * z + x * x + 2 * x * y + 2 * x * z + y * y + 2 * y * z + z * z);
    // total degree
    pt::set_auto_truncate_degree(1);
    CHECK_EQUAL((x + y + 1) * (x + y + 1), 2 * x + 2 * y + 1);
    pt::set_auto_truncate_degree(2);
    pt::unset_auto_truncate_degree(); // build the expected value untruncated
    const pt full = 2 * x + 2 * y + 1 + 2 * x * y + x * x + y * y;
    pt::set_auto_truncate_degree(2);
    CHECK_EQUAL((x + y + 1) * (x + y + 1), full);
    pt::set_auto_truncate_degree(3);
    CHECK_EQUAL((x + y + 1) * (x + y + 1), full);
    pt::set_auto_truncate_degree(0);
    CHECK_EQUAL((x + y + 1) * (x + y + 1), 1);
    pt::set_auto_truncate_degree(-1);
    CHECK_EQUAL((x + y + 1) * (x + y + 1), 0);
    pt::set_auto_truncate_degree(1);
    CHECK_EQUAL((x + y + z + t) * (x + y + z + t), 0);
    // partial degree
    pt::unset_auto_truncate_degree();
    const pt e_x = 2 * x * y + 2 * x + y * y + 2 * y + 1, e_z = x * x + 2 * x * y + 2 * x + y * y + 2 * y + 1;
    const pt e_4 = t * t + 2 * t * x + 2 * t * y + 2 * t * z + 2 * x * y + 2 * x * z + y * y + 2 * y * z + z * z;
    pt::set_auto_truncate_degree(1, {"x"});
    CHECK_EQUAL((x + y + 1) * (x + y + 1), e_x);
    pt::set_auto_truncate_degree(1, {"z"});
    CHECK_EQUAL((x + y + 1) * (x + y + 1), e_z);
    pt::set_auto_truncate_degree(1, {"x", "y"});
    CHECK_EQUAL((x + y + 1) * (x + y + 1), 2 * x + 2 * y + 1);
    pt::set_auto_truncate_degree(1, {"x"});
    CHECK_EQUAL((x + y + z + t) * (x + y + z + t), e_4);
    pt::unset_auto_truncate_degree();
    // the explicit static forms (polynomial_multiplier_03.cpp:75-127)
    CHECK_EQUAL(pt::untruncated_multiplication(x + y, x - y), x * x - y * y);
    CHECK_EQUAL(pt::truncated_multiplication(x + y + 1, x + y + 1, 1), 2 * x + 2 * y + 1);
    CHECK_EQUAL(pt::truncated_multiplication(x + y + 1, x + y + 1, 1, {"x"}), e_x);
    CHECK_EQUAL(pt::truncated_multiplication(x + y + 1, x + y + 1, -1), 0);
    // degree-limit arithmetic is checked: max - d1 overflowing throws (polynomial.hpp:1243-1252)
    CHECK_THROW(pt::truncated_multiplication(1 + x.pow(-1), x, std::numeric_limits<std::int64_t>::max()), std::overflow_error);
    // truncated Fateman: degree 20 keeps exactly the terms of f (benchmarks/fateman1_unpacked_truncation.cpp:56-59, power 10 here)
    auto f = (1 + x + y + z + t).pow(10);
    CHECK_EQUAL(pt::truncated_multiplication(f, f + 1, 10).size(), 1001u);
    CHECK_EQUAL(pt::truncated_multiplication(f, f + 1, 20).size(), 10626u);
}

static void threads_test()
{
    // operator* may be called concurrently from different user threads on different operands
    std::vector<std::size_t> sizes(4, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < 4; ++k)
        th.emplace_back([k, &sizes]() {
            pt x{"x"}, y{"y"}, z{"z"};
            auto f = (1 + x + y + z).pow(6 + k);
            sizes[k] = (f * (f + 1)).size();
        });
    for (auto &t : th) t.join();
    for (int k = 0; k < 4; ++k) {
        const std::size_t n = 2 * (6 + k); // (1+x+y+z)^(2n'): C(n+3, 3) terms
        CHECK_EQUAL(sizes[k], (n + 1) * (n + 2) * (n + 3) / 6);
    }
}

static void shared_operands_test()
{
    // concurrent const use of the SAME operands from several threads (the reference allows it): the lazy device upload
    // and the lazy host download of a series are serialised per series, every thread sees the same product
    pt x{"x"}, y{"y"}, z{"z"};
    const pt f = (1 + x + y + z).pow(9), g = (1 + x - y + z).pow(9);
    const pt shared_product = f * g; // device-resident; the threads below read (download) it concurrently
    std::vector<std::size_t> sizes(6, 0), reads(6, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < 6; ++k)
        th.emplace_back([k, &sizes, &reads, &f, &g, &shared_product]() {
            sizes[k] = (f * g).size();
            reads[k] = shared_product._container().size();
        });
    for (auto &t : th) t.join();
    const pt want = plain_product(f, g);
    for (int k = 0; k < 6; ++k) {
        CHECK_EQUAL(sizes[k], want.size());
        CHECK_EQUAL(reads[k], want.size());
    }
    CHECK_EQUAL(shared_product, want);
    // a series created and destroyed on another thread frees its device terms through the context that owns them
    pt moved;
    std::thread([&moved, &f]() { moved = f * f; }).join();
    CHECK_EQUAL(moved.size(), (18u + 1) * (18u + 2) * (18u + 3) / 6);
}

static void multi_gpu_test()
{
    // settings::set_n_gpus(n): one product spread over n devices behind the same operator* (pk_ctx_create_multi); with
    // fewer physical GPUs the parts share a device (PB200_DEVICES names them), the code path is the same
    pt x{"x"}, y{"y"}, z{"z"}, t{"t"};
    const pt f = (1 + x + y + z + t).pow(10), g = (1 - x + y + z + t).pow(10);
    const pt one = f * g;
    CHECK_EQUAL(one.size(), 5786u); // polynomial_multiplier_01.cpp:218-231
    one.materialise();
    settings::set_n_gpus(3);
    const pt three = f * g; // operands re-uploaded to the new context
    CHECK_EQUAL(three.size(), 5786u);
    CHECK_EQUAL(three, one);
    const pt chained = three * f; // a partitioned product as an operand: replicated device-to-device
    settings::reset_n_gpus();
    CHECK_EQUAL(chained, one * f);
    CHECK_THROW(settings::set_n_gpus(0), std::invalid_argument);
}

static void residency_test()
{
    // a product stays on the device until its terms are read; operands keep the device copy of their first product, so
    // chained products (f *= tmp, pow; series.hpp:799-805, :2333-2403) neither re-upload nor download in between
    pt x{"x"}, y{"y"}, z{"z"};
    const pt base = 1 + x + y + z;
    CHECK(base.is_host_resident());
    CHECK(!base.is_device_resident());
    pt f = base * base;
    CHECK(f.is_device_resident());
    CHECK(!f.is_host_resident());
    CHECK(base.is_device_resident()); // uploaded once, kept
    CHECK_EQUAL(f.size(), 10u);       // answered from the device
    CHECK(!f.is_host_resident());
    pt g = f;
    for (int i = 0; i < 6; ++i) g *= base; // (1+x+y+z)^8 without touching the host
    CHECK(!g.is_host_resident());
    CHECK_EQUAL(g.size(), 165u);
    CHECK_EQUAL(g.find_cf({0, 0, 0}), integer(1)); // first read: materialised
    CHECK(g.is_host_resident());
    CHECK(g.is_device_resident());
    CHECK_EQUAL(g.find_cf({2, 3, 3}), integer(560)); // 8! / (2! 3! 3! 0!)
    g += 1; // a host-side change drops the device copy
    CHECK(!g.is_device_resident());
    CHECK_EQUAL(g.find_cf({0, 0, 0}), integer(2));
    CHECK_EQUAL((g * base).find_cf({0, 0, 0}), integer(2));
    // copies share the device terms; changing one does not change the other
    pt h = f;
    h -= x * x;
    CHECK_EQUAL(f.find_cf({2, 0, 0}), integer(1));
    CHECK_EQUAL(h.find_cf({2, 0, 0}), integer(0));
    // an empty product is an empty series with the symbol set kept
    pt::set_auto_truncate_degree(-1);
    const pt e = base * base;
    pt::unset_auto_truncate_degree();
    CHECK_EQUAL(e.size(), 0u);
    CHECK_EQUAL(e, 0);
    CHECK((e.get_symbol_set() == symbol_fset{"x", "y", "z"}));
}

static void unpacked_test()
{
    // unpacked exponent-vector keys go through the same GPU product (packed on the fly): benchmarks/fateman1_unpacked.cpp
    // (reduced to power 10: 10626 terms) and fateman1_unpacked_truncation.cpp:56-59 (degree 20 keeps exactly the terms of f)
    using ut = polynomial<integer, monomial<int>>;
    ut x{"x"}, y{"y"}, z{"z"}, t{"t"};
    auto f = 1 + x + y + z + t;
    const auto base(f);
    for (int i = 1; i < 10; ++i) f *= base;
    const auto g = f + 1;
    const auto r = f * g;
    CHECK_EQUAL(r.size(), 10626u);
    CHECK_EQUAL(r.find_cf({5, 5, 5, 5}), integer("11732745024"));
    CHECK_EQUAL(r.find_cf({0, 0, 0, 0}), integer(2));
    CHECK_EQUAL(ut::truncated_multiplication(f, g, 10).size(), 1001u);
    ut::set_auto_truncate_degree(10);
    CHECK_EQUAL((f * g).size(), 1001u);
    ut::set_auto_truncate_degree(3, {"x"});
    {
        const ut expect = pt_partial_reference<ut>(); // (leaves the partial truncation set)
        CHECK_EQUAL((x + y + 1).pow(3) * (x + 1), expect);
    }
    ut::unset_auto_truncate_degree();
    // same answer as the Kronecker-keyed series
    pt kx{"x"}, ky{"y"};
    const auto kr = (1 + kx + ky).pow(7) * (1 - kx + ky).pow(5);
    const auto ur = (1 + x + y).pow(7) * (1 - x + y).pow(5);
    CHECK_EQUAL(kr.size(), ur.size());
    for (const auto &term : ur._container()) {
        const auto e = term.m_key.unpack(ur.get_symbol_set());
        CHECK_EQUAL(kr.find_cf({e[0], e[1]}), term.m_cf);
    }
    CHECK_THROW((series_multiplier<ut>(x, y)), std::invalid_argument);
}

static void rational_test()
{
    // tests/polynomial_multiplier_02.cpp:114-154 (polynomial_multiplier_multiplier_finalise_test): rational coefficients
    // are brought to a common denominator, multiplied as integers (on the GPU) and divided back
    using qt = polynomial<rational, k_type>;
    {
        qt x{"x"}, y{"y"};
        CHECK_EQUAL(x * 4 / rational(3) * y * 5 / rational(2), rational(10, 3) * x * y);
        CHECK_EQUAL((x * 4 / rational(3) + y * 5 / rational(2)) * (x.pow(2) * 4 / rational(13) - y * 5 / rational(17)),
                    16 * x.pow(3) / 39 + rational(10, 13) * y * x * x - 20 * x * y / 51 - 25 * y * y / 34);
    }
    pt cmp;
    {
        pt x("x"), y("y"), z("z"), t("t");
        auto f = 1 + x + y + z + t;
        const auto tmp2 = f;
        for (int i = 1; i < 10; ++i) f *= tmp2;
        cmp = f * (f + 1);
    }
    qt res;
    {
        qt x("x"), y("y"), z("z"), t("t");
        auto f = 1 + x + y + z + t;
        const auto tmp2 = f;
        for (int i = 1; i < 10; ++i) f *= tmp2;
        const auto g = f + 1;
        res = f / 2 * g / 3;
    }
    const qt res6 = res * 6;
    CHECK_EQUAL(res6.size(), cmp.size());
    CHECK_EQUAL(cmp.size(), 10626u);
    std::size_t same = 0;
    for (const auto &t : cmp._container()) {
        const auto it = res6._container().find(qt::term_type(rational(0), t.m_key));
        if (it != res6._container().end() && it->m_cf == rational(t.m_cf)) ++same;
    }
    CHECK_EQUAL(same, cmp.size());
    // integer helpers behind the lcm / gcd passes
    CHECK_EQUAL(integer::gcd(integer("123456789012345678901234567890"), integer("987654321098765432109876543210")), integer("9000000000900000000090"));
    CHECK_EQUAL(integer("-1000000000000000000000000000007") / integer("1000000000000"), integer("-1000000000000000000"));
    CHECK_EQUAL(integer("-1000000000000000000000000000007") % integer("1000000000000"), integer(-7));
    CHECK_EQUAL(rational(6, -4), rational(-3, 2));
}

int main()
{
    RUN_TEST(residency_test);
    RUN_TEST(rational_test);
    RUN_TEST(unpacked_test);
    RUN_TEST(bounds_test);
    RUN_TEST(multiplication_test);
    RUN_TEST(pearce_test);
    RUN_TEST(truncation_test);
    RUN_TEST(threads_test);
    RUN_TEST(shared_operands_test);
    RUN_TEST(multi_gpu_test);
    return test_summary();
}
