// CPU-only tests of the host mirror (no GPU, no product): integer arithmetic, Kronecker codification, hash_set,
// series addition and symbol merging.  Follows the reference's tests/integer_01.cpp:149-175 (multiply_accumulate),
// tests/kronecker_array.cpp:52-160 (limits structure, encode/decode round trips), tests/kronecker_monomial_01.cpp:404-461
// (key multiply) and tests/hash_set_01.cpp (low-level interface).
#include <cstdint>
#include <cstring>
#include <atomic>
#include <random>
#include <thread>
#include <string>
#include <vector>

#include "../../piranha_b200/host/polynomial.hpp"
#include "../../piranha_b200/host/rational.hpp"
#include "check.hpp"

using namespace pb200;

static void integer_test()
{
    CHECK_EQUAL(integer(0).sign(), 0);
    CHECK_EQUAL(integer(-5).to_string(), "-5");
    CHECK_EQUAL(integer("123456789012345678901234567890").to_string(), "123456789012345678901234567890");
    CHECK_EQUAL((integer("-99999999999999999999") + integer("99999999999999999999")).is_zero(), true);
    CHECK_EQUAL(integer(std::numeric_limits<long long>::min()).to_string(), "-9223372036854775808");
    std::mt19937_64 rng(42);
    for (int it = 0; it < 20000; ++it) {
        const long long a = static_cast<long long>(rng()) >> (rng() % 40), b = static_cast<long long>(rng()) >> (rng() % 40);
        const __int128 p = static_cast<__int128>(a) * b, s = static_cast<__int128>(a) + b, d = static_cast<__int128>(a) - b;
        auto to128 = [](const integer &n) {
            unsigned __int128 m = (static_cast<unsigned __int128>(n.limb(1)) << 64) | n.limb(0);
            return n.sign() < 0 ? -static_cast<__int128>(m) : static_cast<__int128>(m);
        };
        CHECK(to128(integer(a) * integer(b)) == p);
        CHECK(to128(integer(a) + integer(b)) == s);
        CHECK(to128(integer(a) - integer(b)) == d);
        // multiply_accumulate against ordinary operators (integer_01.cpp:149-175)
        integer x(a), y(b), z(a ^ b);
        integer acc = x;
        math::multiply_accumulate(acc, y, z);
        CHECK(acc == x + y * z);
        CHECK((integer(a) < integer(b)) == (a < b));
    }
    // wide values: (2^64 - 1)^4 and string round trip
    integer w("18446744073709551615");
    integer w4 = w * w * w * w;
    CHECK_EQUAL(w4.to_string(), "115792089237316195398462578067141184799968521174335529155754622898352762650625");
    CHECK_EQUAL(w4.bits(), 256u);
    CHECK(integer(w4.to_string()) == w4);
    // beyond the static storage the value is promoted to the heap (mppp::integer), without limit
    static_assert(sizeof(integer) == 24, "integer: two limbs in place + size, capacity and sign");
    CHECK(integer(5).is_static() && w.is_static() && !w4.is_static());
    const integer w12 = w4 * w4 * w4;
    CHECK_EQUAL(w12.bits(), 768u);
    CHECK_EQUAL(w12, w.pow(12));
    CHECK_EQUAL(w12 / w4, w4 * w4);
    CHECK((w12 % w4).is_zero());
    CHECK_EQUAL(integer(w12.to_string()), w12);
    {
        integer a = w12, b = w4;      // copies and moves between static and promoted values
        a = b;
        CHECK_EQUAL(a, w4);
        a = integer(7);
        CHECK_EQUAL(a, integer(7));
        b = std::move(a);
        CHECK_EQUAL(b, integer(7));
        CHECK(a.is_zero());
        a = w12;
        a -= w12;
        CHECK(a.is_zero() && a.sign() == 0);
        a += w12;
        a += a;                       // aliased operand
        CHECK_EQUAL(a, w12 * integer(2));
        a -= w12;
        a -= w12 - integer(1);
        CHECK_EQUAL(a, integer(1));
        CHECK_EQUAL(a.limb(0), 1u);
        CHECK_EQUAL(a.limb(1), 0u);
    }
    CHECK_EQUAL(integer(3).pow(40).to_string(), "12157665459056928801");
    CHECK_THROW(integer("12x"), std::invalid_argument);
}

static void rational_test()
{
    // division against __int128 on random operands, then the identities a = q * b + r, |r| < |b|, sign(r) = sign(a)
    std::mt19937_64 rng(7);
    for (int i = 0; i < 2000; ++i) {
        const std::int64_t a = static_cast<std::int64_t>(rng()) >> (rng() % 40), b = static_cast<std::int64_t>(rng()) >> (rng() % 60);
        if (!b) continue;
        const integer A(a), B(b);
        const integer W = A * integer(static_cast<std::int64_t>(rng())) * integer(static_cast<std::int64_t>(rng())) + integer(a);
        CHECK_EQUAL(A / B, integer(a / b));
        CHECK_EQUAL(A % B, integer(a % b));
        integer q, r;
        integer::divmod(W, B, q, r);
        CHECK_EQUAL(q * B + r, W);
        CHECK(r.abs() < B.abs());
        CHECK(r.is_zero() || r.sign() == W.sign());
        const integer g = integer::gcd(W, B);
        CHECK((W % g).is_zero() && (B % g).is_zero());
        CHECK_EQUAL(integer::gcd(W / g, B / g), integer(1));
    }
    CHECK_THROW(integer(1) / integer(0), std::domain_error);
    CHECK_EQUAL(integer::gcd(integer(0), integer(-12)), integer(12));
    // canonical form, arithmetic
    CHECK_EQUAL(rational(6, -4), rational(-3, 2));
    CHECK_EQUAL(rational(0, -5), rational(0));
    CHECK_EQUAL(rational(0, -5).den(), integer(1));
    CHECK_THROW(rational(1, 0), std::domain_error);
    CHECK_EQUAL(rational(1, 2) + rational(1, 3), rational(5, 6));
    CHECK_EQUAL(rational(1, 2) - rational(1, 2), rational(0));
    CHECK_EQUAL(rational(4, 3) * rational(5, 2), rational(10, 3));
    CHECK_EQUAL(rational(4, 3) / rational(-2, 9), rational(-6));
    CHECK_THROW(rational(1) / rational(0), std::domain_error);
    CHECK_EQUAL(rational(-2, 3).pow(3), rational(-8, 27));
    // rational series: linear operations stay on the host
    using qt = polynomial<rational, kronecker_monomial<std::int64_t>>;
    qt x{"x"}, y{"y"};
    const qt s = x + y - x;
    CHECK_EQUAL(s, y);
    CHECK_EQUAL((x + x).size(), 1u);
}

static void kronecker_array_test()
{
    using ka = kronecker_array<std::int64_t>;
    const auto &l = ka::get_limits();
    CHECK(l.size() > 1u);
    CHECK(std::get<0>(l[0]).empty());
    for (std::size_t m = 1; m < l.size(); ++m) { // kronecker_array.cpp:52-83
        CHECK_EQUAL(std::get<0>(l[m]).size(), m);
        for (auto v : std::get<0>(l[m])) CHECK(v > 0);
        CHECK(std::get<1>(l[m]) < 0);
        CHECK(std::get<2>(l[m]) > 0);
        CHECK(std::get<3>(l[m]) > 0);
    }
    std::mt19937 rng(7);
    for (std::size_t m = 1; m < std::min<std::size_t>(l.size(), 12); ++m) { // round trips, :100-160
        const auto &M = std::get<0>(l[m]);
        std::vector<std::int64_t> v(m), out(m);
        for (int it = 0; it < 200; ++it) {
            for (std::size_t i = 0; i < m; ++i) v[i] = it == 0 ? -M[i] : (it == 1 ? M[i] : std::uniform_int_distribution<std::int64_t>(-M[i], M[i])(rng));
            ka::decode(out, ka::encode(v));
            CHECK(out == v);
        }
        v[0] = M[0] + 1;
        CHECK_THROW(ka::encode(v), std::invalid_argument);
    }
    std::vector<std::int64_t> empty;
    CHECK_EQUAL(ka::encode(empty), 0);
    CHECK_THROW(ka::encode(std::vector<std::int64_t>(l.size(), 0)), std::invalid_argument);
}

static void kronecker_monomial_test()
{
    using k_type = kronecker_monomial<std::int64_t>;
    using term_type = term<integer, k_type>;
    // kronecker_monomial_01.cpp:404-461
    symbol_fset ed{"x"};
    term_type t1(integer(2), k_type{1}), t2(integer(3), k_type{2}), res;
    k_type::multiply(res, t1, t2, ed);
    CHECK_EQUAL(res.m_cf, integer(6));
    CHECK_EQUAL(res.m_key.get_int(), 3);
    symbol_fset xy{"x", "y"};
    term_type t3(integer(2), k_type{1, -1}), t4(integer(-4), k_type{2, 0});
    k_type::multiply(res, t3, t4, xy);
    CHECK_EQUAL(res.m_cf, integer(-8));
    CHECK((res.m_key.unpack(xy) == std::vector<std::int64_t>{3, -1}));
    CHECK_EQUAL((k_type{1, 2}.degree(xy)), 3);
    CHECK_EQUAL((k_type{1, 2}.degree({1}, xy)), 2);
    CHECK((k_type{1, 2}.pow(3, xy).unpack(xy) == std::vector<std::int64_t>{3, 6}));
    CHECK((k_type{5}.merge_symbols(symbol_fset{"y"}, symbol_fset{"x", "y", "z"}).unpack(symbol_fset{"x", "y", "z"})
           == std::vector<std::int64_t>{0, 5, 0}));
    CHECK(k_type{}.is_compatible(symbol_fset{}));
    CHECK(!k_type::from_int(1).is_compatible(symbol_fset{}));
}

static void hash_set_test()
{
    struct H {
        std::size_t operator()(int v) const { return static_cast<std::size_t>(v); }
    };
    hash_set<int, H> h;
    CHECK(h.empty());
    CHECK_EQUAL(h.bucket_count(), 0u);
    for (int i = 0; i < 1000; ++i) CHECK(h.insert(i * 7).second);
    CHECK(!h.insert(7).second);
    CHECK_EQUAL(h.size(), 1000u);
    CHECK((h.bucket_count() & (h.bucket_count() - 1)) == 0u);
    CHECK(h.load_factor() <= 1.0);
    for (int i = 0; i < 1000; ++i) CHECK(h.find(i * 7) != h.end());
    CHECK(h.find(3) == h.end());
    for (int i = 0; i < 1000; i += 2) h.erase(h.find(i * 7));
    CHECK_EQUAL(h.size(), 500u);
    std::size_t n = 0;
    for (auto it = h.begin(); it != h.end(); ++it) ++n;
    CHECK_EQUAL(n, 500u);
    // low-level interface: rehash, _bucket, _unique_insert, _update_size (base_series_multiplier.cpp:448-515 uses it the same way)
    hash_set<int, H> g;
    g.rehash(100);
    CHECK_EQUAL(g.bucket_count(), 128u);
    for (int i = 0; i < 100; ++i) g._unique_insert(i, g._bucket(i));
    CHECK_EQUAL(g.size(), 0u);
    g._update_size(100);
    CHECK_EQUAL(g.size(), 100u);
    for (int i = 0; i < 100; ++i) CHECK(g._find(i, g._bucket(i)) != g.end());
    g.rehash(1); // would exceed the maximum load factor: refused
    CHECK_EQUAL(g.bucket_count(), 128u);
    g.rehash(1000);
    CHECK_EQUAL(g.bucket_count(), 1024u);
    for (int i = 0; i < 100; ++i) CHECK(g.find(i) != g.end());
    // thread-parallel bulk fill: same contents, same iteration order whatever the thread count; the set stays usable
    const std::size_t nbulk = 50000;
    std::vector<int> order1;
    for (unsigned nt : {1u, 2u, 5u, 8u}) {
        hash_set<int, H> b;
        b.insert(-1); // replaced by the fill
        b._bulk_unique_fill(nbulk, [](std::size_t i) { return static_cast<int>(i * 3); }, [](std::size_t i) { return i * 3; }, nt);
        CHECK_EQUAL(b.size(), nbulk);
        CHECK_EQUAL(b.bucket_count(), 65536u);
        CHECK(b.find(-1) == b.end());
        std::size_t found = 0;
        for (std::size_t i = 0; i < nbulk; ++i) found += b.find(static_cast<int>(i * 3)) != b.end();
        CHECK_EQUAL(found, nbulk);
        CHECK(b.find(1) == b.end());
        std::vector<int> order(b.begin(), b.end());
        if (nt == 1u) order1 = order;
        CHECK(order == order1);
        b.erase(b.find(3));
        CHECK(b.insert(1).second);
        CHECK(!b.insert(6).second);
        CHECK_EQUAL(b.size(), nbulk);
        hash_set<int, H> c(b); // copies and moves of the node pool
        CHECK_EQUAL(c.size(), nbulk);
        CHECK(c.find(1) != c.end() && c.find(3) == c.end());
        hash_set<int, H> d(std::move(c));
        CHECK_EQUAL(d.size(), nbulk);
        CHECK(d.find(9) != d.end());
        // an element constructor that throws leaves an empty, reusable set
        bool thrown = false;
        try {
            b._bulk_unique_fill(nbulk, [](std::size_t i) { if (i == 31234) throw std::overflow_error("x"); return static_cast<int>(i); },
                                [](std::size_t i) { return i; }, nt);
        } catch (const std::overflow_error &) {
            thrown = true;
        }
        CHECK(thrown);
        CHECK(b.empty());
        CHECK(b.insert(5).second);
        // the same fill for elements that arrive in pieces (a result crossing PCIe): a producer thread announces them
        for (unsigned n_chunks : {1u, 7u, 64u}) {
            std::atomic<std::size_t> ready{0};
            std::atomic<bool> failed{false};
            auto wait = [&](std::size_t hi) {
                while (ready.load(std::memory_order_acquire) < hi) {
                    if (failed.load()) return false;
                    std::this_thread::yield();
                }
                return true;
            };
            std::thread producer([&] {
                for (std::size_t r = 0; r < nbulk; r += 777) ready.store(std::min<std::size_t>(nbulk, r + 777), std::memory_order_release);
            });
            hash_set<int, H> e;
            e.insert(-1);
            CHECK(e._bulk_unique_fill_streamed(nbulk, [](std::size_t i) { return static_cast<int>(i * 3); }, [](std::size_t i) { return i * 3; }, nt,
                                               n_chunks, wait));
            producer.join();
            CHECK_EQUAL(e.size(), nbulk);
            CHECK(e.find(-1) == e.end());
            std::size_t hits = 0;
            for (std::size_t i = 0; i < nbulk; ++i) hits += e.find(static_cast<int>(i * 3)) != e.end();
            CHECK_EQUAL(hits, nbulk);
            CHECK((std::vector<int>(e.begin(), e.end()) == order1)); // node order = element order, whatever the pieces
            // pieces = groups of elements that own disjoint bucket ranges (grouped by the top bits of the bucket): plain
            // stores on the bucket heads
            {
                const std::size_t nb = 65536, n_groups = 16;
                std::vector<int> grouped;
                std::vector<std::size_t> goff(n_groups + 1, 0);
                for (std::size_t g = 0; g < n_groups; ++g) {
                    goff[g] = grouped.size();
                    for (std::size_t i = 0; i < nbulk; ++i)
                        if (((i * 3) & (nb - 1)) >> 12 == g) grouped.push_back(static_cast<int>(i * 3));
                }
                goff[n_groups] = grouped.size();
                CHECK_EQUAL(grouped.size(), nbulk);
                std::atomic<std::size_t> ready2{0};
                std::thread producer2([&] {
                    for (std::size_t r = 0; r < nbulk; r += 1000) ready2.store(std::min<std::size_t>(nbulk, r + 1000), std::memory_order_release);
                });
                hash_set<int, H> x;
                CHECK(x._bulk_unique_fill_streamed(
                    nbulk, [&](std::size_t i) { return grouped[i]; }, [&](std::size_t i) { return static_cast<std::size_t>(grouped[i]); }, nt,
                    static_cast<unsigned>(n_groups),
                    [&](std::size_t hi) {
                        while (ready2.load(std::memory_order_acquire) < hi) std::this_thread::yield();
                        return true;
                    },
                    [&](unsigned g) { return std::make_pair(goff[g], goff[g + 1]); }, true));
                producer2.join();
                CHECK_EQUAL(x.size(), nbulk);
                CHECK_EQUAL(x.bucket_count(), nb);
                std::size_t hits2 = 0;
                for (std::size_t i = 0; i < nbulk; ++i) hits2 += x.find(static_cast<int>(i * 3)) != x.end();
                CHECK_EQUAL(hits2, nbulk);
                CHECK(x.find(1) == x.end());
                // ranges that do not tile [0, n): refused, empty set
                CHECK(!x._bulk_unique_fill_streamed(
                    nbulk, [&](std::size_t i) { return grouped[i]; }, [&](std::size_t i) { return static_cast<std::size_t>(grouped[i]); }, nt,
                    static_cast<unsigned>(n_groups), [](std::size_t) { return true; },
                    [&](unsigned g) { return std::make_pair(goff[g], g + 1 == n_groups ? goff[g + 1] - 1 : goff[g + 1]); }, true));
                CHECK(x.empty());
            }
            // a producer that gives up half way: false, and an empty, reusable set
            ready.store(nbulk / 2);
            failed.store(true);
            CHECK(!e._bulk_unique_fill_streamed(nbulk, [](std::size_t i) { return static_cast<int>(i); }, [](std::size_t i) { return i; }, nt,
                                                n_chunks, wait));
            CHECK(e.empty());
            CHECK(e.insert(5).second);
            // an element constructor that throws
            failed.store(false);
            ready.store(nbulk);
            bool thrown2 = false;
            try {
                e._bulk_unique_fill_streamed(nbulk, [](std::size_t i) { if (i == 31234) throw std::overflow_error("x"); return static_cast<int>(i); },
                                             [](std::size_t i) { return i; }, nt, n_chunks, wait);
            } catch (const std::overflow_error &) {
                thrown2 = true;
            }
            CHECK(thrown2);
            CHECK(e.empty());
        }
    }
    // large arrays come from 2 MB aligned blocks that are recycled: the next block of about the same size is the freed one
    {
        using detail::big_alloc;
        using detail::big_free;
        const std::size_t bytes = (std::size_t(9) << 20) + 12345;
        void *p1 = big_alloc(bytes);
        CHECK((reinterpret_cast<std::uintptr_t>(p1) & ((std::size_t(2) << 20) - 1)) == 0u);
        std::memset(p1, 0xAB, bytes);
        big_free(p1, bytes);
        void *p2 = big_alloc(bytes - 4096); // a little smaller: same block
        CHECK(p2 == p1);
        void *p3 = big_alloc(bytes); // the cache is empty now: a new block
        CHECK(p3 != p2);
        void *p4 = big_alloc(bytes * 4); // much larger than anything kept
        big_free(p2, bytes - 4096);
        big_free(p3, bytes);
        big_free(p4, bytes * 4);
        void *small = big_alloc(100); // small arrays go to malloc
        big_free(small, 100);
        big_free(nullptr, 0);
        hash_set<int, H> big1;
        big1._bulk_unique_fill(3u << 20, [](std::size_t i) { return static_cast<int>(i); }, [](std::size_t i) { return i; }, 4u);
        CHECK_EQUAL(big1.size(), std::size_t(3u << 20));
        big1.clear();
        big1._bulk_unique_fill(3u << 20, [](std::size_t i) { return static_cast<int>(i) + 1; }, [](std::size_t i) { return i + 1; }, 4u);
        CHECK(big1.find(3 << 20) != big1.end() && big1.find(0) == big1.end());
    }
    CHECK(settings::get_n_threads() >= 1u);
    settings::set_n_threads(3u);
    CHECK_EQUAL(settings::get_n_threads(), 3u);
    CHECK_THROW(settings::set_n_threads(0u), std::invalid_argument);
    settings::reset_n_threads();
    CHECK(settings::get_n_threads() >= 1u);
}

static void series_test()
{
    using pt = polynomial<integer, kronecker_monomial<std::int64_t>>;
    pt x{"x"}, y{"y"};
    CHECK_EQUAL(x.size(), 1u);
    CHECK((x.get_symbol_set() == symbol_fset{"x"}));
    auto s = x + y + 1;
    CHECK_EQUAL(s.size(), 3u);
    CHECK((s.get_symbol_set() == symbol_fset{"x", "y"}));
    CHECK_EQUAL(s - s, 0);
    CHECK_EQUAL((s - y - x), 1);
    CHECK_EQUAL((x + x).find_cf({1}), integer(2));
    CHECK_EQUAL((2 * x * 3).find_cf({1}), integer(6));
    CHECK_EQUAL((x * 0).size(), 0u);
    CHECK((x.pow(-3).get_symbol_set() == symbol_fset{"x"}));
    CHECK_EQUAL(x.pow(-3).find_cf({-3}), integer(1));
    CHECK_EQUAL(x.pow(5).degree(), 5);
    CHECK(x + y == y + x);
    CHECK(x + y != y - x);
    // the primary template has no call operator: only the specialised series type is multipliable
    static_assert(std::is_empty<series_multiplier<int>>::value, "primary template must be empty");
    // auto-truncation state (polynomial_truncation.cpp:69-80, :170-177)
    auto tup = pt::get_auto_truncate_degree();
    CHECK_EQUAL(std::get<0>(tup), 0);
    pt::set_auto_truncate_degree(3);
    tup = pt::get_auto_truncate_degree();
    CHECK_EQUAL(std::get<0>(tup), 1);
    CHECK_EQUAL(std::get<1>(tup), 3);
    pt::set_auto_truncate_degree(1, {"x", "z"});
    tup = pt::get_auto_truncate_degree();
    CHECK_EQUAL(std::get<0>(tup), 2);
    CHECK((std::get<2>(tup) == symbol_fset{"x", "z"}));
    pt::unset_auto_truncate_degree();
    tup = pt::get_auto_truncate_degree();
    CHECK_EQUAL(std::get<0>(tup), 0);
    CHECK(std::get<2>(tup).empty());
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--limits") { // the table the library codifies with, for tests/test_cpp_host.py
        const auto &l = kronecker_array<std::int64_t>::get_limits();
        for (std::size_t m = 0; m < l.size(); ++m) {
            std::printf("%zu %lld", m, static_cast<long long>(std::get<1>(l[m])));
            for (auto v : std::get<0>(l[m])) std::printf(" %lld", static_cast<long long>(v));
            std::printf("\n");
        }
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "--calc") { // exact arithmetic on request, for tests/test_cpp_host.py
        // lines "op a b" (integers: add sub mul div mod gcd) or "qop an ad bn bd" (rationals: qadd qsub qmul qdiv), decimal
        std::string op, x[4];
        while (std::cin >> op) {
            try {
                if (op[0] == 'q') {
                    std::cin >> x[0] >> x[1] >> x[2] >> x[3];
                    const rational a{integer{x[0]}, integer{x[1]}}, b{integer{x[2]}, integer{x[3]}};
                    const rational r = op == "qadd" ? a + b : op == "qsub" ? a - b : op == "qmul" ? a * b : a / b;
                    std::cout << r.num() << ' ' << r.den() << '\n';
                } else {
                    std::cin >> x[0] >> x[1];
                    const integer a{x[0]}, b{x[1]};
                    const integer r = op == "add" ? a + b : op == "sub" ? a - b : op == "mul" ? a * b : op == "div" ? a / b
                                      : op == "mod" ? a % b : integer::gcd(a, b);
                    std::cout << r << '\n';
                }
            } catch (const std::exception &e) {
                std::cout << "error " << e.what() << '\n';
            }
        }
        return 0;
    }
    RUN_TEST(integer_test);
    RUN_TEST(rational_test);
    RUN_TEST(kronecker_array_test);
    RUN_TEST(kronecker_monomial_test);
    RUN_TEST(hash_set_test);
    RUN_TEST(series_test);
    return test_summary();
}
