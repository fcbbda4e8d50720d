// Host mirror of piranha::kronecker_array<std::int64_t> (reference: include/piranha/kronecker_array.hpp).
//
// Same public surface for the multiplier path: get_limits() (:251-254), encode() (:274-307), decode() (:329-365),
// same exceptions.  The limits table is produced by the same perturbed-prime search (:128-229) driven by
// std::mt19937(m) and std::uniform_int_distribution<int>(-5, 5), so with the same C++ standard library it is the
// table piranha itself would build.  Arithmetic that the reference does with mp++ integers is done here with
// __int128 (every intermediate of the search stays below 2^113) and a deterministic Miller-Rabin next-prime.
#ifndef PB200_KRONECKER_ARRAY_HPP
#define PB200_KRONECKER_ARRAY_HPP

#include <cstddef>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace pb200
{

namespace detail
{
using i128 = __int128;
using u128 = unsigned __int128;

inline u128 mulmod(u128 a, u128 b, u128 m)
{
    if (m <= (u128(1) << 64)) {
        return (a % m) * (b % m) % m; // both factors < 2^64
    }
    u128 r = 0;
    a %= m;
    while (b) { // double-and-add; only reached for moduli above 2^64 (the discarded last search iteration)
        if (b & 1) {
            r = (r >= m - a) ? r - (m - a) : r + a;
        }
        a = (a >= m - a) ? a - (m - a) : a + a;
        b >>= 1;
    }
    return r;
}

inline u128 powmod(u128 b, u128 e, u128 m)
{
    u128 r = 1;
    b %= m;
    while (e) {
        if (e & 1) r = mulmod(r, b, m);
        b = mulmod(b, b, m);
        e >>= 1;
    }
    return r;
}

inline bool is_prime(u128 n)
{
    static const unsigned bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (unsigned p : bases) {
        if (n % p == 0) return n == p;
    }
    u128 d = n - 1;
    unsigned s = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++s;
    }
    for (unsigned a : bases) { // deterministic below 3.3e24
        u128 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (unsigned r = 1; r < s; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                composite = false;
                break;
            }
        }
        if (composite) return false;
    }
    return true;
}

// mppp::integer::nextprime(): the smallest prime strictly greater than n.
inline i128 next_prime(i128 n)
{
    u128 c = n < 1 ? 2 : static_cast<u128>(n) + 1;
    while (!is_prime(c)) ++c;
    return static_cast<i128>(c);
}
} // namespace detail

template <typename SignedInteger = std::int64_t>
class kronecker_array
{
public:
    using int_type = SignedInteger;
    using size_type = std::size_t;
    // (M vector, h_min, h_max, h_max - h_min), kronecker_array.hpp:96-103.
    using limit_type = std::tuple<std::vector<int_type>, int_type, int_type, int_type>;
    using limits_type = std::vector<limit_type>;

private:
    using i128 = detail::i128;
    static bool fits(i128 x)
    {
        return x >= static_cast<i128>(std::numeric_limits<int_type>::min())
               && x <= static_cast<i128>(std::numeric_limits<int_type>::max());
    }
    static i128 dot(const std::vector<i128> &a, const std::vector<i128> &b)
    {
        i128 r = 0;
        for (size_type i = 0; i < a.size(); ++i) r += a[i] * b[i];
        return r;
    }
    static limit_type determine_limit(size_type m)
    {
        std::mt19937 engine(static_cast<std::mt19937::result_type>(m));
        std::uniform_int_distribution<int> dist(-5, 5);
        auto perturb = [&](i128 &arg) {
            arg += (static_cast<i128>(dist(engine)) * arg) / 100;
            arg = detail::next_prime(arg);
        };
        std::vector<i128> lo(m, -1), hi(m, 1), c(m, 1), pc, plo, phi;
        for (size_type i = 1; i < m; ++i) c[i] = c[i - 1] * 3;
        for (;;) {
            const i128 h_min = dot(c, lo), h_max = dot(c, hi);
            if (!(fits(h_min) && fits(h_max) && fits(h_max - h_min + 1))) {
                std::vector<int_type> out;
                if (pc.empty()) return std::make_tuple(out, int_type(0), int_type(0), int_type(0));
                const i128 a = dot(pc, plo), b = dot(pc, phi);
                for (auto v : phi) out.push_back(static_cast<int_type>(v));
                return std::make_tuple(out, static_cast<int_type>(a), static_cast<int_type>(b),
                                       static_cast<int_type>(b - a));
            }
            pc = c;
            plo = lo;
            phi = hi;
            for (size_type i = 1; i < m; ++i) {
                i128 delta = c[i] / pc[i - 1]; // the radix of component i-1
                delta *= 2;
                perturb(delta);
                c[i] = delta * c[i - 1];
            }
            for (size_type i = 0; i + 1 < m; ++i) {
                hi[i] = (c[i + 1] / c[i] - 1) / 2;
                lo[i] = -hi[i];
            }
            hi[m - 1] = (4 * hi[m - 1] + 1) / 2;
            perturb(hi[m - 1]);
            lo[m - 1] = -hi[m - 1];
        }
    }
    static limits_type determine_limits()
    {
        limits_type out;
        out.emplace_back(std::vector<int_type>{}, int_type(0), int_type(0), int_type(0));
        for (size_type m = 1;; ++m) {
            auto l = determine_limit(m);
            if (std::get<0>(l).empty()) break;
            out.push_back(std::move(l));
        }
        return out;
    }

public:
    static const limits_type &get_limits()
    {
        static const limits_type table = determine_limits();
        return table;
    }
    template <typename Vector>
    static int_type encode(const Vector &v)
    {
        const auto &limits = get_limits();
        const auto size = v.size();
        if (size >= limits.size()) throw std::invalid_argument("size of vector to be encoded is too large");
        if (!size) return int_type(0);
        const auto &M = std::get<0>(limits[size]);
        for (size_type i = 0; i < size; ++i) {
            const i128 x = static_cast<i128>(v[i]);
            if (x < -static_cast<i128>(M[i]) || x > static_cast<i128>(M[i])) {
                throw std::invalid_argument("a component of the vector to be encoded is out of bounds");
            }
        }
        int_type code = static_cast<int_type>(static_cast<int_type>(v[0]) + M[0]);
        int_type radix_prod = static_cast<int_type>(2 * M[0] + 1);
        for (size_type i = 1; i < size; ++i) {
            code = static_cast<int_type>(code + (static_cast<int_type>(v[i]) + M[i]) * radix_prod);
            radix_prod = static_cast<int_type>(radix_prod * (2 * M[i] + 1));
        }
        return static_cast<int_type>(code + std::get<1>(limits[size]));
    }
    template <typename Vector>
    static void decode(Vector &retval, const int_type &n)
    {
        using v_type = typename Vector::value_type;
        const auto &limits = get_limits();
        const auto m = retval.size();
        if (m >= limits.size()) throw std::invalid_argument("size of vector to be decoded is too large");
        if (!m) {
            if (n != 0) throw std::invalid_argument("a vector of size 0 must always be encoded as 0");
            return;
        }
        const auto &M = std::get<0>(limits[m]);
        const int_type hmin = std::get<1>(limits[m]), hmax = std::get<2>(limits[m]);
        if (n < hmin || n > hmax) throw std::invalid_argument("the integer to be decoded is out of bounds");
        int_type code = static_cast<int_type>(n - hmin);
        for (size_type i = 0; i < m; ++i) {
            const int_type radix = static_cast<int_type>(2 * M[i] + 1);
            retval[i] = static_cast<v_type>(code % radix - M[i]);
            code = static_cast<int_type>(code / radix);
        }
    }
};

} // namespace pb200

#endif
