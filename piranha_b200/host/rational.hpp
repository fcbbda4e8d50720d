// Host mirror of piranha::rational (mp++ rational) for the multiplier path, and the series_multiplier specialisation for
// rational coefficients.  The reference multiplies series with rational coefficients by
// bringing each operand to a common denominator first (base_series_multiplier.hpp:157-206: lcm of the denominators,
// numerators scaled), multiplying the INTEGER series, and dividing by lcm1 * lcm2 at the end (finalise_series,
// base_series_multiplier.hpp:275-326).  The same here: the integer product is the GPU product, the pre/post passes are
// host code.
#ifndef PB200_RATIONAL_HPP
#define PB200_RATIONAL_HPP

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <type_traits>

#include "polynomial.hpp"

namespace pb200
{

class rational
{
public:
    rational() : m_den(1) {}
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    rational(T n) : m_num(n), m_den(1)
    {
    }
    rational(const integer &n) : m_num(n), m_den(1) {}
    rational(const integer &n, const integer &d) : m_num(n), m_den(d) { canonicalise(); }
    template <typename T, typename U, typename std::enable_if<std::is_integral<T>::value && std::is_integral<U>::value, int>::type = 0>
    rational(T n, U d) : m_num(n), m_den(d)
    {
        canonicalise();
    }
    const integer &num() const { return m_num; }
    const integer &den() const { return m_den; }
    bool is_zero() const { return m_num.is_zero(); }
    rational operator-() const
    {
        rational r(*this);
        r.m_num = -r.m_num;
        return r;
    }
    rational &operator+=(const rational &o)
    {
        m_num = m_num * o.m_den + o.m_num * m_den;
        m_den *= o.m_den;
        canonicalise();
        return *this;
    }
    rational &operator-=(const rational &o) { return *this += -o; }
    rational &operator*=(const rational &o)
    {
        m_num *= o.m_num;
        m_den *= o.m_den;
        canonicalise();
        return *this;
    }
    rational &operator/=(const rational &o)
    {
        if (o.is_zero()) throw std::domain_error("rational division by zero");
        m_num *= o.m_den;
        m_den *= o.m_num;
        canonicalise();
        return *this;
    }
    friend rational operator+(rational a, const rational &b) { return a += b; }
    friend rational operator-(rational a, const rational &b) { return a -= b; }
    friend rational operator*(rational a, const rational &b) { return a *= b; }
    friend rational operator/(rational a, const rational &b) { return a /= b; }
    friend bool operator==(const rational &a, const rational &b) { return a.m_num == b.m_num && a.m_den == b.m_den; }
    friend bool operator!=(const rational &a, const rational &b) { return !(a == b); }
    rational pow(unsigned n) const { return rational(m_num.pow(n), m_den.pow(n)); }
    friend std::ostream &operator<<(std::ostream &os, const rational &q)
    {
        os << q.m_num;
        if (q.m_den != integer(1)) os << '/' << q.m_den;
        return os;
    }

private:
    void canonicalise()
    {
        if (m_den.is_zero()) throw std::domain_error("rational with zero denominator");
        if (m_den.sign() < 0) {
            m_num = -m_num;
            m_den = -m_den;
        }
        if (m_den.size() == 1 && m_den.limb(0) == 1) return; // integral value: nothing to reduce
        const integer g = integer::gcd(m_num, m_den);
        if (g != integer(1) && !g.is_zero()) {
            m_num = m_num / g;
            m_den = m_den / g;
        }
        if (m_num.is_zero()) m_den = integer(1);
    }
    integer m_num, m_den;
};

// series / x and series * q for rational series (tests write x * 4 / 3_q)
template <typename Key, typename X>
inline polynomial<rational, Key> operator/(const polynomial<rational, Key> &p, const X &x)
{
    return p * polynomial<rational, Key>(rational(1) / rational(x));
}
template <typename Key>
inline polynomial<rational, Key> operator*(const polynomial<rational, Key> &p, const rational &q)
{
    return p * polynomial<rational, Key>(q);
}
template <typename Key>
inline polynomial<rational, Key> operator*(const rational &q, const polynomial<rational, Key> &p)
{
    return p * polynomial<rational, Key>(q);
}

template <>
class series_multiplier<polynomial<rational, kronecker_monomial<std::int64_t>>, void>
{
public:
    using key_type = kronecker_monomial<std::int64_t>;
    using series_type = polynomial<rational, key_type>;
    using int_series = polynomial<integer, key_type>;
    using degree_type = series_type::degree_type;

    explicit series_multiplier(const series_type &s1, const series_type &s2)
        : m_t0(std::chrono::steady_clock::now()), m_ss(s1.get_symbol_set()), m_lcm1(lcm_of(s1)), m_lcm2(lcm_of(s2)), m_p1(scaled(s1, m_lcm1)),
          m_p2(scaled(s2, m_lcm2)), m_t1(std::chrono::steady_clock::now()), m_mult(check(s1, s2, m_p1), m_p2)
    {
        if (std::getenv("PB200_TRACE"))
            std::fprintf(stderr, "[pb200 rational] lcm + scaled operands %.3f ms, multiplier construction (bounds, uploads) %.3f ms\n",
                         std::chrono::duration<double, std::milli>(m_t1 - m_t0).count(),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - m_t1).count());
    }
    series_multiplier(const series_multiplier &) = delete;
    series_multiplier &operator=(const series_multiplier &) = delete;

    series_type operator()() const
    {
        const auto t = series_type::get_auto_truncate_degree();
        if (std::get<0>(t) == 0) return _untruncated_multiplication();
        if (std::get<0>(t) == 1) return _truncated_multiplication(std::get<1>(t));
        std::vector<std::string> names(std::get<2>(t).begin(), std::get<2>(t).end());
        std::vector<std::size_t> idx;
        for (const auto &n : names) {
            const std::size_t i = symbol_index(m_ss, n);
            if (i < m_ss.size()) idx.push_back(i);
        }
        return _truncated_multiplication(std::get<1>(t), names, idx);
    }
    series_type _untruncated_multiplication() const
    {
        if (!std::getenv("PB200_TRACE")) return finalise(m_mult._untruncated_multiplication());
        const auto t0 = std::chrono::steady_clock::now();
        const int_series p = m_mult._untruncated_multiplication();
        const auto t1 = std::chrono::steady_clock::now();
        series_type r = finalise(p);
        const auto t2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[pb200 rational] integer product %.3f ms, result over the common denominator %.3f ms\n",
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
        return r;
    }
    series_type _truncated_multiplication(const degree_type &max_degree) const
    {
        return finalise(m_mult._truncated_multiplication(max_degree));
    }
    series_type _truncated_multiplication(const degree_type &max_degree, const std::vector<std::string> &names,
                                          const std::vector<std::size_t> &idx) const
    {
        return finalise(m_mult._truncated_multiplication(max_degree, names, idx));
    }

private:
    static const int_series &check(const series_type &s1, const series_type &s2, const int_series &p1)
    {
        if (s1.get_symbol_set() != s2.get_symbol_set()) throw std::invalid_argument("incompatible arguments sets");
        return p1;
    }
    // lcm of the denominators (base_series_multiplier.hpp:165-180)
    static integer lcm_of(const series_type &s)
    {
        integer l(1);
        for (const auto &t : s._container()) {
            const integer &d = t.m_cf.den();
            if (d.size() == 1 && d.limb(0) == 1) continue; // an integral coefficient leaves the lcm as it is
            l = l / integer::gcd(l, d) * d;
        }
        return l;
    }
    // numerators over the common denominator (base_series_multiplier.hpp:181-200)
    static int_series scaled(const series_type &s, const integer &lcm)
    {
        int_series r;
        r.set_symbol_set(s.get_symbol_set());
        auto &c = r._container();
        c.rehash(std::max<std::size_t>(1, s.size()));
        const bool lcm_is_one = lcm.size() == 1 && lcm.limb(0) == 1;
        std::size_t n = 0;
        for (const auto &t : s._container()) { // (the keys are those of s: unique, so the low-level insertion of the multipliers does)
            const integer &d = t.m_cf.den();
            const bool d_is_one = d.size() == 1 && d.limb(0) == 1;
            integer v = lcm_is_one ? t.m_cf.num() : (d_is_one ? t.m_cf.num() * lcm : t.m_cf.num() * (lcm / d));
            int_series::term_type it(std::move(v), t.m_key);
            const auto b = c._bucket(it);
            c._unique_insert(std::move(it), b);
            ++n;
        }
        c._update_size(n);
        return r;
    }
    // den = lcm1 * lcm2, canonicalised (finalise_impl, base_series_multiplier.hpp:275-326)
    series_type finalise(const int_series &p) const
    {
        const integer den = m_lcm1 * m_lcm2;
        series_type r;
        r.set_symbol_set(m_ss);
        auto &c = r._container();
        if (!p.m_host_valid.load(std::memory_order_acquire)) { // product still in HBM: fill the rational container straight from the downloaded terms
            const gpu::handle h = p.m_dev->h;
            std::lock_guard<std::recursive_mutex> ctx_lock(h->mx);
            gpu::fill_from_device(h->ctx, p.m_dev->p, c, [&den](const pk_result &res, std::size_t i) {
                return series_type::term_type(rational(integer::from_limbs(res.sign[i], res.mag + i * res.limbs, res.limbs), den),
                                              key_type::from_int(res.key[i]));
            });
            return r;
        }
        c.rehash(std::max<std::size_t>(1, p.size()));
        for (const auto &t : p._container()) c.insert(series_type::term_type(rational(t.m_cf, den), t.m_key));
        return r;
    }

    std::chrono::steady_clock::time_point m_t0; // (PB200_TRACE=1 prints the phases)
    symbol_fset m_ss;
    integer m_lcm1, m_lcm2;
    int_series m_p1, m_p2;
    std::chrono::steady_clock::time_point m_t1;
    series_multiplier<int_series> m_mult;
};

} // namespace pb200

#endif
