// Host mirror of piranha::monomial<T> (reference: include/piranha/monomial.hpp): the unpacked exponent-vector key used by
// benchmarks/fateman1_unpacked.cpp, pearce1_unpacked.cpp and fateman1_unpacked_truncation.cpp -- only the interface a
// polynomial and its multiplier need (hash :951-966, equality, unpack, degree, pow, symbol merge).
//
// series_multiplier<polynomial<integer, monomial<T>>> (below) packs the exponent vectors to Kronecker codes on the fly
// (kronecker_array::encode, kronecker_array.hpp:274-307), runs the same GPU product as the Kronecker-keyed polynomial and
// unpacks the result: SURVEY.md section 8f rank 4 -- wider coverage of operator* without a new kernel.
#ifndef PB200_MONOMIAL_HPP
#define PB200_MONOMIAL_HPP

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "polynomial.hpp"

namespace pb200
{

template <typename T = int>
class monomial
{
public:
    using value_type = T;
    using v_type = std::vector<T>;

    monomial() = default;
    explicit monomial(const symbol_fset &args) : m_v(args.size(), T(0)) {}
    monomial(std::initializer_list<T> list) : m_v(list) {}
    explicit monomial(const v_type &v) : m_v(v) {}
    template <typename U>
    static monomial from_vector(const std::vector<U> &v)
    {
        monomial m;
        m.m_v.reserve(v.size());
        for (const U &e : v) {
            if (e > static_cast<U>(std::numeric_limits<T>::max()) || e < static_cast<U>(std::numeric_limits<T>::min()))
                throw std::overflow_error("monomial exponent out of the range of its value type");
            m.m_v.push_back(static_cast<T>(e));
        }
        return m;
    }
    static monomial from_int(std::int64_t) { throw std::logic_error("an unpacked monomial has no integer code"); }

    std::size_t size() const { return m_v.size(); }
    std::size_t hash() const // hash-combine of the exponents (any hash will do: the container only needs consistency)
    {
        std::size_t h = 0;
        for (const T &e : m_v) h ^= std::hash<long long>()(static_cast<long long>(e)) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return h;
    }
    bool operator==(const monomial &o) const { return m_v == o.m_v; }
    bool operator!=(const monomial &o) const { return m_v != o.m_v; }
    bool operator<(const monomial &o) const
    {
        // the order of the Kronecker codes: the last exponent is the most significant digit
        for (std::size_t i = m_v.size(); i-- > 0;)
            if (m_v[i] != o.m_v[i]) return m_v[i] < o.m_v[i];
        return false;
    }
    v_type unpack(const symbol_fset &args) const
    {
        if (args.size() != m_v.size()) throw std::invalid_argument("incompatible symbol set");
        return m_v;
    }
    bool is_compatible(const symbol_fset &args) const { return args.size() == m_v.size(); }
    T degree(const symbol_fset &args) const
    {
        (void)args;
        T d = 0;
        for (const T &e : m_v)
            if (__builtin_add_overflow(d, e, &d)) throw std::overflow_error("overflow in the computation of the degree");
        return d;
    }
    monomial pow(long long n, const symbol_fset &) const
    {
        monomial r(*this);
        for (auto &e : r.m_v) {
            const __int128 p = static_cast<__int128>(e) * n;
            if (p > std::numeric_limits<T>::max() || p < std::numeric_limits<T>::min())
                throw std::overflow_error("overflow in the exponentiation of a monomial");
            e = static_cast<T>(p);
        }
        return r;
    }
    monomial merge_symbols(const symbol_fset &args, const symbol_fset &new_args) const
    {
        monomial r;
        r.m_v.reserve(new_args.size());
        auto it = args.begin();
        std::size_t i = 0;
        for (const auto &name : new_args) {
            if (it != args.end() && *it == name) {
                r.m_v.push_back(m_v[i++]);
                ++it;
            } else {
                r.m_v.push_back(T(0));
            }
        }
        if (it != args.end()) throw std::invalid_argument("invalid argument(s) for symbol set merging");
        return r;
    }

private:
    v_type m_v;
};

// Unpacked keys: pack -> the Kronecker multiplier -> unpack.  Same protocol and members as the Kronecker specialisation;
// exponents outside the Kronecker limits raise std::overflow_error from the constructor (the reference's unpacked
// multiplier has no such limit: it checks the exponent type's own range instead, monomial.hpp:1050-1100).
template <typename T>
class series_multiplier<polynomial<integer, monomial<T>>, void>
{
public:
    using series_type = polynomial<integer, monomial<T>>;
    using packed_type = polynomial<integer, kronecker_monomial<std::int64_t>>;
    using degree_type = typename series_type::degree_type;

    explicit series_multiplier(const series_type &s1, const series_type &s2)
        : m_ss(s1.get_symbol_set()), m_p1(pack(s1)), m_p2(pack(s2)), m_mult(check(s1, s2, m_p1), m_p2)
    {
    }
    series_multiplier(const series_multiplier &) = delete;
    series_multiplier &operator=(const series_multiplier &) = delete;

    series_type operator()() const
    {
        // the auto-truncation state belongs to the series type the user sees
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
    series_type _untruncated_multiplication() const { return unpack(m_mult._untruncated_multiplication()); }
    series_type _truncated_multiplication(const degree_type &max_degree) const
    {
        return unpack(m_mult._truncated_multiplication(static_cast<std::int64_t>(max_degree)));
    }
    series_type _truncated_multiplication(const degree_type &max_degree, const std::vector<std::string> &names,
                                          const std::vector<std::size_t> &idx) const
    {
        return unpack(m_mult._truncated_multiplication(static_cast<std::int64_t>(max_degree), names, idx));
    }

private:
    static const packed_type &check(const series_type &s1, const series_type &s2, const packed_type &p1)
    {
        if (s1.get_symbol_set() != s2.get_symbol_set()) throw std::invalid_argument("incompatible arguments sets");
        return p1;
    }
    static packed_type pack(const series_type &s)
    {
        packed_type r;
        r.set_symbol_set(s.get_symbol_set());
        auto &c = r._container();
        c.rehash(std::max<std::size_t>(1, s.size()));
        std::vector<std::int64_t> e(s.get_symbol_set().size());
        for (const auto &t : s._container()) {
            const auto v = t.m_key.unpack(s.get_symbol_set());
            for (std::size_t i = 0; i < v.size(); ++i) e[i] = static_cast<std::int64_t>(v[i]);
            std::int64_t code;
            try {
                code = kronecker_array<std::int64_t>::encode(e);
            } catch (const std::invalid_argument &) {
                throw std::overflow_error("monomial exponents are outside the Kronecker codification limits");
            }
            c.insert(typename packed_type::term_type(t.m_cf, kronecker_monomial<std::int64_t>::from_int(code)));
        }
        return r;
    }
    series_type unpack(const packed_type &p) const
    {
        series_type r;
        r.set_symbol_set(m_ss);
        auto &c = r._container();
        c.rehash(std::max<std::size_t>(1, p.size()));
        for (const auto &t : p._container())
            c.insert(typename series_type::term_type(t.m_cf, monomial<T>::from_vector(t.m_key.unpack(m_ss))));
        return r;
    }

    symbol_fset m_ss;
    packed_type m_p1, m_p2;
    series_multiplier<packed_type> m_mult;
};

} // namespace pb200

#endif
