// Host mirror of piranha::kronecker_monomial<std::int64_t> (reference: include/piranha/kronecker_monomial.hpp),
// restricted to what the multiplier path uses: get_int/set_int (:325-336), hash (:441-444), operator== / operator<
// (:452-455, :987-990), unpack (:554-557), is_compatible (:352-371), multiply (:427-436), degree (:1083-1119),
// pow (:760-790) and the symbol-set merge that series arithmetic performs before a multiplier is built
// (merge_symbols, :568-610).
#ifndef PB200_KRONECKER_MONOMIAL_HPP
#define PB200_KRONECKER_MONOMIAL_HPP

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "kronecker_array.hpp"

namespace pb200
{

// Sorted set of symbol names (piranha::symbol_fset is a boost flat_set<std::string>, symbol_utils.hpp:59).
using symbol_fset = std::set<std::string>;

inline std::size_t symbol_index(const symbol_fset &s, const std::string &name)
{
    std::size_t i = 0;
    for (const auto &n : s) {
        if (n == name) return i;
        ++i;
    }
    return s.size();
}

template <typename T = std::int64_t>
class kronecker_monomial
{
public:
    using value_type = T;
    using ka = kronecker_array<T>;
    using v_type = std::vector<T>;

    kronecker_monomial() = default;
    explicit kronecker_monomial(const symbol_fset &) {}
    kronecker_monomial(std::initializer_list<T> list) : m_value(ka::encode(v_type(list))) {}
    explicit kronecker_monomial(const v_type &v) : m_value(ka::encode(v)) {}
    static kronecker_monomial from_int(T n)
    {
        kronecker_monomial k;
        k.m_value = n;
        return k;
    }

    T get_int() const { return m_value; }
    void set_int(const T &n) { m_value = n; }
    std::size_t hash() const { return static_cast<std::size_t>(m_value); }
    bool operator==(const kronecker_monomial &o) const { return m_value == o.m_value; }
    bool operator!=(const kronecker_monomial &o) const { return m_value != o.m_value; }
    bool operator<(const kronecker_monomial &o) const { return m_value < o.m_value; }

    v_type unpack(const symbol_fset &args) const
    {
        v_type out(args.size());
        ka::decode(out, m_value);
        return out;
    }
    bool is_compatible(const symbol_fset &args) const
    {
        const auto &limits = ka::get_limits();
        const auto s = args.size();
        if (s == 0) return m_value == 0;
        if (s >= limits.size()) return false;
        return m_value >= std::get<1>(limits[s]) && m_value <= std::get<2>(limits[s]);
    }
    bool is_unitary(const symbol_fset &args) const
    {
        (void)args;
        return m_value == 0; // all exponents zero encode as 0 (kronecker_array.hpp:298-306)
    }
    // Total degree; checked addition like the reference's safe arithmetic.
    T degree(const symbol_fset &args) const
    {
        T d = 0;
        for (const T e : unpack(args)) d = safe_add(d, e);
        return d;
    }
    // Partial degree over the listed positions.
    T degree(const std::vector<std::size_t> &positions, const symbol_fset &args) const
    {
        const v_type v = unpack(args);
        T d = 0;
        for (const std::size_t p : positions) {
            if (p >= v.size()) throw std::invalid_argument("invalid position in the computation of the partial degree");
            d = safe_add(d, v[p]);
        }
        return d;
    }
    // res = t1 * t2: coefficient product, code sum with no overflow check (kronecker_monomial.hpp:417-436);
    // the bounds were verified by the multiplier's constructor.
    template <typename Term>
    static void multiply(Term &res, const Term &t1, const Term &t2, const symbol_fset &)
    {
        res.m_cf = t1.m_cf * t2.m_cf;
        res.m_key.set_int(static_cast<T>(t1.m_key.get_int() + t2.m_key.get_int()));
    }
    kronecker_monomial pow(long long n, const symbol_fset &args) const
    {
        v_type v = unpack(args);
        for (auto &e : v) {
            const __int128 p = static_cast<__int128>(e) * n;
            if (p > std::numeric_limits<T>::max() || p < std::numeric_limits<T>::min())
                throw std::overflow_error("overflow in the exponentiation of a Kronecker monomial");
            e = static_cast<T>(p);
        }
        try {
            return kronecker_monomial(v);
        } catch (const std::invalid_argument &) {
            // the reference reports exponents outside the codification limits as an overflow of the monomial
            throw std::overflow_error("overflow in the exponentiation of a Kronecker monomial");
        }
    }
    // Re-encode for a larger symbol set: `args` is the current set, `new_args` a superset.
    kronecker_monomial merge_symbols(const symbol_fset &args, const symbol_fset &new_args) const
    {
        const v_type old = unpack(args);
        v_type v;
        v.reserve(new_args.size());
        auto it = args.begin();
        std::size_t i = 0;
        for (const auto &name : new_args) {
            if (it != args.end() && *it == name) {
                v.push_back(old[i++]);
                ++it;
            } else {
                v.push_back(T(0));
            }
        }
        if (it != args.end()) throw std::invalid_argument("invalid argument(s) for symbol set merging");
        return kronecker_monomial(v);
    }

private:
    static T safe_add(T a, T b)
    {
        T r;
        if (__builtin_add_overflow(a, b, &r)) throw std::overflow_error("overflow in the computation of the degree");
        return r;
    }
    T m_value = 0;
};

} // namespace pb200

#endif
