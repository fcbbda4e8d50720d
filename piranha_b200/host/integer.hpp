// Host coefficient type of the drop-in: an exact signed integer in sign + magnitude form with 64-bit limbs,
// least significant first -- the layout of mppp::integer's static storage, which is what the C ABI ships
// (include/piranha_b200.h: pk_poly.mag / pk_poly.sign).
//
// Stands in for piranha::integer = mppp::integer<1> (reference: include/piranha/integer.hpp:67) on the host side of
// the multiplier path: construction, +, -, *, comparison, pow, printing.  mp++ promotes to heap storage without
// limit; this type has a fixed capacity of PB200_INTEGER_LIMBS limbs and throws std::overflow_error beyond it
// (a fixed-width path must report, never wrap).
#ifndef PB200_INTEGER_HPP
#define PB200_INTEGER_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>

#ifndef PB200_INTEGER_LIMBS
#define PB200_INTEGER_LIMBS 6
#endif

namespace pb200
{

class integer
{
public:
    static constexpr std::size_t max_limbs = PB200_INTEGER_LIMBS;
    using limb_t = std::uint64_t;

    integer() = default;
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    integer(T n)
    {
        if (n < 0) {
            m_neg = true;
            // two's complement negate in the unsigned domain (also right for the most negative value)
            m_mag[0] = static_cast<limb_t>(0) - static_cast<limb_t>(static_cast<long long>(n));
        } else {
            m_mag[0] = static_cast<limb_t>(n);
        }
        m_size = m_mag[0] ? 1u : 0u;
        if (!m_size) m_neg = false;
    }
    explicit integer(const std::string &s)
    {
        std::size_t i = 0;
        bool neg = false;
        if (i < s.size() && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
        if (i == s.size()) throw std::invalid_argument("invalid string for the construction of an integer");
        for (; i < s.size(); ++i) {
            if (s[i] < '0' || s[i] > '9') throw std::invalid_argument("invalid string for the construction of an integer");
            mul_small(10u);
            add_small(static_cast<limb_t>(s[i] - '0'));
        }
        m_neg = neg && m_size;
    }
    // sign in {-1, 0, +1} and `n` magnitude limbs (the C-ABI term format)
    static integer from_limbs(int sign, const limb_t *mag, std::size_t n)
    {
        integer r;
        while (n && !mag[n - 1]) --n;
        if (n > max_limbs) throw std::overflow_error("integer: magnitude exceeds the static capacity");
        std::copy(mag, mag + n, r.m_mag.begin());
        r.m_size = static_cast<unsigned>(n);
        r.m_neg = sign < 0 && n;
        return r;
    }

    int sign() const { return m_size ? (m_neg ? -1 : 1) : 0; }
    bool is_zero() const { return m_size == 0; }
    std::size_t size() const { return m_size; } // limbs in use
    limb_t limb(std::size_t i) const { return i < m_size ? m_mag[i] : 0; }
    const limb_t *limbs() const { return m_mag.data(); }
    unsigned bits() const { return m_size ? 64u * (m_size - 1u) + (64u - static_cast<unsigned>(__builtin_clzll(m_mag[m_size - 1]))) : 0u; }

    integer operator-() const
    {
        integer r(*this);
        if (r.m_size) r.m_neg = !r.m_neg;
        return r;
    }
    integer &operator+=(const integer &o)
    {
        add_signed(o, o.m_neg);
        return *this;
    }
    integer &operator-=(const integer &o)
    {
        add_signed(o, !o.m_neg);
        return *this;
    }
    integer &operator*=(const integer &o)
    {
        *this = *this * o;
        return *this;
    }
    friend integer operator+(integer a, const integer &b) { return a += b; }
    friend integer operator-(integer a, const integer &b) { return a -= b; }
    friend integer operator*(const integer &a, const integer &b)
    {
        integer r;
        if (!a.m_size || !b.m_size) return r;
        limb_t tmp[2 * max_limbs] = {};
        for (unsigned i = 0; i < a.m_size; ++i) {
            limb_t carry = 0;
            for (unsigned j = 0; j < b.m_size; ++j) {
                const unsigned __int128 t = static_cast<unsigned __int128>(a.m_mag[i]) * b.m_mag[j] + tmp[i + j] + carry;
                tmp[i + j] = static_cast<limb_t>(t);
                carry = static_cast<limb_t>(t >> 64);
            }
            tmp[i + b.m_size] = carry;
        }
        unsigned n = a.m_size + b.m_size;
        while (n && !tmp[n - 1]) --n;
        if (n > max_limbs) throw std::overflow_error("integer: product exceeds the static capacity");
        std::copy(tmp, tmp + n, r.m_mag.begin());
        r.m_size = n;
        r.m_neg = a.m_neg != b.m_neg;
        return r;
    }
    // multiply-accumulate: *this += a * b  (math::multiply_accumulate, reference integer.hpp:72-85)
    integer &multiply_accumulate(const integer &a, const integer &b) { return *this += a * b; }

    friend bool operator==(const integer &a, const integer &b)
    {
        return a.m_size == b.m_size && a.m_neg == b.m_neg && std::equal(a.m_mag.begin(), a.m_mag.begin() + a.m_size, b.m_mag.begin());
    }
    friend bool operator!=(const integer &a, const integer &b) { return !(a == b); }
    friend bool operator<(const integer &a, const integer &b)
    {
        if (a.m_neg != b.m_neg) return a.m_neg;
        const int c = cmp_mag(a, b);
        return a.m_neg ? c > 0 : c < 0;
    }
    friend bool operator>(const integer &a, const integer &b) { return b < a; }
    friend bool operator<=(const integer &a, const integer &b) { return !(b < a); }
    friend bool operator>=(const integer &a, const integer &b) { return !(a < b); }

    // truncating division (quotient rounded towards zero, remainder with the sign of the dividend), by binary long division:
    // only the rational coefficient path uses it (lcm / gcd of denominators), never the product itself
    static void divmod(const integer &a, const integer &b, integer &q, integer &r)
    {
        if (!b.m_size) throw std::domain_error("integer division by zero");
        integer quo, rem;
        const unsigned nb = a.bits();
        for (unsigned i = nb; i-- > 0;) {
            rem.shl1((a.m_mag[i / 64] >> (i % 64)) & 1u);
            if (cmp_mag(rem, b) >= 0) {
                integer d(b);
                d.m_neg = false;
                rem.m_neg = false;
                rem -= d;
                quo.m_mag[i / 64] |= limb_t(1) << (i % 64);
                if (i / 64 + 1 > quo.m_size) quo.m_size = i / 64 + 1;
            }
        }
        quo.m_neg = quo.m_size && (a.m_neg != b.m_neg);
        rem.m_neg = rem.m_size && a.m_neg;
        q = quo;
        r = rem;
    }
    friend integer operator/(const integer &a, const integer &b)
    {
        integer q, r;
        divmod(a, b, q, r);
        return q;
    }
    friend integer operator%(const integer &a, const integer &b)
    {
        integer q, r;
        divmod(a, b, q, r);
        return r;
    }
    integer abs() const
    {
        integer r(*this);
        r.m_neg = false;
        return r;
    }
    static integer gcd(integer a, integer b)
    {
        a.m_neg = b.m_neg = false;
        while (b.m_size) {
            if (a.m_size <= 1 && b.m_size <= 1) { // both fit one limb: finish in machine arithmetic
                limb_t x = a.m_mag[0], y = b.m_mag[0];
                while (y) {
                    const limb_t t = x % y;
                    x = y;
                    y = t;
                }
                return integer(x);
            }
            integer t = a % b;
            a = b;
            b = t;
        }
        return a;
    }

    integer pow(unsigned n) const
    {
        integer r(1), b(*this);
        while (n) {
            if (n & 1u) r *= b;
            n >>= 1;
            if (n) b *= b;
        }
        return r;
    }
    std::string to_string() const
    {
        if (!m_size) return "0";
        integer t(*this);
        std::string s;
        while (t.m_size) s.push_back(static_cast<char>('0' + t.divmod_small(10u)));
        if (m_neg) s.push_back('-');
        std::reverse(s.begin(), s.end());
        return s;
    }
    friend std::ostream &operator<<(std::ostream &os, const integer &n) { return os << n.to_string(); }
    explicit operator long long() const
    {
        if (m_size > 1 || (m_size == 1 && m_mag[0] > (m_neg ? (limb_t(1) << 63) : (limb_t(1) << 63) - 1)))
            throw std::overflow_error("integer: value does not fit a long long");
        return m_neg ? static_cast<long long>(0 - m_mag[0]) : static_cast<long long>(m_mag[0]);
    }

private:
    static int cmp_mag(const integer &a, const integer &b)
    {
        if (a.m_size != b.m_size) return a.m_size < b.m_size ? -1 : 1;
        for (unsigned i = a.m_size; i-- > 0;)
            if (a.m_mag[i] != b.m_mag[i]) return a.m_mag[i] < b.m_mag[i] ? -1 : 1;
        return 0;
    }
    void add_signed(const integer &o, bool o_neg)
    {
        if (!o.m_size) return;
        if (!m_size) {
            *this = o;
            m_neg = o_neg;
            return;
        }
        if (m_neg == o_neg) { // magnitudes add
            limb_t carry = 0;
            const unsigned n = std::max(m_size, o.m_size);
            for (unsigned i = 0; i < n; ++i) {
                const unsigned __int128 t = static_cast<unsigned __int128>(limb(i)) + o.limb(i) + carry;
                m_mag[i] = static_cast<limb_t>(t);
                carry = static_cast<limb_t>(t >> 64);
            }
            m_size = n;
            if (carry) {
                if (n == max_limbs) throw std::overflow_error("integer: sum exceeds the static capacity");
                m_mag[m_size++] = carry;
            }
        } else { // magnitudes subtract, the larger one gives the sign
            const int c = cmp_mag(*this, o);
            if (c == 0) {
                *this = integer();
                return;
            }
            const integer &big = c > 0 ? *this : o, &small = c > 0 ? o : *this;
            const bool neg = c > 0 ? m_neg : o_neg;
            limb_t borrow = 0;
            std::array<limb_t, max_limbs> out{};
            for (unsigned i = 0; i < big.m_size; ++i) {
                const limb_t x = big.m_mag[i], y = small.limb(i);
                const limb_t d = x - y - borrow;
                borrow = (x < y || (x == y && borrow)) ? 1u : 0u;
                out[i] = d;
            }
            unsigned n = big.m_size;
            while (n && !out[n - 1]) --n;
            m_mag = out;
            m_size = n;
            m_neg = neg;
        }
    }
    void shl1(limb_t low_bit) // magnitude = magnitude * 2 + low_bit
    {
        limb_t carry = low_bit;
        for (unsigned i = 0; i < m_size; ++i) {
            const limb_t nc = m_mag[i] >> 63;
            m_mag[i] = (m_mag[i] << 1) | carry;
            carry = nc;
        }
        if (carry) {
            if (m_size == max_limbs) throw std::overflow_error("integer: value exceeds the static capacity");
            m_mag[m_size++] = carry;
        }
    }
    void mul_small(limb_t m)
    {
        limb_t carry = 0;
        for (unsigned i = 0; i < m_size; ++i) {
            const unsigned __int128 t = static_cast<unsigned __int128>(m_mag[i]) * m + carry;
            m_mag[i] = static_cast<limb_t>(t);
            carry = static_cast<limb_t>(t >> 64);
        }
        if (carry) {
            if (m_size == max_limbs) throw std::overflow_error("integer: value exceeds the static capacity");
            m_mag[m_size++] = carry;
        }
    }
    void add_small(limb_t a)
    {
        for (unsigned i = 0; a && i < max_limbs; ++i) {
            const limb_t s = m_mag[i] + a;
            a = s < a ? 1u : 0u;
            m_mag[i] = s;
            if (i >= m_size) m_size = i + 1;
        }
        if (a) throw std::overflow_error("integer: value exceeds the static capacity");
        while (m_size && !m_mag[m_size - 1]) --m_size;
    }
    limb_t divmod_small(limb_t d)
    {
        unsigned __int128 rem = 0;
        for (unsigned i = m_size; i-- > 0;) {
            const unsigned __int128 cur = (rem << 64) | m_mag[i];
            m_mag[i] = static_cast<limb_t>(cur / d);
            rem = cur % d;
        }
        while (m_size && !m_mag[m_size - 1]) --m_size;
        return static_cast<limb_t>(rem);
    }

    std::array<limb_t, max_limbs> m_mag{};
    unsigned m_size = 0;
    bool m_neg = false;
};

namespace math
{
// any coefficient type with an is_zero() member (integer, rational)
template <typename T>
inline auto is_zero(const T &n) -> decltype(n.is_zero())
{
    return n.is_zero();
}
inline void multiply_accumulate(integer &x, const integer &y, const integer &z) { x.multiply_accumulate(y, z); }
inline void mul3(integer &r, const integer &a, const integer &b) { r = a * b; }
inline integer pow(const integer &b, unsigned n) { return b.pow(n); }
} // namespace math

} // namespace pb200

#endif
