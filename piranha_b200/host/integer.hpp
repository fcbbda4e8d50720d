// Host coefficient type of the drop-in: an exact signed integer in sign + magnitude form with 64-bit limbs,
// least significant first -- the layout of mppp::integer's static storage, which is what the C ABI ships
// (include/piranha_b200.h: pk_poly.mag / pk_poly.sign).
//
// Stands in for piranha::integer = mppp::integer<1> (reference: include/piranha/integer.hpp:67) on the host side of
// the multiplier path: construction, +, -, *, /, comparison, pow, printing.  Like mp++ it keeps small values in place
// (PB200_INTEGER_STATIC_LIMBS limbs, default 2: every coefficient of the benchmark results) and promotes to heap
// storage beyond, without limit; sizeof is 24 bytes, so a polynomial term {coefficient, Kronecker key} is 32 bytes
// and a node of the result container 40 -- the fill of a 5.8 M term result is bound by the bytes it writes.
// (The GPU product itself takes operands of up to four limbs and reports PK_E_CF_WIDTH beyond, SURVEY.md section 8b;
// that limit is the C ABI's, not this type's.)
#ifndef PB200_INTEGER_HPP
#define PB200_INTEGER_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifndef PB200_INTEGER_STATIC_LIMBS
#define PB200_INTEGER_STATIC_LIMBS 2
#endif

namespace pb200
{

class integer
{
public:
    static constexpr unsigned static_limbs = PB200_INTEGER_STATIC_LIMBS;
    static_assert(static_limbs >= 1, "at least one limb in place");
    using limb_t = std::uint64_t;

    integer() noexcept : m_size(0), m_cap(static_limbs), m_neg(0) { std::memset(m_u.s, 0, sizeof(m_u.s)); }
    ~integer()
    {
        if (!is_static()) delete[] m_u.d;
    }
    integer(const integer &o) : integer()
    {
        if (o.is_static()) { // (the common case: the whole object is its value)
            std::memcpy(static_cast<void *>(this), static_cast<const void *>(&o), sizeof(integer));
            return;
        }
        assign(o.data(), o.m_size);
        m_neg = o.m_neg;
    }
    integer(integer &&o) noexcept { steal(o); }
    integer &operator=(const integer &o)
    {
        if (this != &o) {
            if (o.is_static() && is_static()) {
                std::memcpy(static_cast<void *>(this), static_cast<const void *>(&o), sizeof(integer));
                return *this;
            }
            assign(o.data(), o.m_size);
            m_neg = o.m_neg;
        }
        return *this;
    }
    integer &operator=(integer &&o) noexcept
    {
        if (this != &o) {
            if (!is_static()) delete[] m_u.d;
            steal(o);
        }
        return *this;
    }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    integer(T n) : integer()
    {
        if (n < 0) {
            m_neg = true;
            // two's complement negate in the unsigned domain (also right for the most negative value)
            m_u.s[0] = static_cast<limb_t>(0) - static_cast<limb_t>(static_cast<long long>(n));
        } else {
            m_u.s[0] = static_cast<limb_t>(n);
        }
        m_size = m_u.s[0] ? 1u : 0u;
        if (!m_size) m_neg = false;
    }
    explicit integer(const std::string &s) : integer()
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
        if (n > 0x7FFFFFFFu) throw std::overflow_error("integer: too many limbs");
        r.assign(mag, static_cast<unsigned>(n));
        r.m_neg = sign < 0 && n;
        return r;
    }

    int sign() const { return m_size ? (m_neg ? -1 : 1) : 0; }
    bool is_zero() const { return m_size == 0; }
    std::size_t size() const { return m_size; } // limbs in use
    limb_t limb(std::size_t i) const { return i < m_size ? data()[i] : 0; }
    const limb_t *limbs() const { return data(); }
    bool is_static() const { return m_cap == static_limbs; } // (mppp::integer::is_static)
    unsigned bits() const { return m_size ? 64u * (m_size - 1u) + (64u - static_cast<unsigned>(__builtin_clzll(data()[m_size - 1]))) : 0u; }

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
    friend integer operator+(integer a, const integer &b) { return std::move(a += b); }
    friend integer operator-(integer a, const integer &b) { return std::move(a -= b); }
    friend integer operator*(const integer &a, const integer &b)
    {
        integer r;
        if (!a.m_size || !b.m_size) return r;
        const unsigned na = a.m_size, nb = b.m_size;
        scratch tmp(na + nb);
        const limb_t *pa = a.data(), *pb = b.data();
        for (unsigned i = 0; i < na; ++i) {
            limb_t carry = 0;
            for (unsigned j = 0; j < nb; ++j) {
                const unsigned __int128 t = static_cast<unsigned __int128>(pa[i]) * pb[j] + tmp.p[i + j] + carry;
                tmp.p[i + j] = static_cast<limb_t>(t);
                carry = static_cast<limb_t>(t >> 64);
            }
            tmp.p[i + nb] = carry;
        }
        unsigned n = na + nb;
        while (n && !tmp.p[n - 1]) --n;
        r.assign(tmp.p, n);
        r.m_neg = a.m_neg != b.m_neg;
        return r;
    }
    // multiply-accumulate: *this += a * b  (math::multiply_accumulate, reference integer.hpp:72-85)
    integer &multiply_accumulate(const integer &a, const integer &b) { return *this += a * b; }

    friend bool operator==(const integer &a, const integer &b)
    {
        return a.m_size == b.m_size && a.m_neg == b.m_neg && std::equal(a.data(), a.data() + a.m_size, b.data());
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
        if (a.m_size <= 2 && b.m_size <= 2) { // both fit 128 bits (every value of the benchmarks' rational passes): machine division
            const unsigned __int128 x = a.mag128(), y = b.mag128();
            const bool qneg = a.m_neg != b.m_neg, rneg = a.m_neg;
            q = from_u128(x / y, qneg);
            r = from_u128(x % y, rneg);
            return;
        }
        if (b.m_size == 1) { // one-limb divisor (the common denominators of the rational path): one pass of 128 / 64 divisions
            integer quo(a);
            const limb_t rm = quo.divmod_small(b.data()[0]);
            quo.m_neg = quo.m_size && (a.m_neg != b.m_neg);
            integer rem(rm);
            rem.m_neg = rem.m_size && a.m_neg;
            q = std::move(quo);
            r = std::move(rem);
            return;
        }
        integer quo, rem;
        const unsigned nb = a.bits();
        quo.reserve((nb + 63u) / 64u);
        for (unsigned i = nb; i-- > 0;) {
            rem.shl1((a.data()[i / 64] >> (i % 64)) & 1u);
            if (cmp_mag(rem, b) >= 0) {
                rem.m_neg = false;
                rem.add_signed(b, true); // rem -= |b|
                quo.data()[i / 64] |= limb_t(1) << (i % 64);
                if (i / 64 + 1 > quo.m_size) quo.m_size = i / 64 + 1;
            }
        }
        quo.m_neg = quo.m_size && (a.m_neg != b.m_neg);
        rem.m_neg = rem.m_size && a.m_neg;
        q = std::move(quo);
        r = std::move(rem);
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
            if (a.m_size <= 2 && b.m_size <= 2) { // both fit 128 bits: finish in machine arithmetic
                unsigned __int128 x = a.mag128(), y = b.mag128();
                while (y) {
                    if (!(x >> 64) && !(y >> 64)) { // down to one limb each
                        limb_t u = static_cast<limb_t>(x), v = static_cast<limb_t>(y);
                        while (v) {
                            const limb_t t = u % v;
                            u = v;
                            v = t;
                        }
                        return integer(u);
                    }
                    const unsigned __int128 t = x % y;
                    x = y;
                    y = t;
                }
                return from_u128(x, false);
            }
            integer t = a % b;
            a = std::move(b);
            b = std::move(t);
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
        const limb_t m = limb(0);
        if (m_size > 1 || (m_size == 1 && m > (m_neg ? (limb_t(1) << 63) : (limb_t(1) << 63) - 1)))
            throw std::overflow_error("integer: value does not fit a long long");
        return m_neg ? static_cast<long long>(0 - m) : static_cast<long long>(m);
    }

private:
    unsigned __int128 mag128() const { return (static_cast<unsigned __int128>(limb(1)) << 64) | limb(0); } // (m_size <= 2)
    static integer from_u128(unsigned __int128 v, bool neg)
    {
        const limb_t l[2] = {static_cast<limb_t>(v), static_cast<limb_t>(v >> 64)};
        return from_limbs(neg ? -1 : 1, l, 2);
    }
    // Storage.  Invariant: the limbs [m_size, m_cap) are zero.  m_cap == static_limbs: the limbs live in m_u.s.
    limb_t *data() { return is_static() ? m_u.s : m_u.d; }
    const limb_t *data() const { return is_static() ? m_u.s : m_u.d; }
    void reserve(unsigned n)
    {
        if (n <= m_cap) return;
        const unsigned cap = std::max(n, 2u * m_cap);
        limb_t *p = new limb_t[cap]();
        std::copy(data(), data() + m_size, p);
        if (!is_static()) delete[] m_u.d;
        m_u.d = p;
        m_cap = cap;
    }
    void set_size(unsigned n) // shrink (or keep) the size, restoring the invariant
    {
        limb_t *p = data();
        for (unsigned i = n; i < m_size; ++i) p[i] = 0;
        m_size = n;
    }
    void trim()
    {
        const limb_t *p = data();
        while (m_size && !p[m_size - 1]) --m_size;
    }
    void assign(const limb_t *src, unsigned n) // (src is never this object's own storage)
    {
        reserve(n);
        limb_t *p = data();
        std::copy(src, src + n, p);
        for (unsigned i = n; i < m_size; ++i) p[i] = 0;
        m_size = n;
    }
    void steal(integer &o) noexcept
    {
        std::memcpy(&m_u, &o.m_u, sizeof(m_u));
        m_size = o.m_size;
        m_cap = o.m_cap;
        m_neg = o.m_neg;
        std::memset(o.m_u.s, 0, sizeof(o.m_u.s));
        o.m_size = 0;
        o.m_cap = static_limbs;
        o.m_neg = false;
    }
    // a few limbs of scratch space: on the stack for the sizes that occur in practice
    struct scratch {
        explicit scratch(unsigned n) : p(n <= 16 ? local : (heap = new limb_t[n]))
        {
            std::fill(p, p + n, limb_t(0));
        }
        ~scratch() { delete[] heap; }
        scratch(const scratch &) = delete;
        scratch &operator=(const scratch &) = delete;
        limb_t local[16];
        limb_t *heap = nullptr;
        limb_t *p;
    };

    static int cmp_mag(const integer &a, const integer &b)
    {
        if (a.m_size != b.m_size) return a.m_size < b.m_size ? -1 : 1;
        const limb_t *pa = a.data(), *pb = b.data();
        for (unsigned i = a.m_size; i-- > 0;)
            if (pa[i] != pb[i]) return pa[i] < pb[i] ? -1 : 1;
        return 0;
    }
    void add_signed(const integer &o, bool o_neg) // (o may be *this)
    {
        if (!o.m_size) return;
        if (!m_size) {
            *this = o;
            m_neg = o_neg;
            return;
        }
        if (m_neg == o_neg) { // magnitudes add
            const unsigned n = std::max(m_size, o.m_size);
            reserve(n + 1u);
            limb_t *p = data();
            const limb_t *po = o.data(); // (after reserve: o may be this object)
            const unsigned no = o.m_size;
            limb_t carry = 0;
            for (unsigned i = 0; i < n; ++i) {
                const unsigned __int128 t = static_cast<unsigned __int128>(p[i]) + (i < no ? po[i] : 0) + carry;
                p[i] = static_cast<limb_t>(t);
                carry = static_cast<limb_t>(t >> 64);
            }
            m_size = n;
            if (carry) p[m_size++] = carry;
        } else { // magnitudes subtract, the larger one gives the sign
            const int c = cmp_mag(*this, o);
            if (c == 0) {
                set_size(0);
                m_neg = false;
                return;
            }
            const integer &big = c > 0 ? *this : o, &small = c > 0 ? o : *this;
            const bool neg = c > 0 ? m_neg : o_neg;
            const unsigned nb = big.m_size, ns = small.m_size;
            scratch out(nb);
            const limb_t *pb = big.data(), *ps = small.data();
            limb_t borrow = 0;
            for (unsigned i = 0; i < nb; ++i) {
                const limb_t x = pb[i], y = i < ns ? ps[i] : 0;
                const limb_t d = x - y - borrow;
                borrow = (x < y || (x == y && borrow)) ? 1u : 0u;
                out.p[i] = d;
            }
            unsigned n = nb;
            while (n && !out.p[n - 1]) --n;
            assign(out.p, n);
            m_neg = neg;
        }
    }
    void shl1(limb_t low_bit) // magnitude = magnitude * 2 + low_bit
    {
        reserve(m_size + 1u);
        limb_t *p = data();
        limb_t carry = low_bit;
        for (unsigned i = 0; i < m_size; ++i) {
            const limb_t nc = p[i] >> 63;
            p[i] = (p[i] << 1) | carry;
            carry = nc;
        }
        if (carry) p[m_size++] = carry;
    }
    void mul_small(limb_t m)
    {
        reserve(m_size + 1u);
        limb_t *p = data();
        limb_t carry = 0;
        for (unsigned i = 0; i < m_size; ++i) {
            const unsigned __int128 t = static_cast<unsigned __int128>(p[i]) * m + carry;
            p[i] = static_cast<limb_t>(t);
            carry = static_cast<limb_t>(t >> 64);
        }
        if (carry) p[m_size++] = carry;
    }
    void add_small(limb_t a)
    {
        reserve(m_size + 1u);
        limb_t *p = data();
        for (unsigned i = 0; a; ++i) { // (limbs at and above m_size are zero, and there is one of them)
            const limb_t s = p[i] + a;
            a = s < a ? 1u : 0u;
            p[i] = s;
            if (i >= m_size) m_size = i + 1;
        }
        trim();
    }
    limb_t divmod_small(limb_t d)
    {
        limb_t *p = data();
        unsigned __int128 rem = 0;
        for (unsigned i = m_size; i-- > 0;) {
            const unsigned __int128 cur = (rem << 64) | p[i];
            p[i] = static_cast<limb_t>(cur / d);
            rem = cur % d;
        }
        trim();
        return static_cast<limb_t>(rem);
    }

    union storage {
        limb_t s[static_limbs];
        limb_t *d;
    } m_u;
    std::uint32_t m_size;
    std::uint32_t m_cap : 31;
    std::uint32_t m_neg : 1;
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
