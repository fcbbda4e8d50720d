// Host mirror of the reference's operator interface for the multiplier path:
//
//   term<Cf, Key>                      include/piranha/term.hpp (mutable coefficient, hash/equality on the key only)
//   polynomial<Cf, Key>                the part of include/piranha/polynomial.hpp + series.hpp that operator* needs:
//                                      construction from a symbol name or a constant, + - *, pow, size, equality,
//                                      get_symbol_set, _container, the auto-truncation statics (:646-740) and
//                                      truncated_multiplication / untruncated_multiplication (:813-893)
//   series_multiplier<polynomial<integer, kronecker_monomial<int64_t>>>
//                                      THE DROP-IN: the class-template protocol of series_multiplier.hpp:37-56 --
//                                      construct from two series, operator()() returns the product -- implemented
//                                      by marshalling the terms to the C ABI (include/piranha_b200.h) and filling
//                                      the result container with rehash / _unique_insert / _update_size exactly as
//                                      sparse_kronecker_multiplication does (polynomial.hpp:1666-1667, :1750-1754,
//                                      base_series_multiplier.hpp:934).
//
// The arithmetic of the product runs ONLY on the GPU (libpiranha_b200.so); this header contains no multiplication
// loop.  Series addition/subtraction and symbol merging are host code, as in the reference.
#ifndef PB200_POLYNOMIAL_HPP
#define PB200_POLYNOMIAL_HPP

#include <atomic>
#include <condition_variable>
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/piranha_b200.h"
#include "hash_set.hpp"
#include "integer.hpp"
#include "settings.hpp"
#include "kronecker_monomial.hpp"
#include "series_multiplier.hpp"

namespace pb200
{

template <typename Cf, typename Key>
struct term {
    using cf_type = Cf;
    using key_type = Key;
    term() = default;
    term(Cf cf, Key key) : m_cf(std::move(cf)), m_key(std::move(key)) {}
    bool operator==(const term &o) const { return m_key == o.m_key; }
    std::size_t hash() const { return m_key.hash(); }
    bool is_zero(const symbol_fset &) const { return math::is_zero(m_cf); }
    bool is_compatible(const symbol_fset &args) const { return m_key.is_compatible(args); }
    mutable Cf m_cf{};
    Key m_key{};
};

// std::atomic<bool> that copies by value (series are copyable; a copy starts from the source's current state)
struct copyable_atomic_bool {
    std::atomic<bool> v;
    copyable_atomic_bool(bool b = true) : v(b) {}
    copyable_atomic_bool(const copyable_atomic_bool &o) : v(o.v.load(std::memory_order_acquire)) {}
    copyable_atomic_bool &operator=(const copyable_atomic_bool &o)
    {
        v.store(o.v.load(std::memory_order_acquire), std::memory_order_release);
        return *this;
    }
    bool load(std::memory_order m = std::memory_order_seq_cst) const { return v.load(m); }
    void store(bool b, std::memory_order m = std::memory_order_seq_cst) { v.store(b, m); }
};

struct term_hasher {
    template <typename Term>
    std::size_t operator()(const Term &t) const
    {
        return t.hash();
    }
};

// ONE pk_ctx per process, shared and reference-counted: every device-resident series keeps the context that owns its
// terms alive (so a static / global polynomial can be destroyed after main() returns), and calls into the context are
// serialised by its mutex -- operator* may be called concurrently from different threads on shared const operands
// (base_series_multiplier.hpp:404-408), a pk_ctx is used by one thread at a time.  The context spreads every product over
// settings::get_n_gpus() devices (pk_ctx_create_multi; environment PB200_DEVICES = "0,1,2,3" or PB200_N_GPUS = 4), the
// analogue of settings::set_n_threads() sizing the reference's thread pool.
class gpu
{
public:
    struct holder {
        pk_ctx *ctx = nullptr;
        std::recursive_mutex mx;
        unsigned n_gpus = 1;
        ~holder()
        {
            if (ctx) pk_ctx_destroy(ctx);
        }
    };
    using handle = std::shared_ptr<holder>;
    // the current context (created on first use, re-created when settings::set_n_gpus() changed the device count)
    static handle shared()
    {
        std::lock_guard<std::mutex> lk(slot_mutex());
        handle &cur = slot();
        const unsigned want = settings::get_n_gpus();
        if (cur && cur->n_gpus == want) return cur;
        auto h = std::make_shared<holder>();
        std::vector<int> ids = device_list(want);
        int rc;
        if (ids.size() == 1) rc = pk_ctx_create(ids[0], nullptr, &h->ctx);
        else rc = pk_ctx_create_multi(static_cast<std::uint32_t>(ids.size()), ids.data(), &h->ctx);
        if (rc != PK_OK) throw std::runtime_error("piranha_b200: no usable sm_100 GPU (pk_ctx_create failed); there is no CPU fallback");
        h->n_gpus = want;
        cur = h;
        return cur;
    }
    // status code -> the exception the reference throws (include/piranha_b200.h)
    static void check(pk_ctx *ctx, int rc)
    {
        if (rc == PK_OK) return;
        const std::string msg = pk_last_error(ctx);
        switch (rc) {
        case PK_E_ARGS: throw std::invalid_argument(msg);
        case PK_E_KEY_RANGE:
        case PK_E_DEGREE:
        case PK_E_CF_WIDTH: throw std::overflow_error(msg);
        case PK_E_NOMEM: throw std::bad_alloc();
        default: throw std::runtime_error(msg);
        }
    }

    // Device terms -> host container: rehash + _unique_insert + _update_size, as sparse_kronecker_multiplication fills its
    // result (polynomial.hpp:1666-1667, :1750-1754, base_series_multiplier.hpp:934).  make(result, i) builds term i.  The
    // caller holds the context's lock.
    // Large results arrive grouped by the top bits of their bucket in the table they are about to fill (power-of-two bucket
    // count, bucket = hash mod count, hash = the packed key), so every fill thread walks the bucket heads window by window
    // (128 KB each, cache resident) instead of at random over tens of megabytes -- and in pieces, so the fill threads work
    // on the first pieces while the rest is still crossing PCIe.
    template <typename Container, typename Make>
    static void fill_from_device(pk_ctx *ctx, const pk_dpoly *p, Container &c, Make make)
    {
        pk_result r{};
        const std::uint64_t n_dev = pk_dpoly_size(p);
        const unsigned nt = settings::get_n_threads();
        if (nt > 1u && n_dev >= (1u << 20) && n_dev < (std::uint64_t(1) << 32) && pk_ctx_device_count(ctx) == 1u) {
            unsigned b = 0;
            while ((std::uint64_t(1) << b) < n_dev) ++b; // the bucket count the fill will choose
            struct progress { // (the fill threads SLEEP while they wait: the thread that follows the copy needs a core to wake up on)
                std::mutex mx;
                std::condition_variable cv;
                std::uint64_t ready = 0;
                bool failed = false;
                const pk_result *r = nullptr;
            } pg;
            c.clear();
            std::exception_ptr fill_err;
            bool filled = false;
            // one piece of the fill per group: the groups own disjoint ranges of buckets, so the bucket heads need no atomics
            constexpr unsigned group_bits = 8u, n_groups = 1u << group_bits;
            std::vector<std::uint64_t> goff(n_groups + 1u, 0); // (written by the library before the first progress call)
            std::thread filler([&] {
                try {
                    filled = c._bulk_unique_fill_streamed(
                        n_dev, [&pg, &make](std::size_t i) { return make(*pg.r, i); },
                        [&pg](std::size_t i) { return static_cast<std::size_t>(pg.r->key[i]); }, nt, n_groups,
                        [&pg](std::size_t hi) {
                            std::unique_lock<std::mutex> lk(pg.mx);
                            pg.cv.wait(lk, [&] { return pg.ready >= hi || pg.failed; });
                            return pg.ready >= hi;
                        },
                        [&goff](unsigned g) { return std::make_pair(static_cast<std::size_t>(goff[g]), static_cast<std::size_t>(goff[g + 1u])); },
                        true);
                } catch (...) {
                    fill_err = std::current_exception();
                }
            });
            const int rc = pk_download_grouped_progress(
                ctx, p, b - group_bits, group_bits, 4u * nt,
                [](void *user, const pk_result *out, std::uint64_t ready) {
                    auto *g = static_cast<progress *>(user);
                    {
                        std::lock_guard<std::mutex> lk(g->mx);
                        g->r = out;
                        g->ready = ready;
                    }
                    g->cv.notify_all();
                },
                &pg, goff.data(), &r);
            {
                std::lock_guard<std::mutex> lk(pg.mx);
                if (rc != PK_OK || pg.ready < n_dev) pg.failed = true;
            }
            pg.cv.notify_all();
            filler.join();
            if (rc == PK_OK) pk_result_free(ctx, &r);
            if (rc != PK_OK || fill_err || !filled) c.clear();
            check(ctx, rc);
            if (fill_err) std::rethrow_exception(fill_err);
            if (!filled) throw std::runtime_error("polynomial: the fill of the result container was interrupted");
            return;
        }
        check(ctx, pk_download(ctx, p, &r));
        try {
            c.clear();
            if (nt > 1u && r.n >= (1u << 16)) { // large results: thread-parallel fill
                c._bulk_unique_fill(
                    r.n, [&r, &make](std::size_t i) { return make(r, i); }, [&r](std::size_t i) { return static_cast<std::size_t>(r.key[i]); },
                    nt);
            } else {
                c.rehash(std::max<std::size_t>(1, r.n));
                for (std::uint64_t i = 0; i < r.n; ++i) {
                    if (i + 24 < r.n) c._prefetch_bucket(c._bucket_from_hash(static_cast<std::size_t>(r.key[i + 24])));
                    c._unique_insert(make(r, i), c._bucket_from_hash(static_cast<std::size_t>(r.key[i])));
                }
                c._update_size(r.n);
            }
        } catch (...) {
            pk_result_free(ctx, &r);
            c.clear();
            throw;
        }
        pk_result_free(ctx, &r);
    }

private:
    static std::mutex &slot_mutex()
    {
        static std::mutex m;
        return m;
    }
    static handle &slot()
    {
        static handle h;
        return h;
    }
    static std::vector<int> device_list(unsigned want)
    {
        std::vector<int> ids;
        if (const char *list = std::getenv("PB200_DEVICES")) {
            const char *p = list;
            while (*p) {
                char *end = nullptr;
                const long v = std::strtol(p, &end, 10);
                if (end == p) break;
                ids.push_back(static_cast<int>(v));
                p = (*end == ',') ? end + 1 : end;
            }
        }
        if (ids.empty()) { // devices first, first + 1, ...; more parts than the box has GPUs share the devices round-robin
            const char *dev = std::getenv("PB200_DEVICE");
            const int first = dev ? std::atoi(dev) : 0;
            const int have = std::max(1, pk_device_count());
            for (unsigned i = 0; i < want; ++i) ids.push_back((first + static_cast<int>(i)) % have);
        } else {
            const std::size_t listed = ids.size();
            for (std::size_t i = listed; i < want; ++i) ids.push_back(ids[i % listed]);
            ids.resize(std::max<std::size_t>(1, want));
        }
        return ids;
    }
};

template <typename Cf, typename Key>
class polynomial
{
public:
    using term_type = term<Cf, Key>;
    using container_type = hash_set<term_type, term_hasher>;
    using size_type = typename container_type::size_type;
    using degree_type = typename Key::value_type;

    polynomial() = default;
    // pt{"x"}: the symbol x with unit coefficient
    explicit polynomial(const std::string &name) : m_symbol_set{name} { insert(term_type(Cf(1), Key{degree_type(1)})); }
    explicit polynomial(const char *name) : polynomial(std::string(name)) {}
    // pt{3}: a constant
    template <typename T, typename std::enable_if<std::is_integral<T>::value || std::is_same<T, Cf>::value, int>::type = 0>
    polynomial(const T &c)
    {
        insert(term_type(Cf(c), Key{}));
    }

    // number of terms; a device-resident product answers without touching the host
    size_type size() const
    {
        if (m_host_valid.load(std::memory_order_acquire)) return m_container.size();
        std::lock_guard<std::recursive_mutex> lk(m_lazy.mx);
        return m_host_valid.load(std::memory_order_relaxed) ? m_container.size() : static_cast<size_type>(pk_dpoly_size(m_dev->p));
    }
    bool empty() const { return size() == 0; }
    const symbol_fset &get_symbol_set() const { return m_symbol_set; }
    void set_symbol_set(const symbol_fset &s)
    {
        if (!empty()) throw std::invalid_argument("cannot set arguments on a non-empty series");
        m_symbol_set = s;
    }
    // the term container (series.hpp:2817-2828).  Reading it brings a device-resident result to the host (once);
    // the non-const form also drops the device copy, because the caller may change the terms.
    container_type &_container() { return host_mut(); }
    const container_type &_container() const { return host(); }
    // ---- device residency: a product stays in HBM until its terms are asked for, and
    // an operand keeps the device copy made for its first product, so chains like f *= tmp (series.hpp:799-805) and pow
    // never re-upload f nor download intermediate results
    bool is_device_resident() const
    {
        std::lock_guard<std::recursive_mutex> lk(m_lazy.mx);
        return static_cast<bool>(m_dev);
    }
    bool is_host_resident() const { return m_host_valid.load(std::memory_order_acquire); }
    void materialise() const { (void)host(); }

    // series::insert: add to an existing term or create it, drop zeros (series.hpp:2190)
    template <bool Sign = true>
    void insert(term_type t)
    {
        if (!t.is_compatible(m_symbol_set)) throw std::invalid_argument("cannot insert incompatible term");
        if (t.is_zero(m_symbol_set)) return;
        container_type &c = host_mut();
        auto it = c.find(t);
        if (it == c.end()) {
            if (!Sign) t.m_cf = -t.m_cf;
            c.insert(std::move(t));
        } else {
            if (Sign) it->m_cf += t.m_cf;
            else it->m_cf -= t.m_cf;
            if (math::is_zero(it->m_cf)) c.erase(it);
        }
    }

    // ---- arithmetic
    polynomial operator-() const
    {
        polynomial r(*this);
        for (const auto &t : r.host_mut()) t.m_cf = -t.m_cf;
        return r;
    }
    polynomial &operator+=(const polynomial &o) { return *this = merged_binary(*this, o, true); }
    polynomial &operator-=(const polynomial &o) { return *this = merged_binary(*this, o, false); }
    // x *= y  is  x = x * y  (series.hpp:799-805)
    polynomial &operator*=(const polynomial &o) { return *this = *this * o; }
    friend polynomial operator+(const polynomial &a, const polynomial &b) { return merged_binary(a, b, true); }
    friend polynomial operator-(const polynomial &a, const polynomial &b) { return merged_binary(a, b, false); }
    // dispatch_binary_mul -> series_merge_f -> series_multiplier<Series>(x, y)()   (series.hpp:738-752, :2855-2893)
    friend polynomial operator*(const polynomial &a, const polynomial &b)
    {
        if (a.m_symbol_set == b.m_symbol_set) return series_multiplier<polynomial>(a, b)();
        symbol_fset merged(a.m_symbol_set);
        merged.insert(b.m_symbol_set.begin(), b.m_symbol_set.end());
        const polynomial a2 = a.extend_symbol_set(merged), b2 = b.extend_symbol_set(merged);
        return series_multiplier<polynomial>(a2, b2)();
    }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator+(const polynomial &a, T n) { return a + polynomial(n); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator+(T n, const polynomial &a) { return polynomial(n) + a; }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator-(const polynomial &a, T n) { return a - polynomial(n); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator-(T n, const polynomial &a) { return polynomial(n) - a; }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator*(const polynomial &a, T n) { return a.scaled(Cf(n)); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend polynomial operator*(T n, const polynomial &a) { return a.scaled(Cf(n)); }

    friend bool operator==(const polynomial &a, const polynomial &b)
    {
        if (a.m_symbol_set != b.m_symbol_set) {
            symbol_fset merged(a.m_symbol_set);
            merged.insert(b.m_symbol_set.begin(), b.m_symbol_set.end());
            return a.extend_symbol_set(merged).same_terms(b.extend_symbol_set(merged));
        }
        return a.same_terms(b);
    }
    friend bool operator!=(const polynomial &a, const polynomial &b) { return !(a == b); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend bool operator==(const polynomial &a, T n) { return a == polynomial(n); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    friend bool operator==(T n, const polynomial &a) { return a == polynomial(n); }

    // Exponentiation: single-term series go through the key (so x.pow(-1) and x.pow(limit) work as in the reference,
    // series.hpp:2333-2403), everything else by repeated multiplication through operator*.
    polynomial pow(long long n) const
    {
        if (size() == 1u) {
            const term_type &t = *host().begin();
            if (n >= 0 || t.m_cf == Cf(1) || t.m_cf == Cf(-1)) {
                polynomial r;
                r.m_symbol_set = m_symbol_set;
                Cf c = n >= 0 ? t.m_cf.pow(static_cast<unsigned>(n)) : ((n % 2) ? t.m_cf : Cf(1));
                r.insert(term_type(std::move(c), t.m_key.pow(n, m_symbol_set)));
                return r;
            }
        }
        if (n < 0) throw std::invalid_argument("negative powers are supported only for single-term series with unit coefficient");
        polynomial r(1), b(*this);
        r.m_symbol_set = m_symbol_set;
        r.host_mut().clear();
        r.insert(term_type(Cf(1), Key(std::vector<degree_type>(m_symbol_set.size(), 0))));
        unsigned long long e = static_cast<unsigned long long>(n);
        while (e) {
            if (e & 1u) r = r * b;
            e >>= 1;
            if (e) b = b * b;
        }
        return r;
    }

    degree_type degree() const
    {
        if (empty()) return degree_type(0);
        bool first = true;
        degree_type d = 0;
        for (const auto &t : host()) {
            const degree_type td = t.m_key.degree(m_symbol_set);
            if (first || td > d) d = td;
            first = false;
        }
        return d;
    }
    // coefficient of the monomial with the given exponents (find_cf)
    Cf find_cf(const std::vector<degree_type> &exponents) const
    {
        if (exponents.size() != m_symbol_set.size()) throw std::invalid_argument("invalid number of exponents");
        const auto it = host().find(term_type(Cf(0), Key(exponents)));
        return it == host().end() ? Cf(0) : it->m_cf;
    }

    // ---- auto-truncation (polynomial.hpp:646-740): per-type static state under a mutex
    static void set_auto_truncate_degree(const degree_type &max_degree)
    {
        std::lock_guard<std::mutex> lock(s_at_mutex());
        s_at_mode() = 1;
        s_at_degree() = max_degree;
        s_at_names().clear();
    }
    static void set_auto_truncate_degree(const degree_type &max_degree, const std::vector<std::string> &names)
    {
        std::lock_guard<std::mutex> lock(s_at_mutex());
        s_at_mode() = 2;
        s_at_degree() = max_degree;
        s_at_names() = symbol_fset(names.begin(), names.end());
    }
    static void unset_auto_truncate_degree()
    {
        std::lock_guard<std::mutex> lock(s_at_mutex());
        s_at_mode() = 0;
        s_at_degree() = degree_type(0);
        s_at_names().clear();
    }
    static std::tuple<int, degree_type, symbol_fset> get_auto_truncate_degree()
    {
        std::lock_guard<std::mutex> lock(s_at_mutex());
        return std::make_tuple(s_at_mode(), s_at_degree(), s_at_names());
    }
    // explicit (non-automatic) forms, polynomial.hpp:813-893
    static polynomial untruncated_multiplication(const polynomial &p1, const polynomial &p2)
    {
        return with_merged(p1, p2, [](const polynomial &a, const polynomial &b) { return series_multiplier<polynomial>(a, b)._untruncated_multiplication(); });
    }
    static polynomial truncated_multiplication(const polynomial &p1, const polynomial &p2, const degree_type &max_degree)
    {
        return with_merged(p1, p2, [&](const polynomial &a, const polynomial &b) { return series_multiplier<polynomial>(a, b)._truncated_multiplication(max_degree); });
    }
    static polynomial truncated_multiplication(const polynomial &p1, const polynomial &p2, const degree_type &max_degree,
                                               const std::vector<std::string> &names)
    {
        return with_merged(p1, p2, [&](const polynomial &a, const polynomial &b) {
            std::vector<std::size_t> idx;
            for (const auto &n : symbol_fset(names.begin(), names.end())) {
                const std::size_t i = symbol_index(a.get_symbol_set(), n);
                if (i < a.get_symbol_set().size()) idx.push_back(i);
            }
            return series_multiplier<polynomial>(a, b)._truncated_multiplication(max_degree, names, idx);
        });
    }

    // canonical printing (sorted by key) for test diagnostics
    friend std::ostream &operator<<(std::ostream &os, const polynomial &p)
    {
        std::vector<const term_type *> v;
        for (const auto &t : p.host()) v.push_back(&t);
        std::sort(v.begin(), v.end(), [](const term_type *a, const term_type *b) { return a->m_key < b->m_key; });
        if (v.empty()) return os << "0";
        bool first = true;
        for (const term_type *t : v) {
            if (!first) os << " + ";
            first = false;
            os << t->m_cf;
            const auto e = t->m_key.unpack(p.m_symbol_set);
            std::size_t i = 0;
            for (const auto &n : p.m_symbol_set) {
                if (e[i]) os << "*" << n << "**" << e[i];
                ++i;
            }
        }
        return os;
    }

    polynomial extend_symbol_set(const symbol_fset &merged) const
    {
        if (merged == m_symbol_set) return *this;
        polynomial r;
        r.m_symbol_set = merged;
        r.m_container.rehash(host().bucket_count());
        for (const auto &t : host()) r.m_container.insert(term_type(t.m_cf, t.m_key.merge_symbols(m_symbol_set, merged)));
        return r;
    }

private:
    template <typename F>
    static polynomial with_merged(const polynomial &a, const polynomial &b, F &&f)
    {
        if (a.m_symbol_set == b.m_symbol_set) return f(a, b);
        symbol_fset merged(a.m_symbol_set);
        merged.insert(b.m_symbol_set.begin(), b.m_symbol_set.end());
        return f(a.extend_symbol_set(merged), b.extend_symbol_set(merged));
    }
    static polynomial merged_binary(const polynomial &a, const polynomial &b, bool plus)
    {
        symbol_fset merged(a.m_symbol_set);
        merged.insert(b.m_symbol_set.begin(), b.m_symbol_set.end());
        polynomial r = a.extend_symbol_set(merged);
        const polynomial b2 = b.extend_symbol_set(merged);
        for (const auto &t : b2.host()) {
            if (plus) r.template insert<true>(t);
            else r.template insert<false>(t);
        }
        return r;
    }
    polynomial scaled(const Cf &c) const
    {
        polynomial r;
        r.m_symbol_set = m_symbol_set;
        if (math::is_zero(c)) return r;
        r.m_container.rehash(host().bucket_count());
        for (const auto &t : host()) r.m_container.insert(term_type(t.m_cf * c, t.m_key));
        return r;
    }
    bool same_terms(const polynomial &o) const
    {
        if (size() != o.size()) return false;
        for (const auto &t : host()) {
            const auto it = o.host().find(t);
            if (it == o.host().end() || it->m_cf != t.m_cf) return false;
        }
        return true;
    }
    static std::mutex &s_at_mutex()
    {
        static std::mutex m;
        return m;
    }
    static int &s_at_mode()
    {
        static int v = 0;
        return v;
    }
    static degree_type &s_at_degree()
    {
        static degree_type v = 0;
        return v;
    }
    static symbol_fset &s_at_names()
    {
        static symbol_fset v;
        return v;
    }

    // ---- host / device state
    struct device_terms { // owns one pk_dpoly (immutable once created, so copies of the series share it) and keeps the
                          // context that created it alive; freed through that context, whichever thread drops the last copy
        gpu::handle h;
        pk_dpoly *p = nullptr;
        ~device_terms()
        {
            if (p && h) {
                std::lock_guard<std::recursive_mutex> lk(h->mx);
                pk_dpoly_free(h->ctx, p);
            }
        }
    };
    // The host container and the device copy are two lazily filled caches of the same terms, filled from const member
    // functions.  The reference allows concurrent const use of a series, so both fills are serialised per series.
    struct lazy_state {
        mutable std::recursive_mutex mx;
        lazy_state() = default;
        lazy_state(const lazy_state &) {}
        lazy_state &operator=(const lazy_state &) { return *this; }
    };
    static constexpr bool gpu_type = std::is_same<Cf, integer>::value && std::is_same<Key, kronecker_monomial<std::int64_t>>::value;
    const container_type &host() const
    {
        if (!m_host_valid.load(std::memory_order_acquire)) {
            std::lock_guard<std::recursive_mutex> lk(m_lazy.mx);
            if (!m_host_valid.load(std::memory_order_relaxed)) download();
        }
        return m_container;
    }
    container_type &host_mut()
    {
        host();
        std::lock_guard<std::recursive_mutex> lk(m_lazy.mx);
        m_dev.reset();
        return m_container;
    }
    // device terms -> host container: rehash + _unique_insert + _update_size, as sparse_kronecker_multiplication fills its
    // result (polynomial.hpp:1666-1667, :1750-1754, base_series_multiplier.hpp:934)
    void download() const
    {
        if constexpr (gpu_type) { // (called with m_lazy.mx held)
            const gpu::handle h = m_dev->h;
            std::unique_lock<std::recursive_mutex> ctx_lock(h->mx);
            pk_ctx *ctx = h->ctx;
            gpu::fill_from_device(ctx, m_dev->p, m_container, [](const pk_result &q, std::size_t i) {
                return term_type(integer::from_limbs(q.sign[i], q.mag + i * q.limbs, q.limbs), Key::from_int(q.key[i]));
            });
            m_host_valid.store(true, std::memory_order_release);
        }
    }
    template <typename, typename>
    friend class series_multiplier;

    symbol_fset m_symbol_set;
    mutable container_type m_container;
    mutable copyable_atomic_bool m_host_valid{true};
    mutable std::shared_ptr<device_terms> m_dev;
    lazy_state m_lazy;
};

namespace math
{
template <typename Cf, typename Key>
inline polynomial<Cf, Key> pow(const polynomial<Cf, Key> &p, long long n)
{
    return p.pow(n);
}
} // namespace math

// ------------------------------------------------------------------------------------------------ the drop-in

// series_multiplier<polynomial<integer, kronecker_monomial<int64_t>>>: same protocol and member names as
// polynomial.hpp:991-1968, with the work done by the C ABI.  Operands are brought to the device once (the copy stays
// with the series), the product is pk_mul_dev and stays on the device until the caller reads its terms.
template <>
class series_multiplier<polynomial<integer, kronecker_monomial<std::int64_t>>, void>
{
public:
    using series_type = polynomial<integer, kronecker_monomial<std::int64_t>>;
    using key_type = kronecker_monomial<std::int64_t>;
    using term_type = series_type::term_type;
    using degree_type = series_type::degree_type;

    // base_series_multiplier ctor (base_series_multiplier.hpp:363-401) + check_bounds (polynomial.hpp:1118-1194):
    // throws std::invalid_argument on mismatching symbol sets and std::overflow_error when an exponent of the
    // product could leave the Kronecker limits -- from the constructor, like the reference.
    explicit series_multiplier(const series_type &s1, const series_type &s2) : m_ss(s1.get_symbol_set())
    {
        if (s1.get_symbol_set() != s2.get_symbol_set()) throw std::invalid_argument("incompatible arguments sets");
        m_s1 = &s1;
        m_s2 = &s2;
        if (m_s1->size() < m_s2->size()) std::swap(m_s1, m_s2); // larger series first
        if (m_s1->empty() || m_s2->empty()) return;
        m_h = gpu::shared();
        std::lock_guard<std::recursive_mutex> lk(m_h->mx);
        gpu::check(m_h->ctx, pk_check_bounds(m_h->ctx, device(*m_s1, m_h), device(*m_s2, m_h)));
    }
    series_multiplier(const series_multiplier &) = delete;
    series_multiplier &operator=(const series_multiplier &) = delete;

    // execute(): the auto-truncation state of the series type selects the path (polynomial.hpp:1596-1628)
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
        pk_opts o = make_opts();
        return run(o);
    }
    series_type _truncated_multiplication(const degree_type &max_degree) const
    {
        pk_opts o = make_opts();
        o.trunc_mode = 1;
        o.max_degree = max_degree;
        return run(o);
    }
    series_type _truncated_multiplication(const degree_type &max_degree, const std::vector<std::string> &,
                                          const std::vector<std::size_t> &idx) const
    {
        std::vector<std::uint8_t> mask(std::max<std::size_t>(1, m_ss.size()), 0);
        for (const std::size_t i : idx) {
            if (i >= m_ss.size()) throw std::invalid_argument("invalid symbol position in a truncated multiplication");
            mask[i] = 1;
        }
        pk_opts o = make_opts();
        o.trunc_mode = 2;
        o.max_degree = max_degree;
        o.degree_mask = mask.data();
        return run(o);
    }
    // statistics of the last product on the process' context (benchmarks)
    static pk_stats last_stats()
    {
        pk_stats st{};
        const gpu::handle h = gpu::shared();
        std::lock_guard<std::recursive_mutex> lk(h->mx);
        pk_get_stats(h->ctx, &st);
        return st;
    }

private:
    static pk_opts make_opts()
    {
        pk_opts o{};
        o.struct_size = sizeof(pk_opts);
        return o;
    }
    // The series' terms in HBM: uploaded on first use (the role of fill_term_pointers, base_series_multiplier.hpp:89-154:
    // walk the container once, here into the structure-of-arrays form of the C ABI) and kept with the series.
    static pk_dpoly *device(const series_type &s, const gpu::handle &h)
    {
        std::lock_guard<std::recursive_mutex> lazy(s.m_lazy.mx);
        if (s.m_dev && s.m_dev->h == h) return s.m_dev->p;
        // (a device copy that belongs to an earlier context -- settings::set_n_gpus() changed -- is read back first)
        const auto &c = s.host();
        const std::size_t n = c.size();
        std::size_t L = 1;
        for (const auto &t : c) L = std::max(L, t.m_cf.size());
        // wider operands than the C ABI takes are reported by pk_upload as PK_E_CF_WIDTH -> std::overflow_error
        std::vector<std::int64_t> key(n);
        std::vector<std::uint64_t> mag(n * L, 0);
        std::vector<std::int8_t> sign(n);
        std::size_t i = 0;
        for (const auto &t : c) {
            key[i] = t.m_key.get_int();
            for (std::size_t l = 0; l < t.m_cf.size(); ++l) mag[i * L + l] = t.m_cf.limb(l);
            sign[i] = static_cast<std::int8_t>(t.m_cf.sign());
            ++i;
        }
        const pk_poly hp{n, static_cast<std::uint32_t>(s.get_symbol_set().size()), static_cast<std::uint32_t>(L), key.data(),
                         mag.data(), sign.data()};
        auto dev = std::make_shared<series_type::device_terms>();
        dev->h = h;
        std::lock_guard<std::recursive_mutex> lk(h->mx);
        gpu::check(h->ctx, pk_upload(h->ctx, &hp, &dev->p));
        s.m_dev = std::move(dev);
        return s.m_dev->p;
    }
    series_type run(const pk_opts &o) const
    {
        series_type retval;
        retval.set_symbol_set(m_ss); // polynomial.hpp:1651-1652
        if (m_s1->empty() || m_s2->empty()) return retval; // :1654-1656
        auto dev = std::make_shared<series_type::device_terms>();
        dev->h = m_h;
        {
            std::lock_guard<std::recursive_mutex> lk(m_h->mx);
            gpu::check(m_h->ctx, pk_mul_dev(m_h->ctx, device(*m_s1, m_h), device(*m_s2, m_h), &o, &dev->p));
        }
        // the result stays in HBM; reading its terms (or _container()) downloads it once into the hash_set
        retval.m_dev = std::move(dev);
        retval.m_host_valid.store(false, std::memory_order_release);
        return retval;
    }

    symbol_fset m_ss;
    const series_type *m_s1 = nullptr, *m_s2 = nullptr;
    gpu::handle m_h;
};

} // namespace pb200

#endif
