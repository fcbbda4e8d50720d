// Host container of the drop-in: the subset of piranha::hash_set (reference: include/piranha/hash_set.hpp) that the
// multiplier path touches -- the unchecked low-level interface through which a series_multiplier fills its result
// (rehash :1125-1160, _bucket :1298-1317, _find :1274-1289, _unique_insert :1251-1260, _update_size :1324-1327) plus
// the ordinary interface series arithmetic needs (find / insert / erase / iteration).
//
// Own design, not the reference's inline-first-node lists: nodes live in one contiguous pool and buckets hold pool
// indices, so iteration is a linear sweep and filling from the device result is one append per term.  Same
// observable contract: power-of-two bucket count, maximum load factor 1, bucket = hash mod bucket_count.
#ifndef PB200_HASH_SET_HPP
#define PB200_HASH_SET_HPP

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <functional>
#include <iterator>
#include <mutex>
#include <new>
#include <stdexcept>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace pb200
{

namespace detail
{
// Contiguous pool of constructed nodes over raw storage: what hash_set needs from std::vector, plus the one thing
// std::vector cannot do -- hand out uninitialised storage for n nodes so that several threads construct (and first-touch
// the pages of) disjoint ranges at once.
// Large arrays (node pools and bucket heads of results with millions of terms) are 2 MB aligned and advised to use huge
// pages: a bulk fill touches the bucket heads in random order, one TLB entry per element with 4 KB pages.
// Freed large blocks are kept (a few, up to PB200_HOST_CACHE_MB megabytes in total, default 1024; 0 = keep nothing) and
// handed to the next request of about the same size: a benchmark loop or a chain of products builds results of the same
// size over and over, and fresh pages cost a fault and a kernel zero-fill each (a quarter of the fill of a 5.8 M term
// result).  glibc does the same below its mmap threshold; these blocks are above it.
struct big_block_cache {
    static constexpr std::size_t slots = 8;
    struct block {
        void *p;
        std::size_t bytes;
    };
    std::mutex m;
    block kept[slots] = {};
    std::size_t total = 0, limit;
    big_block_cache()
    {
        const char *e = std::getenv("PB200_HOST_CACHE_MB");
        limit = (e ? std::strtoull(e, nullptr, 10) : 1024ull) << 20;
    }
    void *take(std::size_t bytes)
    {
        std::lock_guard<std::mutex> g(m);
        block *best = nullptr;
        for (auto &b : kept)
            if (b.p && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 && (!best || b.bytes < best->bytes)) best = &b;
        if (!best) return nullptr;
        void *p = best->p;
        total -= best->bytes;
        *best = block{nullptr, 0};
        return p;
    }
    bool keep(void *p, std::size_t bytes)
    {
        std::lock_guard<std::mutex> g(m);
        if (total + bytes > limit) return false;
        for (auto &b : kept)
            if (!b.p) {
                b = block{p, bytes};
                total += bytes;
                return true;
            }
        return false;
    }
    static big_block_cache &get()
    {
        static big_block_cache *c = new big_block_cache; // (never destroyed: containers with static storage duration may free after main)
        return *c;
    }
};
constexpr std::size_t big_block_min = std::size_t(4) << 20;
inline void *big_alloc(std::size_t bytes)
{
    constexpr std::size_t huge = std::size_t(2) << 20;
    if (bytes >= big_block_min) {
        const std::size_t rounded = (bytes + huge - 1) & ~(huge - 1);
        if (void *p = big_block_cache::get().take(rounded)) return p;
        // (the size is kept in front of the block: one huge page of slack keeps the block itself 2 MB aligned)
        void *raw = nullptr;
        if (::posix_memalign(&raw, huge, rounded + huge) != 0 || !raw) throw std::bad_alloc();
#if defined(__linux__) && defined(MADV_HUGEPAGE)
        ::madvise(raw, rounded + huge, MADV_HUGEPAGE);
#endif
        char *p = static_cast<char *>(raw) + huge;
        reinterpret_cast<std::size_t *>(p)[-1] = rounded;
        return p;
    }
    void *p = std::malloc(bytes ? bytes : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
// bytes: the size the block was requested with
inline void big_free(void *p, std::size_t bytes)
{
    if (!p) return;
    if (bytes >= big_block_min) {
        const std::size_t rounded = reinterpret_cast<std::size_t *>(p)[-1];
        if (big_block_cache::get().keep(p, rounded)) return;
        std::free(static_cast<char *>(p) - (std::size_t(2) << 20));
        return;
    }
    std::free(p);
}

// std::allocator that leaves trivially constructible elements uninitialised on resize() (the caller fills them, possibly
// from several threads) and allocates through big_alloc
template <typename T>
struct raw_big_allocator {
    using value_type = T;
    raw_big_allocator() = default;
    template <typename U>
    raw_big_allocator(const raw_big_allocator<U> &) noexcept
    {
    }
    T *allocate(std::size_t n) { return static_cast<T *>(big_alloc(n * sizeof(T))); }
    void deallocate(T *p, std::size_t n) noexcept { big_free(p, n * sizeof(T)); }
    template <typename U>
    void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value)
    {
        ::new (static_cast<void *>(p)) U; // default-initialisation: no zero fill
    }
    template <typename U, typename... Args>
    void construct(U *p, Args &&...args)
    {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
    template <typename U>
    bool operator==(const raw_big_allocator<U> &) const noexcept
    {
        return true;
    }
    template <typename U>
    bool operator!=(const raw_big_allocator<U> &) const noexcept
    {
        return false;
    }
};

template <typename Node>
class node_pool
{
public:
    node_pool() = default;
    node_pool(const node_pool &o)
    {
        reserve(o.m_size);
        for (; m_size < o.m_size; ++m_size) ::new (static_cast<void *>(m_p + m_size)) Node(o.m_p[m_size]);
    }
    node_pool(node_pool &&o) noexcept : m_p(o.m_p), m_size(o.m_size), m_cap(o.m_cap) { o.m_p = nullptr, o.m_size = o.m_cap = 0; }
    node_pool &operator=(node_pool o) noexcept
    {
        swap(o);
        return *this;
    }
    ~node_pool() { release(); }
    void swap(node_pool &o) noexcept
    {
        std::swap(m_p, o.m_p);
        std::swap(m_size, o.m_size);
        std::swap(m_cap, o.m_cap);
    }
    std::size_t size() const { return m_size; }
    Node &operator[](std::size_t i) { return m_p[i]; }
    const Node &operator[](std::size_t i) const { return m_p[i]; }
    void clear() { release(); }
    void reserve(std::size_t n)
    {
        if (n <= m_cap) return;
        Node *np = static_cast<Node *>(big_alloc(n * sizeof(Node)));
        for (std::size_t i = 0; i < m_size; ++i) {
            ::new (static_cast<void *>(np + i)) Node(std::move(m_p[i]));
            m_p[i].~Node();
        }
        big_free(m_p, m_cap * sizeof(Node));
        m_p = np;
        m_cap = n;
    }
    void push_back(Node &&v)
    {
        if (m_size == m_cap) reserve(m_cap ? 2 * m_cap : 8);
        ::new (static_cast<void *>(m_p + m_size)) Node(std::move(v));
        ++m_size;
    }
    // raw storage for n nodes in an EMPTY pool; the caller constructs nodes and then calls commit(n) (all constructed) or
    // abandon() (none left constructed)
    Node *raw(std::size_t n)
    {
        release();
        reserve(n);
        return m_p;
    }
    void commit(std::size_t n) { m_size = n; }
    void abandon() { m_size = 0; }

private:
    void release()
    {
        if (!std::is_trivially_destructible<Node>::value && m_size) {
            // (a coefficient may own heap storage: every node is visited, by several threads when the pool is large --
            // 5.8 M nodes take 19 ms on one thread, the time of the fill itself)
            const std::size_t hw = std::max(1u, std::thread::hardware_concurrency());
            const std::size_t nt = m_size >= (std::size_t(1) << 20) ? std::min<std::size_t>({16, hw, m_size >> 18}) : 1;
            Node *p = m_p;
            const std::size_t n = m_size;
            auto destroy = [p, n, nt](std::size_t t) {
                for (std::size_t i = n * t / nt, hi = n * (t + 1) / nt; i < hi; ++i) p[i].~Node();
            };
            std::vector<std::thread> th;
            try {
                for (std::size_t t = 1; t < nt; ++t) th.emplace_back(destroy, t);
            } catch (...) { // (no more threads: the rest is done here)
                for (std::size_t t = th.size() + 1; t < nt; ++t) destroy(t);
            }
            destroy(0);
            for (auto &x : th) x.join();
        }
        big_free(m_p, m_cap * sizeof(Node));
        m_p = nullptr;
        m_size = m_cap = 0;
    }
    Node *m_p = nullptr;
    std::size_t m_size = 0, m_cap = 0;
};
} // namespace detail

template <typename T, typename Hash = std::hash<T>, typename Pred = std::equal_to<T>>
class hash_set
{
    static constexpr std::uint32_t npos = 0xFFFFFFFFu;
    struct node {
        T value;
        std::uint32_t next;
        bool live;
    };

public:
    using key_type = T;
    using size_type = std::size_t;

    template <bool Const>
    class iter
    {
        friend class hash_set;
        using set_ptr = typename std::conditional<Const, const hash_set *, hash_set *>::type;
        set_ptr m_set = nullptr;
        std::size_t m_idx = 0;
        void skip()
        {
            while (m_idx < m_set->m_nodes.size() && !m_set->m_nodes[m_idx].live) ++m_idx;
        }

    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = T;
        using difference_type = std::ptrdiff_t;
        using pointer = const T *;
        using reference = const T &;
        iter() = default;
        iter(set_ptr s, std::size_t i) : m_set(s), m_idx(i) { skip(); }
        template <bool C = Const, typename std::enable_if<C, int>::type = 0>
        iter(const iter<false> &o) : m_set(o.m_set), m_idx(o.m_idx)
        {
        }
        reference operator*() const { return m_set->m_nodes[m_idx].value; }
        pointer operator->() const { return &m_set->m_nodes[m_idx].value; }
        iter &operator++()
        {
            ++m_idx;
            skip();
            return *this;
        }
        iter operator++(int)
        {
            iter t(*this);
            ++*this;
            return t;
        }
        friend bool operator==(const iter &a, const iter &b) { return a.m_idx == b.m_idx; }
        friend bool operator!=(const iter &a, const iter &b) { return a.m_idx != b.m_idx; }
    };
    using iterator = iter<false>;
    using const_iterator = iter<true>;

    hash_set() = default;
    explicit hash_set(size_type n_buckets) { rehash(n_buckets); }

    iterator begin() { return iterator(this, 0); }
    iterator end() { return iterator(this, m_nodes.size()); }
    const_iterator begin() const { return const_iterator(this, 0); }
    const_iterator end() const { return const_iterator(this, m_nodes.size()); }
    size_type size() const { return m_n_elements; }
    bool empty() const { return m_n_elements == 0; }
    size_type bucket_count() const { return m_heads.size(); }
    double load_factor() const { return m_heads.empty() ? 0. : static_cast<double>(m_n_elements) / static_cast<double>(m_heads.size()); }
    double max_load_factor() const { return 1.; }

    void clear()
    {
        m_nodes.clear();
        m_heads.clear();
        m_free = npos;
        m_n_elements = 0;
    }
    void swap(hash_set &o)
    {
        m_nodes.swap(o.m_nodes);
        m_heads.swap(o.m_heads);
        std::swap(m_free, o.m_free);
        std::swap(m_n_elements, o.m_n_elements);
    }

    // At least new_size buckets (rounded up to a power of two); refuses to exceed the maximum load factor, in which
    // case nothing changes -- the contract of hash_set::rehash.  The thread-count argument of the reference
    // (parallel bucket initialisation) is accepted and ignored.
    void rehash(size_type new_size, unsigned = 1u)
    {
        size_type nb = 1;
        while (nb < new_size) nb <<= 1;
        if (new_size == 0) {
            if (m_n_elements == 0) clear();
            return;
        }
        if (static_cast<double>(m_n_elements) / static_cast<double>(nb) > max_load_factor()) return;
        m_nodes.reserve(nb < (size_type(1) << 31) ? std::max(m_nodes.size(), std::min<size_type>(nb, new_size)) : m_nodes.size());
        m_heads.assign(nb, npos);
        for (std::uint32_t i = 0; i < m_nodes.size(); ++i) {
            if (!m_nodes[i].live) continue;
            const size_type b = _bucket(m_nodes[i].value);
            m_nodes[i].next = m_heads[b];
            m_heads[b] = i;
        }
    }

    const_iterator find(const T &k) const
    {
        if (m_heads.empty()) return end();
        return _find(k, _bucket(k));
    }
    iterator find(const T &k)
    {
        if (m_heads.empty()) return end();
        const const_iterator it = static_cast<const hash_set *>(this)->_find(k, _bucket(k));
        return iterator(this, it.m_idx);
    }
    std::pair<iterator, bool> insert(T k)
    {
        if (m_heads.empty()) rehash(1);
        size_type b = _bucket(k);
        const const_iterator it = static_cast<const hash_set *>(this)->_find(k, b);
        if (it != end()) return {iterator(this, it.m_idx), false};
        if (m_n_elements + 1 > m_heads.size()) { // keep the load factor <= 1
            rehash(m_heads.size() * 2);
            b = _bucket(k);
        }
        iterator r = _unique_insert(std::move(k), b);
        ++m_n_elements;
        return {r, true};
    }
    iterator erase(const_iterator it)
    {
        const std::uint32_t idx = static_cast<std::uint32_t>(it.m_idx);
        const size_type b = _bucket(m_nodes[idx].value);
        std::uint32_t *link = &m_heads[b];
        while (*link != idx) link = &m_nodes[*link].next;
        *link = m_nodes[idx].next;
        m_nodes[idx].live = false;
        m_nodes[idx].next = m_free;
        m_free = idx;
        --m_n_elements;
        return iterator(this, idx + 1);
    }

    // ---- low-level interface (no checks, no size bookkeeping: the caller finishes with _update_size)
    size_type _bucket_from_hash(std::size_t h) const { return h & (m_heads.size() - 1); }
    size_type _bucket(const T &k) const { return _bucket_from_hash(Hash()(k)); }
    const_iterator _find(const T &k, size_type bucket) const
    {
        for (std::uint32_t i = m_heads[bucket]; i != npos; i = m_nodes[i].next)
            if (Pred()(m_nodes[i].value, k)) return const_iterator(this, i);
        return end();
    }
    template <typename U>
    iterator _unique_insert(U &&k, size_type bucket)
    {
        std::uint32_t idx;
        if (m_free != npos) {
            idx = m_free;
            m_free = m_nodes[idx].next;
            m_nodes[idx].value = std::forward<U>(k);
            m_nodes[idx].live = true;
        } else {
            if (m_nodes.size() >= npos) throw std::overflow_error("hash_set: too many elements");
            idx = static_cast<std::uint32_t>(m_nodes.size());
            m_nodes.push_back(node{std::forward<U>(k), npos, true});
        }
        m_nodes[idx].next = m_heads[bucket];
        m_heads[bucket] = idx;
        return iterator(this, idx);
    }
    void _update_size(size_type n) { m_n_elements = n; }
    // Replace the contents with n elements known to be distinct: make(i) builds element i and hash_of(i) is its hash.
    // n_threads threads construct disjoint node ranges in place and push them on their buckets with an atomic exchange
    // of the bucket head (insert-only, no readers: lock-free and exact).  The thread-parallel counterpart of
    // rehash + n x _unique_insert + _update_size; the reference fills results from several threads by bucket ranges
    // (base_series_multiplier.hpp:885-948).  Node order is i, so iteration order does not depend on the thread count.
    template <typename Make, typename HashOf>
    void _bulk_unique_fill(size_type n, Make make, HashOf hash_of, unsigned n_threads)
    {
        if (n >= npos) throw std::overflow_error("hash_set: too many elements");
        clear();
        if (!n) return;
        size_type nb = 1;
        while (nb < n) nb <<= 1;
        n_threads = static_cast<unsigned>(std::max<size_type>(1, std::min<size_type>(n_threads, n / 4096 + 1)));
        node *pool = m_nodes.raw(n);
        m_heads.resize(nb); // (left uninitialised by the allocator: the slices are set to npos by the threads below)
        std::uint32_t *heads = m_heads.data();
        std::vector<size_type> built(n_threads, 0);
        std::vector<std::exception_ptr> err(n_threads);
        auto clear_heads = [&](unsigned t) {
            const size_type lo = nb * t / n_threads, hi = nb * (t + 1) / n_threads;
            std::fill(heads + lo, heads + hi, npos);
        };
        auto build = [&](unsigned t) {
            const size_type lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
            size_type i = lo;
            try {
                for (; i < hi; ++i) {
                    if (i + 16 < hi) __builtin_prefetch(heads + (hash_of(i + 16) & (nb - 1)), 1, 1);
                    ::new (static_cast<void *>(pool + i)) node{make(i), npos, true};
                    pool[i].next = __atomic_exchange_n(heads + (hash_of(i) & (nb - 1)), static_cast<std::uint32_t>(i), __ATOMIC_RELAXED);
                }
            } catch (...) {
                err[t] = std::current_exception();
            }
            built[t] = i - lo;
        };
        auto run = [&](auto &&f) {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(f, t);
            f(0u);
            for (auto &x : th) x.join();
        };
        run(clear_heads);
        run(build);
        for (unsigned t = 0; t < n_threads; ++t) {
            if (!err[t]) continue;
            for (unsigned u = 0; u < n_threads; ++u) {
                const size_type lo = n * u / n_threads;
                for (size_type i = lo; i < lo + built[u]; ++i) pool[i].~node();
            }
            m_nodes.abandon();
            clear();
            std::rethrow_exception(err[t]);
        }
        m_nodes.commit(n);
        m_n_elements = n;
    }
    // _bulk_unique_fill for elements that are still ARRIVING (a result crossing PCIe in pieces): the index range is cut in
    // n_chunks pieces dealt round robin over the threads, and a thread starts on a piece when ready_upto(hi) says that the
    // elements [0, hi) are there (it blocks until then; false = the producer gave up, and the container is left empty).
    // Returns false when it gave up.
    template <typename Make, typename HashOf, typename Ready>
    bool _bulk_unique_fill_streamed(size_type n, Make make, HashOf hash_of, unsigned n_threads, unsigned n_chunks, Ready ready_upto)
    {
        n_threads = static_cast<unsigned>(std::max<size_type>(1, std::min<size_type>(n_threads, n / 4096 + 1)));
        n_chunks = static_cast<unsigned>(std::max<size_type>(n_threads, std::min<size_type>(n_chunks, std::max<size_type>(n, 1))));
        return _bulk_unique_fill_streamed(
            n, make, hash_of, n_threads, n_chunks, ready_upto,
            [n, n_chunks](unsigned c) { return std::make_pair(n * c / n_chunks, n * (c + 1) / n_chunks); }, false);
    }
    // The general form: piece c is the elements [range_of(c).first, range_of(c).second) -- the pieces tile [0, n), in
    // order; range_of is first called after ready_upto(1) returned true (the producer may deliver the ranges with the
    // first elements).  exclusive_buckets: the caller guarantees that no two pieces share a bucket (elements grouped by
    // the top bits of their bucket, one piece per group): bucket heads are then updated with plain stores.
    template <typename Make, typename HashOf, typename Ready, typename Range>
    bool _bulk_unique_fill_streamed(size_type n, Make make, HashOf hash_of, unsigned n_threads, unsigned n_chunks, Ready ready_upto,
                                    Range range_of, bool exclusive_buckets)
    {
        if (n >= npos) throw std::overflow_error("hash_set: too many elements");
        clear();
        if (!n) return true;
        size_type nb = 1;
        while (nb < n) nb <<= 1;
        n_threads = std::max(1u, std::min(n_threads, n_chunks));
        node *pool = m_nodes.raw(n);
        m_heads.resize(nb);
        std::uint32_t *heads = m_heads.data();
        std::vector<std::pair<size_type, size_type>> built(n_chunks, {0, 0}); // {first node, nodes constructed} per piece
        std::vector<std::exception_ptr> err(n_threads);
        std::atomic<bool> stop{false};
        auto clear_heads = [&](unsigned t) {
            const size_type lo = nb * t / n_threads, hi = nb * (t + 1) / n_threads;
            std::fill(heads + lo, heads + hi, npos);
        };
        auto build = [&](unsigned t) {
            unsigned c = t;
            size_type lo = 0, i = 0; // (the count of a piece is stored once, at its end: the counters of the pieces share cache lines)
            try {
                if (!ready_upto(1)) {
                    stop.store(true, std::memory_order_relaxed);
                    return;
                }
                for (; c < n_chunks && !stop.load(std::memory_order_relaxed); c += n_threads) {
                    const auto range = range_of(c);
                    lo = i = range.first;
                    const size_type hi = range.second;
                    if (hi <= lo) continue;
                    if (hi > n || !ready_upto(hi)) {
                        stop.store(true, std::memory_order_relaxed);
                        return;
                    }
                    if (exclusive_buckets) {
                        for (; i < hi; ++i) {
                            if (i + 16 < hi) __builtin_prefetch(heads + (hash_of(i + 16) & (nb - 1)), 1, 1);
                            std::uint32_t &head = heads[hash_of(i) & (nb - 1)];
                            ::new (static_cast<void *>(pool + i)) node{make(i), head, true};
                            head = static_cast<std::uint32_t>(i);
                        }
                    } else {
                        for (; i < hi; ++i) {
                            if (i + 16 < hi) __builtin_prefetch(heads + (hash_of(i + 16) & (nb - 1)), 1, 1);
                            ::new (static_cast<void *>(pool + i)) node{make(i), npos, true};
                            pool[i].next = __atomic_exchange_n(heads + (hash_of(i) & (nb - 1)), static_cast<std::uint32_t>(i), __ATOMIC_RELAXED);
                        }
                    }
                    built[c] = {lo, hi - lo};
                }
            } catch (...) {
                if (c < n_chunks) built[c] = {lo, i - lo};
                err[t] = std::current_exception();
                stop.store(true, std::memory_order_relaxed);
            }
        };
        auto run = [&](auto &&f) {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(f, t);
            f(0u);
            for (auto &x : th) x.join();
        };
        run(clear_heads);
        run(build);
        size_type total = 0;
        for (unsigned c = 0; c < n_chunks; ++c) total += built[c].second;
        if (stop.load() || total != n) { // (total != n: the pieces did not tile [0, n))
            for (unsigned c = 0; c < n_chunks; ++c)
                for (size_type i = built[c].first; i < built[c].first + built[c].second; ++i) pool[i].~node();
            m_nodes.abandon();
            clear();
            for (unsigned t = 0; t < n_threads; ++t)
                if (err[t]) std::rethrow_exception(err[t]);
            return false;
        }
        m_nodes.commit(n);
        m_n_elements = n;
        return true;
    }
    // hint that `bucket` is about to receive an element (bulk fills touch the buckets in random order: one cache miss
    // per element unless the bucket heads are requested a few elements ahead)
    void _prefetch_bucket(size_type bucket) const { __builtin_prefetch(&m_heads[bucket], 1, 1); }

private:
    detail::node_pool<node> m_nodes;
    std::vector<std::uint32_t, detail::raw_big_allocator<std::uint32_t>> m_heads;
    std::uint32_t m_free = npos;
    size_type m_n_elements = 0;
};

} // namespace pb200

#endif
