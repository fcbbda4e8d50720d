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
#include <cstddef>
#include <cstdint>
#include <functional>
#include <iterator>
#include <stdexcept>
#include <utility>
#include <vector>

namespace pb200
{

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
    // hint that `bucket` is about to receive an element (bulk fills touch the buckets in random order: one cache miss
    // per element unless the bucket heads are requested a few elements ahead)
    void _prefetch_bucket(size_type bucket) const { __builtin_prefetch(&m_heads[bucket], 1, 1); }

private:
    std::vector<node> m_nodes;
    std::vector<std::uint32_t> m_heads;
    std::uint32_t m_free = npos;
    size_type m_n_elements = 0;
};

} // namespace pb200

#endif
