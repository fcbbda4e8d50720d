/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of piranha's sparse polynomial multiplier.
 *
 * Nothing under piranha_b200/ links, loads or calls this file.  It is the checker for tests/, for
 * __graft_entry__.smoke() and the `cpu_baseline` / `--impl reference` legs of bench.py.
 *
 * The reference itself cannot be compiled in this image (it needs Boost >= 1.58, mp++ >= 0.9 and gmp.h:
 * /root/reference/CMakeLists.txt:163-164, :196-198; none of the three is installed), so this is a labelled
 * restatement, not piranha.  PARITY PINNING: checked against every known answer the reference's own tests
 * and benchmarks hold for the path (tests/test_oracle_golden.py) and against the independent Python
 * big-integer oracle (oracle/pyoracle.py) coefficient by coefficient.
 *
 * What it follows (paths relative to /root/reference/include/piranha/):
 *   - base_series_multiplier ctor: larger operand first, thread count = use_threads(n1*n2, 250000)
 *       base_series_multiplier.hpp:363-401, thread_pool.hpp:463-490, settings.hpp:57
 *   - check_bounds (Kronecker): unpack every term, per-variable min/max, sums inside +-limits
 *       polynomial.hpp:1118-1194, kronecker_array.hpp:329-365
 *   - untruncated_kronecker_mult: estimate -> rehash(2^k >= estimate) -> sparse_kronecker_multiplication
 *       polynomial.hpp:1632-1672
 *   - estimate_final_series_size: 15 trials, multiplier 2, mt19937 seeded with the trial index
 *       base_series_multiplier.hpp:563-718   (sizes the table only; never changes the mathematical result)
 *   - sparse_kronecker_multiplication: operands stable-sorted by destination bucket, tasks of <= 256
 *       products, n_threads*10 zones of buckets, binary-searched (i, j0, j1) tasks, two batches for the
 *       modular wrap-around, threads claim zones with test_and_set          polynomial.hpp:1673-1967
 *   - task_consume: key add, bucket = key mod 2^k, chain probe, mul3 on miss / multiply_accumulate on hit
 *       polynomial.hpp:1723-1759, hash_set.hpp:1251-1327
 *   - truncated path: degrees, operand 2 stable-sorted by degree, skip limits by upper_bound, plain
 *       multiplication in 256x256 blocks with one spinlock per bucket when threaded
 *       polynomial.hpp:1402-1531, base_series_multiplier.hpp:449-504, :987-1102
 *   - sanitise_series: recount and drop zero coefficients            base_series_multiplier.hpp:855-949
 * Coefficients: mp++ integers never overflow (static storage of N limbs, promotion beyond); here they are fixed-width
 * two's complement accumulators of PC_L = 2, 3 or 4 64-bit limbs fed by operands of at most two limbs (sign + magnitude,
 * little-endian limbs).  This file is compiled once per width (oracle/Makefile: -DPC_L=2/3/4 -Dpc_mul=pc_mul_L<n>) and
 * oracle/pc_dispatch.c picks the narrowest width the bound |a|max * |b|max * min(n1, n2) fits, the way mppp::integer<N>
 * stays in its static limbs: a table node is {key, PC_L limbs, next} = 32 / 40 / 48 bytes (the reference's node is
 * term + next pointer).  Exact for every configuration in BASELINE.json (largest result coefficient 2^127.96 -> 3 limbs).
 * The bucket array is mapped with MADV_HUGEPAGE (the products of a task land ~2700 buckets apart, one TLB entry each with
 * 4 KB pages), first touched by the worker threads like the reference's parallel rehash (hash_set.hpp:463-531); worker
 * threads live in a persistent pool (thread_pool.hpp:75-215) and the estimator's trials run on it in parallel
 * (base_series_multiplier.hpp:592-718).
 */
#define _GNU_SOURCE
#include <sys/mman.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sched.h>
#include <unistd.h>

#define PC_OK 0
#define PC_E_ARGS 1
#define PC_E_KEY_RANGE 2
#define PC_E_NOMEM 4

#ifndef PC_L
#define PC_L 4
#endif
#define PC_ACC_LIMBS PC_L
#define PC_BLOCK 256u           /* tuning::multiplication_block_size, tuning.hpp:57 */
#define PC_EST_THRESHOLD 200u   /* tuning::estimate_threshold, tuning.hpp:60 */
#define PC_MIN_WORK 250000ull   /* settings::min_work_per_thread, settings.hpp:57 */
#define PC_ZM 10u               /* zones per thread, polynomial.hpp:1788 */

typedef unsigned __int128 u128;

typedef struct {
    uint64_t w[PC_ACC_LIMBS]; /* two's complement, little endian */
} acc_t;

typedef struct {
    int64_t key;
    uint64_t m0, m1; /* magnitude */
    int32_t neg;
    uint32_t pad;
    int64_t deg;
} term_t;

/* next == NULL: an empty bucket head; PC_END: a used node that ends its chain (calloc'ed / fresh pages = all empty) */
typedef struct node {
    int64_t key;
    acc_t cf;
    struct node *next;
} node_t;
#define PC_END ((node_t *)(uintptr_t)1)
#define PC_USED(n) ((n)->next != NULL)

typedef struct arena {
    struct arena *prev;
    size_t used, cap;
    node_t nodes[];
} arena_t;

typedef struct {
    node_t *b;      /* bucket array: first node inline, hash_set.hpp:128-174 */
    uint64_t nb;    /* power of two */
    uint64_t mask;
    size_t bytes;   /* mapped size */
} table_t;

typedef struct {
    uint64_t n;
    int64_t *key;
    uint64_t *mag; /* n * limbs */
    int8_t *sign;
    uint32_t limbs;
    double seconds; /* the region the reference's simple_timer encloses (benchmarks/fateman1.hpp:52-55) */
    uint64_t buckets;
    uint32_t threads_used;
} pc_result;

/* ---------------------------------------------------------------- multiprecision helpers */

static inline void acc_zero(acc_t *a) { memset(a, 0, sizeof(*a)); }
static inline int acc_is_zero(const acc_t *a)
{
    uint64_t o = 0;
    for (int i = 0; i < PC_L; ++i) o |= a->w[i];
    return o == 0;
}

/* |t1| * |t2| -> 256-bit magnitude */
static inline void mag_mul(const term_t *x, const term_t *y, uint64_t p[4])
{
    u128 t = (u128)x->m0 * y->m0;
    p[0] = (uint64_t)t;
    u128 c = t >> 64;
    if ((x->m1 | y->m1) == 0) {
        p[1] = (uint64_t)c;
        p[2] = p[3] = 0;
        return;
    }
    u128 a = (u128)x->m0 * y->m1, b = (u128)x->m1 * y->m0;
    c += (uint64_t)a;
    c += (uint64_t)b;
    p[1] = (uint64_t)c;
    c >>= 64;
    c += (a >> 64) + (b >> 64);
    u128 d = (u128)x->m1 * y->m1;
    c += (uint64_t)d;
    p[2] = (uint64_t)c;
    c >>= 64;
    c += d >> 64;
    p[3] = (uint64_t)c;
}

static inline void acc_add_mag(acc_t *a, const uint64_t p[4], int neg)
{
    /* (the width bound guarantees p[PC_L..3] == 0 and no carry out of the top limb) */
    if (!neg) {
        u128 c = 0;
        for (int i = 0; i < PC_L; ++i) {
            c += (u128)a->w[i] + p[i];
            a->w[i] = (uint64_t)c;
            c >>= 64;
        }
    } else {
        u128 br = 0;
        for (int i = 0; i < PC_L; ++i) {
            u128 d = (u128)a->w[i] - p[i] - br;
            a->w[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }
}

/* ---------------------------------------------------------------- mt19937 (estimator only) */

typedef struct {
    uint32_t mt[624];
    int idx;
} mt_t;

static void mt_seed(mt_t *m, uint32_t s)
{
    m->mt[0] = s;
    for (int i = 1; i < 624; ++i) m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(mt_t *m)
{
    if (m->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7fffffffu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
/* uniform integer in [0, range) -- same nearly-divisionless scheme libstdc++ uses */
static uint32_t mt_below(mt_t *m, uint32_t range)
{
    uint64_t prod = (uint64_t)mt_next(m) * range;
    uint32_t low = (uint32_t)prod;
    if (low < range) {
        uint32_t thr = (uint32_t)(-range) % range;
        while (low < thr) {
            prod = (uint64_t)mt_next(m) * range;
            low = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

/* ---------------------------------------------------------------- persistent thread pool
 * thread_pool.hpp:75-215: one worker thread per slot, created once and kept; a job is handed to slot t and the caller
 * waits for all of them (future_list::wait_all, thread_pool.hpp:516-612).  The benchmarks bind worker t to core t
 * (settings::set_thread_binding(true), benchmarks/fateman1.cpp:50-54); so does this pool. */

typedef struct {
    pthread_t th;
    pthread_mutex_t m;
    pthread_cond_t cv_job, cv_done;
    void *(*fn)(void *);
    void *arg;
    int has, done, started;
    uint32_t slot;
} pool_slot_t;

#define PC_POOL_MAX 1024u
static pool_slot_t *g_pool[PC_POOL_MAX];
static pthread_mutex_t g_pool_m = PTHREAD_MUTEX_INITIALIZER;

static void *pool_main(void *p)
{
    pool_slot_t *s = (pool_slot_t *)p;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu > 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET((int)(s->slot % (uint32_t)ncpu), &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    pthread_mutex_lock(&s->m);
    for (;;) {
        while (!s->has) pthread_cond_wait(&s->cv_job, &s->m);
        s->has = 0;
        void *(*fn)(void *) = s->fn;
        void *arg = s->arg;
        pthread_mutex_unlock(&s->m);
        fn(arg);
        pthread_mutex_lock(&s->m);
        s->done = 1;
        pthread_cond_signal(&s->cv_done);
    }
    return NULL;
}

/* fn(args + t * stride) on pool slot t for t in [0, nt); returns when all are done */
static int pool_run(uint32_t nt, void *(*fn)(void *), void *args, size_t stride)
{
    if (nt > PC_POOL_MAX) return PC_E_ARGS;
    pthread_mutex_lock(&g_pool_m);
    for (uint32_t t = 0; t < nt; ++t) {
        if (g_pool[t]) continue;
        pool_slot_t *s = (pool_slot_t *)calloc(1, sizeof(pool_slot_t));
        if (!s) {
            pthread_mutex_unlock(&g_pool_m);
            return PC_E_NOMEM;
        }
        pthread_mutex_init(&s->m, NULL);
        pthread_cond_init(&s->cv_job, NULL);
        pthread_cond_init(&s->cv_done, NULL);
        s->slot = t;
        if (pthread_create(&s->th, NULL, pool_main, s)) {
            free(s);
            pthread_mutex_unlock(&g_pool_m);
            return PC_E_NOMEM;
        }
        pthread_detach(s->th);
        g_pool[t] = s;
    }
    for (uint32_t t = 0; t < nt; ++t) {
        pool_slot_t *s = g_pool[t];
        pthread_mutex_lock(&s->m);
        s->fn = fn;
        s->arg = (char *)args + (size_t)t * stride;
        s->done = 0;
        s->has = 1;
        pthread_cond_signal(&s->cv_job);
        pthread_mutex_unlock(&s->m);
    }
    for (uint32_t t = 0; t < nt; ++t) {
        pool_slot_t *s = g_pool[t];
        pthread_mutex_lock(&s->m);
        while (!s->done) pthread_cond_wait(&s->cv_done, &s->m);
        pthread_mutex_unlock(&s->m);
    }
    pthread_mutex_unlock(&g_pool_m);
    return 0;
}

/* ---------------------------------------------------------------- multiplier state */

typedef struct {
    term_t *v1, *v2; /* m_v1 (larger operand), m_v2 */
    uint64_t n1, n2;
    uint32_t n_threads;
    uint32_t n_vars;
    const int64_t *limits; /* kronecker_array limits M_i for n_vars */
    int64_t h_min;
    table_t tab;
    arena_t **arenas; /* one per thread */
    const uint64_t *sl; /* skip limits (truncated path) or NULL */
    atomic_flag *locks; /* one per bucket (threaded plain path) */
} mult_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static node_t *arena_new_node(arena_t **ap)
{
    arena_t *a = *ap;
    if (!a || a->used == a->cap) {
        size_t cap = a ? a->cap * 2 : 4096;
        arena_t *n = (arena_t *)malloc(sizeof(arena_t) + cap * sizeof(node_t));
        if (!n) return NULL;
        n->prev = a;
        n->used = 0;
        n->cap = cap;
        *ap = a = n;
    }
    return &a->nodes[a->used++];
}
static void arena_free_all(arena_t *a)
{
    while (a) {
        arena_t *p = a->prev;
        free(a);
        a = p;
    }
}

/* kronecker_monomial::unpack -> kronecker_array::decode, kronecker_array.hpp:353-364 */
static void unpack(const mult_t *m, int64_t key, int64_t *out)
{
    int64_t code = key - m->h_min;
    for (uint32_t i = 0; i < m->n_vars; ++i) {
        int64_t r = 2 * m->limits[i] + 1;
        out[i] = code % r - m->limits[i];
        code /= r;
    }
}

/* hash_set::_bucket_from_hash with hash(term) = size_t(key): kronecker_monomial.hpp:441-444 */
static inline uint64_t bucket_of(const table_t *t, int64_t key) { return (uint64_t)key & t->mask; }

/* one term-by-term product into the table: polynomial.hpp:1742-1757 */
static inline int consume_pair(mult_t *m, arena_t **ar, const term_t *t1, const term_t *t2)
{
    const int64_t key = (int64_t)((uint64_t)t1->key + (uint64_t)t2->key);
    node_t *b = &m->tab.b[bucket_of(&m->tab, key)];
    uint64_t p[4];
    mag_mul(t1, t2, p);
    const int neg = t1->neg ^ t2->neg;
    if (!PC_USED(b)) { /* empty bucket: inline node, hash_set.hpp:295-310 */
        b->key = key;
        acc_zero(&b->cf);
        acc_add_mag(&b->cf, p, neg);
        b->next = PC_END;
        return 0;
    }
    node_t *cur = b, *last = b;
    while (cur != PC_END) { /* _find */
        if (cur->key == key) {
            acc_add_mag(&cur->cf, p, neg); /* multiply_accumulate */
            return 0;
        }
        last = cur;
        cur = cur->next;
    }
    node_t *n = arena_new_node(ar); /* _unique_insert: new node at the chain end */
    if (!n) return PC_E_NOMEM;
    n->key = key;
    acc_zero(&n->cf);
    acc_add_mag(&n->cf, p, neg);
    n->next = PC_END;
    last->next = n;
    return 0;
}

/* ---------------------------------------------------------------- estimate_final_series_size */

typedef struct {
    int64_t *keys;
    uint64_t cap, mask;
    uint8_t *used;
} kset_t;

static int kset_insert(kset_t *s, int64_t k) /* returns 1 if new */
{
    uint64_t h = ((uint64_t)k * 0x9E3779B97F4A7C15ull) >> 17;
    for (;;) {
        h &= s->mask;
        if (!s->used[h]) {
            s->used[h] = 1;
            s->keys[h] = k;
            return 1;
        }
        if (s->keys[h] == k) return 0;
        ++h;
    }
}

static uint64_t lf_of(const mult_t *m, uint64_t i) { return m->sl ? (m->sl[i] < m->n2 ? m->sl[i] : m->n2) : m->n2; }

typedef struct {
    const mult_t *m;
    unsigned tr0, tr1; /* trials [tr0, tr1) */
    u128 total;
    int rc;
} est_worker_t;

/* the trials [tr0, tr1): every trial is seeded by its own index, so the sum does not depend on how they are dealt out */
static void *estimate_trials(void *p)
{
    est_worker_t *w = (est_worker_t *)p;
    const mult_t *m = w->m;
    const uint64_t n1 = m->n1;
    const unsigned multiplier = 2;
    kset_t s;
    s.cap = 1;
    while (s.cap < 2 * n1 + 2) s.cap <<= 1;
    s.mask = s.cap - 1;
    s.keys = (int64_t *)malloc(s.cap * sizeof(int64_t));
    s.used = (uint8_t *)malloc(s.cap);
    uint64_t *idx = (uint64_t *)malloc(n1 * sizeof(uint64_t));
    if (!s.keys || !s.used || !idx) {
        free(s.keys);
        free(s.used);
        free(idx);
        w->rc = PC_E_NOMEM;
        return NULL;
    }
    u128 total = 0;
    for (unsigned tr = w->tr0; tr < w->tr1; ++tr) {
        mt_t eng;
        mt_seed(&eng, tr); /* seeded by the global trial index: base_series_multiplier.hpp:627-629 */
        for (uint64_t i = 0; i < n1; ++i) idx[i] = i;
        for (uint64_t i = n1 - 1; i > 0; --i) { /* shuffle */
            uint64_t j = mt_below(&eng, (uint32_t)(i + 1));
            uint64_t t = idx[i];
            idx[i] = idx[j];
            idx[j] = t;
        }
        memset(s.used, 0, s.cap);
        uint64_t count = 0, it = 0;
        u128 acc_s2 = 0;
        for (; it < n1; ++it) {
            const uint64_t lim = lf_of(m, idx[it]);
            if (!lim) continue;
            acc_s2 += lim;
            const uint64_t j = mt_below(&eng, (uint32_t)lim);
            const int64_t k = (int64_t)((uint64_t)m->v1[idx[it]].key + (uint64_t)m->v2[j].key);
            if (!kset_insert(&s, k)) break; /* first duplicate */
            ++count;
        }
        u128 add = (it == n1) ? acc_s2 : (u128)multiplier * count * count;
        if (add == 0) add = 1;
        total += add;
    }
    free(s.keys);
    free(s.used);
    free(idx);
    w->total = total;
    return NULL;
}

static uint64_t estimate_final_series_size(const mult_t *m)
{
    const uint64_t n1 = m->n1, n2 = m->n2;
    if (!n1 || !n2) return 1;
    if (n1 == 1 || n2 == 1) return n1 * n2;
    const unsigned n_trials = 15;
    /* the trials are dealt out over the threads (base_series_multiplier.hpp:592-718) */
    uint32_t nt = m->n_threads < n_trials ? m->n_threads : n_trials;
    if (nt < 1) nt = 1;
    est_worker_t ws[15];
    for (uint32_t t = 0; t < nt; ++t) {
        ws[t].m = m;
        ws[t].tr0 = n_trials * t / nt;
        ws[t].tr1 = n_trials * (t + 1) / nt;
        ws[t].total = 0;
        ws[t].rc = 0;
    }
    if (nt == 1) estimate_trials(&ws[0]);
    else if (pool_run(nt, estimate_trials, ws, sizeof(est_worker_t))) return n1; /* (cannot size: a small table still gives the right result) */
    u128 total = 0;
    for (uint32_t t = 0; t < nt; ++t) total += ws[t].total;
    uint64_t est = (uint64_t)(total / n_trials);
    return est ? est : 1;
}

static int table_rehash(table_t *t, uint64_t want, uint32_t n_threads)
{
    (void)n_threads; /* calloc pages are zeroed by the kernel on first touch by the worker threads */
    uint64_t nb = 1;
    while (nb < want) nb <<= 1;
    /* fresh anonymous pages are zero = empty buckets; huge pages when the kernel grants them */
    const size_t bytes = (((size_t)nb * sizeof(node_t)) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    void *mem = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) return PC_E_NOMEM;
    madvise(mem, bytes, MADV_HUGEPAGE);
    t->b = (node_t *)mem;
    t->bytes = bytes;
    t->nb = nb;
    t->mask = nb - 1;
    return 0;
}

/* ---------------------------------------------------------------- stable sorts (merge sort on indices) */

typedef struct {
    uint64_t k;
    uint64_t i;
} ki_t;

static void ki_msort(ki_t *a, ki_t *tmp, uint64_t n)
{
    if (n < 2) return;
    uint64_t h = n / 2;
    ki_msort(a, tmp, h);
    ki_msort(a + h, tmp, n - h);
    uint64_t i = 0, j = h, o = 0;
    while (i < h && j < n) tmp[o++] = (a[j].k < a[i].k) ? a[j++] : a[i++];
    while (i < h) tmp[o++] = a[i++];
    while (j < n) tmp[o++] = a[j++];
    memcpy(a, tmp, n * sizeof(ki_t));
}

/* stable-sort terms by an unsigned 64-bit rank */
static int sort_terms_by(term_t *v, uint64_t n, const uint64_t *rank)
{
    ki_t *a = (ki_t *)malloc(2 * n * sizeof(ki_t));
    term_t *cp = (term_t *)malloc(n * sizeof(term_t));
    if (!a || !cp) {
        free(a);
        free(cp);
        return PC_E_NOMEM;
    }
    for (uint64_t i = 0; i < n; ++i) {
        a[i].k = rank[i];
        a[i].i = i;
    }
    ki_msort(a, a + n, n);
    for (uint64_t i = 0; i < n; ++i) cp[i] = v[a[i].i];
    memcpy(v, cp, n * sizeof(term_t));
    free(a);
    free(cp);
    return 0;
}

/* ---------------------------------------------------------------- tasks and zones */

typedef struct {
    uint64_t i, j0, j1;
} task_t;

typedef struct {
    task_t *t;
    uint64_t n, cap;
} tvec_t;

static int tvec_push(tvec_t *v, uint64_t i, uint64_t j0, uint64_t j1)
{
    if (v->n == v->cap) {
        uint64_t c = v->cap ? v->cap * 2 : 1024;
        task_t *p = (task_t *)realloc(v->t, c * sizeof(task_t));
        if (!p) return PC_E_NOMEM;
        v->t = p;
        v->cap = c;
    }
    v->t[v->n].i = i;
    v->t[v->n].j0 = j0;
    v->t[v->n].j1 = j1;
    ++v->n;
    return 0;
}

/* task_split, polynomial.hpp:1709-1718 */
static int task_split(tvec_t *out, uint64_t i, uint64_t start, uint64_t end)
{
    while (end - start > PC_BLOCK) {
        if (tvec_push(out, i, start, start + PC_BLOCK)) return PC_E_NOMEM;
        start += PC_BLOCK;
    }
    if (end != start && tvec_push(out, i, start, end)) return PC_E_NOMEM;
    return 0;
}

static int sort_tasks(const mult_t *m, tvec_t *tv)
{
    if (tv->n < 2) return 0;
    ki_t *a = (ki_t *)malloc(2 * tv->n * sizeof(ki_t));
    task_t *cp = (task_t *)malloc(tv->n * sizeof(task_t));
    if (!a || !cp) {
        free(a);
        free(cp);
        return PC_E_NOMEM;
    }
    for (uint64_t q = 0; q < tv->n; ++q) { /* task_cmp, polynomial.hpp:1702-1705 */
        a[q].k = bucket_of(&m->tab, m->v1[tv->t[q].i].key) + bucket_of(&m->tab, m->v2[tv->t[q].j0].key);
        a[q].i = q;
    }
    ki_msort(a, a + tv->n, tv->n);
    for (uint64_t q = 0; q < tv->n; ++q) cp[q] = tv->t[a[q].i];
    memcpy(tv->t, cp, tv->n * sizeof(task_t));
    free(a);
    free(cp);
    return 0;
}

/* l_bound, polynomial.hpp:1800-1826: first j with bucket(v2[j]) >= zb - bucket(v1[i]) */
static uint64_t l_bound(const mult_t *m, uint64_t zb, uint64_t i)
{
    const uint64_t ib = bucket_of(&m->tab, m->v1[i].key);
    if (zb < ib) return 0;
    const uint64_t cmp = zb - ib;
    uint64_t first = 0, count = m->n2;
    while (count > 0) {
        uint64_t step = count / 2, idx = first + step;
        if (bucket_of(&m->tab, m->v2[idx].key) < cmp) {
            first = idx + 1;
            count -= step + 1;
        } else {
            count = step;
        }
    }
    return first;
}

typedef struct {
    mult_t *m;
    tvec_t *zones;
    uint64_t n_zones;
    atomic_flag *claimed;
    uint32_t tid;
    int rc;
    uint64_t bpz;
} worker_t;

static int task_consume(mult_t *m, arena_t **ar, const task_t *t)
{
    const term_t *t1 = &m->v1[t->i];
    for (uint64_t j = t->j0; j < t->j1; ++j) {
        int rc = consume_pair(m, ar, t1, &m->v2[j]);
        if (rc) return rc;
    }
    return 0;
}

/* table_filler, polynomial.hpp:1828-1867 (one call per thread: its zm zones) */
static void *zone_builder(void *p)
{
    worker_t *w = (worker_t *)p;
    mult_t *m = w->m;
    const uint64_t nb = m->tab.nb;
    for (unsigned n = 0; n < PC_ZM; ++n) {
        tvec_t *tv = &w->zones[(uint64_t)w->tid * PC_ZM + n];
        uint64_t a = (uint64_t)w->tid * w->bpz * PC_ZM + n * w->bpz;
        uint64_t b = (n == PC_ZM - 1 && w->tid == m->n_threads - 1) ? nb : a + w->bpz;
        for (int batch = 0; batch < 2; ++batch) { /* second batch: bucket sums that wrapped past 2^k */
            const uint64_t off = batch ? nb : 0;
            for (uint64_t i = 0; i < m->n1; ++i) {
                uint64_t j0 = l_bound(m, a + off, i), j1 = l_bound(m, b + off, i);
                if (j0 == 0 && j1 == 0) break;
                if (task_split(tv, i, j0, j1)) {
                    w->rc = PC_E_NOMEM;
                    return NULL;
                }
            }
        }
        if (sort_tasks(m, tv)) {
            w->rc = PC_E_NOMEM;
            return NULL;
        }
    }
    return NULL;
}

/* thread_functor, polynomial.hpp:1921-1947 */
static void *zone_worker(void *p)
{
    worker_t *w = (worker_t *)p;
    uint64_t t_idx = (uint64_t)w->tid * PC_ZM;
    const uint64_t start = t_idx;
    for (;;) {
        if (!atomic_flag_test_and_set(&w->claimed[t_idx])) {
            const tvec_t *tv = &w->zones[t_idx];
            for (uint64_t q = 0; q < tv->n; ++q) {
                int rc = task_consume(w->m, &w->m->arenas[w->tid], &tv->t[q]);
                if (rc) {
                    w->rc = rc;
                    return NULL;
                }
            }
        }
        if (++t_idx == w->n_zones) t_idx = 0;
        if (t_idx == start) break;
    }
    return NULL;
}

static int sparse_kronecker_multiplication(mult_t *m)
{
    int rc = 0;
    /* sort both operands by destination bucket: polynomial.hpp:1691-1695 */
    for (int s = 0; s < 2 && !rc; ++s) {
        term_t *v = s ? m->v2 : m->v1;
        uint64_t n = s ? m->n2 : m->n1;
        uint64_t *rank = (uint64_t *)malloc(n * sizeof(uint64_t));
        if (!rank) return PC_E_NOMEM;
        for (uint64_t i = 0; i < n; ++i) rank[i] = bucket_of(&m->tab, v[i].key);
        rc = sort_terms_by(v, n, rank);
        free(rank);
    }
    if (rc) return rc;
    if (m->n_threads == 1) { /* polynomial.hpp:1760-1782 */
        tvec_t tv = {0, 0, 0};
        for (uint64_t i = 0; i < m->n1 && !rc; ++i) rc = task_split(&tv, i, 0, m->n2);
        if (!rc) rc = sort_tasks(m, &tv);
        for (uint64_t q = 0; q < tv.n && !rc; ++q) rc = task_consume(m, &m->arenas[0], &tv.t[q]);
        free(tv.t);
        return rc;
    }
    const uint64_t n_zones = (uint64_t)m->n_threads * PC_ZM;
    tvec_t *zones = (tvec_t *)calloc(n_zones, sizeof(tvec_t));
    atomic_flag *claimed = (atomic_flag *)malloc(n_zones * sizeof(atomic_flag));
    worker_t *ws = (worker_t *)calloc(m->n_threads, sizeof(worker_t));
    for (uint64_t z = 0; z < n_zones; ++z) atomic_flag_clear(&claimed[z]);
    for (uint32_t t = 0; t < m->n_threads; ++t) {
        ws[t].m = m;
        ws[t].zones = zones;
        ws[t].n_zones = n_zones;
        ws[t].claimed = claimed;
        ws[t].tid = t;
        ws[t].bpz = m->tab.nb / n_zones;
    }
    for (int phase = 0; phase < 2 && !rc; ++phase) {
        const double tp = now_s();
        rc = pool_run(m->n_threads, phase ? zone_worker : zone_builder, ws, sizeof(worker_t));
        for (uint32_t t = 0; t < m->n_threads; ++t)
            if (ws[t].rc) rc = ws[t].rc;
        if (getenv("PC_TIMING")) fprintf(stderr, "[pc] %s %.1f ms\n", phase ? "zone workers" : "zone builders", (now_s() - tp) * 1e3);
    }
    for (uint64_t z = 0; z < n_zones; ++z) free(zones[z].t);
    free(zones);
    free(claimed);
    free(ws);
    return rc;
}

/* ---------------------------------------------------------------- plain (blocked) multiplication */

typedef struct {
    mult_t *m;
    uint64_t start1, end1;
    uint32_t tid;
    int locked;
    int rc;
} pworker_t;

static inline int plain_pair(pworker_t *w, uint64_t i, uint64_t j)
{
    mult_t *m = w->m;
    if (!w->locked) return consume_pair(m, &m->arenas[w->tid], &m->v1[i], &m->v2[j]);
    /* per-bucket spinlock: base_series_multiplier.hpp:1073-1081, detail/atomic_lock_guard.hpp:45-60 */
    const int64_t key = (int64_t)((uint64_t)m->v1[i].key + (uint64_t)m->v2[j].key);
    atomic_flag *l = &m->locks[bucket_of(&m->tab, key)];
    while (atomic_flag_test_and_set_explicit(l, memory_order_acquire)) {
    }
    int rc = consume_pair(m, &m->arenas[w->tid], &m->v1[i], &m->v2[j]);
    atomic_flag_clear_explicit(l, memory_order_release);
    return rc;
}

/* blocked_multiplication, base_series_multiplier.hpp:449-504 */
static void *blocked_multiplication(void *p)
{
    pworker_t *w = (pworker_t *)p;
    mult_t *m = w->m;
    const uint64_t bs = PC_BLOCK, n2 = m->n2;
    const uint64_t nblocks1 = (w->end1 - w->start1) / bs, nblocks2 = n2 / bs;
    const uint64_t i_ir_start = nblocks1 * bs + w->start1, i_ir_end = w->end1;
    const uint64_t j_ir_start = nblocks2 * bs, j_ir_end = n2;
#define PC_ROWS(i_lo, i_hi, j_lo, j_hi)                                          \
    for (uint64_t i = (i_lo); i < (i_hi); ++i) {                                 \
        const uint64_t lf = lf_of(m, i), lim = lf < (j_hi) ? lf : (j_hi);        \
        for (uint64_t j = (j_lo); j < lim; ++j)                                  \
            if ((w->rc = plain_pair(w, i, j))) return NULL;                      \
    }
    for (uint64_t b1 = 0; b1 < nblocks1; ++b1) {
        const uint64_t i0 = b1 * bs + w->start1, i1 = i0 + bs;
        for (uint64_t b2 = 0; b2 < nblocks2; ++b2) PC_ROWS(i0, i1, b2 * bs, b2 * bs + bs)
        PC_ROWS(i0, i1, j_ir_start, j_ir_end)
    }
    for (uint64_t b2 = 0; b2 < nblocks2; ++b2) PC_ROWS(i_ir_start, i_ir_end, b2 * bs, b2 * bs + bs)
    PC_ROWS(i_ir_start, i_ir_end, j_ir_start, j_ir_end)
#undef PC_ROWS
    return NULL;
}

static int plain_multiplication(mult_t *m)
{
    /* base_series_multiplier.hpp:1009-1029: estimate unless the product is tiny and single-threaded */
    uint64_t want = 1;
    const int estimate = !((u128)m->n1 * m->n2 < (u128)PC_EST_THRESHOLD * PC_EST_THRESHOLD && m->n_threads == 1);
    if (estimate) want = estimate_final_series_size(m);
    else {
        /* without an estimate the reference inserts through series::insert(), which grows the table as
         * needed; a table sized by the pair count is an equivalent container for a restatement. */
        want = m->n1 * m->n2;
        if (want < 1) want = 1;
    }
    int rc = table_rehash(&m->tab, want, m->n_threads);
    if (rc) return rc;
    if (m->n_threads == 1) {
        pworker_t w = {m, 0, m->n1, 0, 0, 0};
        blocked_multiplication(&w);
        return w.rc;
    }
    m->locks = (atomic_flag *)malloc(m->tab.nb * sizeof(atomic_flag));
    if (!m->locks) return PC_E_NOMEM;
    for (uint64_t i = 0; i < m->tab.nb; ++i) atomic_flag_clear(&m->locks[i]);
    pworker_t *ws = (pworker_t *)calloc(m->n_threads, sizeof(pworker_t));
    const uint64_t block = m->n1 / m->n_threads; /* base_series_multiplier.hpp:1054 */
    for (uint32_t t = 0; t < m->n_threads; ++t) {
        ws[t].m = m;
        ws[t].start1 = t * block;
        ws[t].end1 = (t == m->n_threads - 1) ? m->n1 : (t + 1) * block;
        ws[t].tid = t;
        ws[t].locked = 1;
    }
    rc = pool_run(m->n_threads, blocked_multiplication, ws, sizeof(pworker_t));
    for (uint32_t t = 0; t < m->n_threads; ++t)
        if (ws[t].rc) rc = ws[t].rc;
    free(ws);
    return rc;
}

/* ---------------------------------------------------------------- sanitise (count pass) */

typedef struct {
    const table_t *tab;
    uint64_t b0, b1, count;
} sanitise_t;

static void *sanitise_worker(void *p)
{
    sanitise_t *w = (sanitise_t *)p;
    uint64_t cnt = 0;
    for (uint64_t b = w->b0; b < w->b1; ++b)
        if (PC_USED(&w->tab->b[b]))
            for (const node_t *c = &w->tab->b[b]; c != PC_END; c = c->next)
                if (!acc_is_zero(&c->cf)) ++cnt;
    w->count = cnt;
    return NULL;
}

/* ---------------------------------------------------------------- public entry point */

static uint32_t use_threads(uint64_t n1, uint64_t n2, uint32_t n_threads, uint64_t min_work)
{
    /* thread_pool.hpp:463-490: min(pool size, work / min_work), at least 1 */
    if (!n1 || !n2 || n_threads <= 1) return 1;
    u128 work = (u128)n1 * n2;
    u128 q = work / (min_work ? min_work : 1);
    if (q < 1) q = 1;
    return (uint32_t)(q < n_threads ? q : n_threads);
}

static void load_terms(term_t *v, uint64_t n, const int64_t *key, const uint64_t *mag, uint32_t limbs, const int8_t *sign)
{
    for (uint64_t i = 0; i < n; ++i) {
        v[i].key = key[i];
        v[i].m0 = mag[i * limbs];
        v[i].m1 = limbs > 1 ? mag[i * limbs + 1] : 0;
        v[i].neg = sign[i] < 0;
        v[i].deg = 0;
    }
}

static int upper_bound_deg(const term_t *v, uint64_t n, int64_t c, uint64_t *out)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (c < v[mid].deg) hi = mid;
        else lo = mid + 1;
    }
    *out = lo;
    return 0;
}

/*
 * trunc_mode: 0 none, 1 total degree, 2 partial degree over the variables with a non-zero entry in
 * deg_mask[n_vars].  min_work: settings::min_work_per_thread (pass 0 for the default 250000).
 */
int pc_mul(uint64_t n1, const int64_t *key1, const uint64_t *mag1, uint32_t limbs1, const int8_t *sign1,
           uint64_t n2, const int64_t *key2, const uint64_t *mag2, uint32_t limbs2, const int8_t *sign2,
           uint32_t n_vars, const int64_t *limits, int64_t h_min, uint32_t n_threads, uint64_t min_work,
           int trunc_mode, int64_t max_degree, const uint8_t *deg_mask, pc_result *out)
{
    if (!out || limbs1 < 1 || limbs1 > 2 || limbs2 < 1 || limbs2 > 2 || n_threads < 1) return PC_E_ARGS;
    memset(out, 0, sizeof(*out));
    out->limbs = PC_ACC_LIMBS;
    const double t0 = now_s();
    mult_t m;
    memset(&m, 0, sizeof(m));
    m.n_vars = n_vars;
    m.limits = limits;
    m.h_min = h_min;
    /* larger series first: base_series_multiplier.hpp:368-372 */
    const int swap = n1 < n2;
    m.n1 = swap ? n2 : n1;
    m.n2 = swap ? n1 : n2;
    m.v1 = (term_t *)malloc((m.n1 ? m.n1 : 1) * sizeof(term_t));
    m.v2 = (term_t *)malloc((m.n2 ? m.n2 : 1) * sizeof(term_t));
    if (!m.v1 || !m.v2) return PC_E_NOMEM;
    if (swap) {
        load_terms(m.v1, n2, key2, mag2, limbs2, sign2);
        load_terms(m.v2, n1, key1, mag1, limbs1, sign1);
    } else {
        load_terms(m.v1, n1, key1, mag1, limbs1, sign1);
        load_terms(m.v2, n2, key2, mag2, limbs2, sign2);
    }
    m.n_threads = use_threads(m.n1, m.n2, n_threads, min_work ? min_work : PC_MIN_WORK);
    out->threads_used = m.n_threads;
    int rc = 0;
    uint64_t *sl = NULL;
    if (!m.n1 || !m.n2) goto done; /* empty operand -> empty result, polynomial.hpp:1654-1656 */
    /* check_bounds + degrees */
    if (n_vars) {
        int64_t mn[2][64], mx[2][64], e[64];
        if (n_vars > 64) {
            rc = PC_E_ARGS;
            goto done;
        }
        for (int s = 0; s < 2; ++s) {
            term_t *v = s ? m.v2 : m.v1;
            uint64_t n = s ? m.n2 : m.n1;
            for (uint64_t i = 0; i < n; ++i) {
                unpack(&m, v[i].key, e);
                int64_t d = 0;
                for (uint32_t k = 0; k < n_vars; ++k) {
                    if (i == 0 || e[k] < mn[s][k]) mn[s][k] = e[k];
                    if (i == 0 || e[k] > mx[s][k]) mx[s][k] = e[k];
                    if (trunc_mode == 1 || (trunc_mode == 2 && deg_mask && deg_mask[k])) d += e[k];
                }
                v[i].deg = d;
            }
        }
        for (uint32_t k = 0; k < n_vars; ++k) {
            __int128 lo = (__int128)mn[0][k] + mn[1][k], hi = (__int128)mx[0][k] + mx[1][k];
            if (lo < -(__int128)limits[k] || hi > (__int128)limits[k]) {
                rc = PC_E_KEY_RANGE; /* std::overflow_error, polynomial.hpp:1190-1192 */
                goto done;
            }
        }
    }
    m.arenas = (arena_t **)calloc(m.n_threads, sizeof(arena_t *));
    if (trunc_mode) {
        /* _truncated_multiplication, polynomial.hpp:1402-1449 */
        uint64_t *rank = (uint64_t *)malloc(m.n2 * sizeof(uint64_t));
        for (uint64_t j = 0; j < m.n2; ++j) rank[j] = (uint64_t)m.v2[j].deg ^ 0x8000000000000000ull;
        rc = sort_terms_by(m.v2, m.n2, rank);
        free(rank);
        if (rc) goto done;
        sl = (uint64_t *)malloc(m.n1 * sizeof(uint64_t));
        for (uint64_t i = 0; i < m.n1; ++i) { /* _get_skip_limits, polynomial.hpp:1495-1509 */
            __int128 c = (__int128)max_degree - m.v1[i].deg;
            if (c > INT64_MAX || c < INT64_MIN) {
                rc = PC_E_KEY_RANGE; /* safe_int_sub overflow -> std::overflow_error */
                goto done;
            }
            upper_bound_deg(m.v2, m.n2, (int64_t)c, &sl[i]);
        }
        m.sl = sl;
        rc = plain_multiplication(&m);
    } else if ((u128)m.n1 * m.n2 < (u128)PC_EST_THRESHOLD * PC_EST_THRESHOLD && m.n_threads == 1) {
        rc = plain_multiplication(&m); /* polynomial.hpp:1640-1649 */
    } else {
        const uint64_t est = estimate_final_series_size(&m);
        if (getenv("PC_TIMING")) fprintf(stderr, "[pc] estimate done at %.1f ms\n", (now_s() - t0) * 1e3);
        rc = table_rehash(&m.tab, est, m.n_threads);
        if (!rc) rc = sparse_kronecker_multiplication(&m);
    }
    if (rc) goto done;
    if (getenv("PC_TIMING")) fprintf(stderr, "[pc] multiplication done at %.1f ms\n", (now_s() - t0) * 1e3);
    /* sanitise_series: count the non-zero terms (zeros are dropped at extraction below); one bucket range per thread,
     * base_series_multiplier.hpp:885-948 */
    {
        const uint32_t nt = (m.n_threads > 1 && m.tab.nb >= 4096) ? m.n_threads : 1;
        sanitise_t *sw = (sanitise_t *)calloc(nt, sizeof(sanitise_t));
        if (!sw) {
            rc = PC_E_NOMEM;
            goto done;
        }
        for (uint32_t t = 0; t < nt; ++t) {
            sw[t].tab = &m.tab;
            sw[t].b0 = m.tab.nb / nt * t;
            sw[t].b1 = (t + 1 == nt) ? m.tab.nb : m.tab.nb / nt * (t + 1);
        }
        if (nt == 1) sanitise_worker(&sw[0]);
        else rc = pool_run(nt, sanitise_worker, sw, sizeof(sanitise_t));
        uint64_t cnt = 0;
        for (uint32_t t = 0; t < nt; ++t) cnt += sw[t].count;
        free(sw);
        if (rc) goto done;
        out->n = cnt;
        out->buckets = m.tab.nb;
    }
    if (getenv("PC_TIMING")) fprintf(stderr, "[pc] sanitise done at %.1f ms\n", (now_s() - t0) * 1e3);
done:
    out->seconds = now_s() - t0;
    if (!rc && out->n) { /* extraction to SoA: outside the timed region */
        out->key = (int64_t *)malloc(out->n * sizeof(int64_t));
        out->mag = (uint64_t *)malloc(out->n * PC_ACC_LIMBS * sizeof(uint64_t));
        out->sign = (int8_t *)malloc(out->n);
        if (!out->key || !out->mag || !out->sign) rc = PC_E_NOMEM;
        uint64_t o = 0;
        for (uint64_t b = 0; b < m.tab.nb && !rc; ++b) {
            if (!PC_USED(&m.tab.b[b])) continue;
            for (node_t *c = &m.tab.b[b]; c != PC_END; c = c->next) {
                if (acc_is_zero(&c->cf)) continue;
                acc_t v = c->cf;
                const int neg = (int)(v.w[PC_L - 1] >> 63);
                if (neg) { /* two's complement -> magnitude */
                    u128 cy = 1;
                    for (int i = 0; i < PC_L; ++i) {
                        cy += (uint64_t)~v.w[i];
                        v.w[i] = (uint64_t)cy;
                        cy >>= 64;
                    }
                }
                out->key[o] = c->key;
                memcpy(&out->mag[o * PC_ACC_LIMBS], v.w, sizeof(v.w));
                out->sign[o] = neg ? -1 : 1;
                ++o;
            }
        }
    }
    if (m.tab.b) munmap(m.tab.b, m.tab.bytes);
    free(m.locks);
    if (m.arenas) {
        for (uint32_t t = 0; t < m.n_threads; ++t) arena_free_all(m.arenas[t]);
        free(m.arenas);
    }
    free(sl);
    free(m.v1);
    free(m.v2);
    if (rc) {
        free(out->key);
        free(out->mag);
        free(out->sign);
        out->key = NULL;
        out->mag = NULL;
        out->sign = NULL;
        out->n = 0;
    }
    return rc;
}

#ifndef PC_NO_RESULT_FREE
void pc_result_free(pc_result *r)
{
    if (!r) return;
    free(r->key);
    free(r->mag);
    free(r->sign);
    memset(r, 0, sizeof(*r));
}
#endif
