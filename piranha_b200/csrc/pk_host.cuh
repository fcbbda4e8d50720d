// Host-side objects shared by the C ABI implementation (pk_abi.cu) and the zone path (pk_zone.cuh).
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "pk_device.cuh"

// ------------------------------------------------------------------------------------------------ objects

namespace pk
{
struct ZoneTuning {
    uint32_t target_zones = 0;    // 0 = 12 per SM
    uint32_t zones_per_chunk = 2;
    uint32_t span_log2_max = 18;  // bitmap zones: at most 2^18 codes (32 KB bitmap + 32 KB prefix)
    uint32_t force_cap = 0;       // tests: force a small accumulator capacity to exercise multi-pass zones
    uint32_t force_mode = 0;      // 0 auto, 1 dense, 2 bitmap
    uint32_t max_zones = 1u << 20;
    uint32_t cursor_bytes_max = 64u << 20; // row cursors live in global memory; 0 disables them
    uint32_t disable = 0;
    uint32_t tpb = 0;              // threads per block: 0 auto, 256 / 512 / 1024
    uint32_t uniform = 0;          // 1 = equal-span zones instead of equal-work (quantile) zones
};

struct BlockTuning {
    uint32_t disable = 0;
    uint32_t tile_target = 1024;  // products per tile of the block path (sweep: 256 is 8 % slower, 1024-2048 best)
    uint32_t zone_work = 1u << 16; // aim for about this many products per zone (more blocks per zone when blocks are light)
    uint32_t min_tile = 8;        // decline the block path when the average pair of groups has fewer products
    uint32_t force_mode = 0;      // 0 auto, 1 dense, 2 bitmap
    uint32_t force_hz = 0;        // tests: blocks per zone
    uint32_t natural = 0;         // 1 = claim zones in code order instead of heaviest first
    uint32_t equal_zones = 1;     // bitmap mode: zones cut at equal work (tile-count quantiles) instead of Hz blocks each:
                                  // 0 never, 1 output partitions only (default), 2 always
    uint32_t zone_products = 0;   // equal-work zones: products per zone (0 = zone_work)
    uint32_t wide = 1;            // 1 = 64-term chunks of gb (two terms per lane) where the group is large enough
    uint32_t pair_walk = 1;       // one-limb operands: the accumulating walk takes two products at a time (two ATOMS chains in flight)
    uint32_t permute = 0;         // 1 = re-order the terms inside B's groups so that the lanes of a tile hit distinct banks
                                  // (measured: 2-3 % on the pairing kernel, the price of one more launch; off)
    uint32_t hash = 0;            // sparse zones: 1 = try a one-walk shared-memory hash accumulation before the two-walk
                                  // bitmap scheme (measured slower on pearce1: 0.60 vs 0.39 ms; experiment, off)
    uint32_t smem_limit = 0;      // bytes of shared memory a zone may use (0 = all): what is left over serves as L1
    uint32_t radix0_residue = 0;  // experiment: pad the radix of digit 0 so that radix mod 32 == this (1..32); 0 = off
    uint32_t unit_codes = 1u << 15; // sparse ranges: codes per zone unit aimed at, times the blocks per SM (a unit is a whole number of blocks)
    uint32_t self_clear = 1;      // sparse ranges: zones clear their bits of the context's bitmap, so the next product skips the memset of the code range
    uint32_t flatcol = 0;         // sparse ranges, accumulating walk: ragged chunks with fixed columns per lane instead of flattened products
    uint32_t stage_rows = 1;      // dense ranges: the rows of A of a full tile are staged in shared memory by a bulk copy (TMA), double
                                  // buffered per warp against the walk
    uint32_t size_hint = 1;       // sparse ranges: size the result from the last product of the same operands and options (skips the
                                  // mid-product read-back; verified on the device, repeated the long way on a miss)
    uint32_t ctas = 4;            // sparse ranges: thread blocks per SM of the accumulating kernel (1 x 1024, 2 x 512 or 4 x 256 threads;
                                  // measured on pearce1: 0.42 / 0.34 / 0.32 ms -- while one block emits or waits at a barrier the others walk)
    uint32_t merge = 1;           // 1 = a tile's rows span all the groups of A that land in the same zone unit with its group of B
};

} // namespace pk

// What the sparse block path remembers of its last product on a context (see block_multiply).
struct SizeHint {
    const void *a = nullptr, *b = nullptr;
    uint64_t nA = 0, nB = 0, range = 0, U = 0;
    uint32_t part = 0, parts = 0, trunc = 0;
    int64_t max_degree = 0;
    uint64_t entries = 0;
    bool valid = false;
    bool same_key(const SizeHint &o) const
    {
        return a == o.a && b == o.b && nA == o.nA && nB == o.nB && range == o.range && U == o.U && part == o.part && parts == o.parts
               && trunc == o.trunc && max_degree == o.max_degree;
    }
};

struct PinnedBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    bool in_use = false;
};

struct pk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr; // device-to-host copies of a pipelined pk_mul (created on first use)
    cudaEvent_t ev_part = nullptr;
    uint32_t pipe_parts = 0;           // output parts of a pipelined pk_mul; 0 or 1 = off (the default: measured on
                                       // pearce1 the copy is 2.9 ms of a 3.9 ms call and the product only 0.65 ms, so
                                       // overlapping them gains under 10 % and costs the per-part fixed work)
    uint64_t pipe_min_products = 1ull << 23;
    uint32_t pipe_test_small = 0;      // tests: undersize the host arrays so that the re-copy path runs
    std::string err;
    std::vector<std::vector<int64_t>> limits; // [n_vars] -> M_k
    std::vector<int64_t> h_min, h_max;
    pk_stats stats{};
    uint64_t total_launches = 0;
    uint32_t call_launches = 0;
    cudaEvent_t ev_begin = nullptr, ev_mul0 = nullptr, ev_mul1 = nullptr, ev_end = nullptr;
    cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr; // the marking walk of the sparse block path
    bool ev_valid = false, aux_valid = false;
    void *tiles_buf = nullptr;  // tile list of the block path, kept across products (grow-only)
    uint64_t tiles_cap = 0;     // in tiles
    SizeHint size_hint;
    // Idle device blocks of this context (>= 256 KB), reused by size before the stream-ordered pool is asked.  The pool splits
    // freed blocks for smaller requests, so a product whose result is tens of MB kept paying for fresh physical memory every
    // few calls (measured: 1-5 ms stalls in the result allocation of partitioned pearce1 products, for a dozen products in a row);
    // everything here lives on the context's one stream, so reuse is stream-ordered like the pool's own.
    struct IdleBlock {
        void *p;
        size_t bytes;
    };
    std::vector<IdleBlock> idle_blocks;
    size_t idle_bytes = 0;
    void *bitmap_buf = nullptr; // sparse block path: bitmap of the code range (k_block_mark -> k_block_acc), grow-only
    size_t bitmap_bytes = 0;
    std::vector<cudaEvent_t> chunk_events; // pk_download_grouped_progress: one event per piece of the copy
    bool bitmap_clean = false;  // the whole buffer is zero (every zone of the last product cleared the bits it consumed)
    void *h_scratch = nullptr; // small pinned read-back area
    pk::MulCounters *d_ctr = nullptr;
    pk::MetaDev *d_meta = nullptr;
    pk::u64 *d_digest = nullptr;
    std::vector<PinnedBuf> pinned;
    int sm_count = 148;
    pk::ZoneTuning zone_tuning;
    pk::BlockTuning block_tuning;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    size_t total_mem = 0;
    pk_dpoly *staging = nullptr; // unordered per-zone output of the shared-memory paths, kept across products (grow-only)
    uint64_t staging_cap = 0;
    cudaMemPool_t pool = nullptr; // private stream-ordered pool: freed blocks stay cached, the device's default pool is untouched
    // multi-device context (pk_ctx_create_multi): one sub-context + one worker thread per device; the master owns the
    // pinned result buffers and no device state of its own
    std::vector<pk_ctx *> subs;
    struct pk_multi_state *multi = nullptr;
};

struct pk_dpoly {
    uint64_t n = 0, stride = 0;
    uint32_t n_vars = 0, limbs = 1;
    uint32_t alloc_limbs = 1;   // planes allocated (limbs may be trimmed below it)
    size_t key_bytes = 0, planes_bytes = 0, sign_bytes = 0; // allocated sizes (blocks may come from the context's idle list)
    pk::i64 *key = nullptr;     // n (+2 padding) piranha codes, ascending
    pk::u64 *planes = nullptr;  // limbs planes of `stride` words
    signed char *sign = nullptr;
    bool has_meta = false;
    std::vector<int64_t> emin, emax;
    std::vector<uint64_t> egcd;
    uint32_t max_bits = 0;
    bool any_neg = false;
    bool any_zero = false;      // some coefficient is zero (accepted on upload, contributes nothing; results never hold one)
    std::mutex meta_mutex; // has_meta and the fields above are filled lazily (results of pk_mul_dev): guarded
    // multi-device polynomial (created through a multi-device context): rep[d] = full replica on device d (nullptr = not
    // materialised), part[r] = the ordered parts of a partitioned product (part r lives on device r); n = total terms
    bool is_multi = false;
    std::vector<pk_dpoly *> rep, part;
};


namespace pk
{

constexpr uint32_t PK_ALGO_AUTO_FALLBACK = 100u; // internal: pick dense or hash by size

constexpr size_t kScratchBytes = 1 << 16;

struct PkError {
    int code;
    std::string msg;
};

[[noreturn]] inline void fail(int code, std::string msg) { throw PkError{code, std::move(msg)}; }

#define CUDA_CHECK(expr)                                                                                    \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            if (e__ == cudaErrorMemoryAllocation) fail(PK_E_NOMEM, std::string("CUDA out of memory: ") + #expr); \
            fail(PK_E_CUDA, std::string(cudaGetErrorString(e__)) + " at " + #expr);                         \
        }                                                                                                   \
    } while (0)

// count a launch of one of OUR kernels (cub's internal launches are library launches and are not counted)
#define PK_LAUNCH(ctx, kernel, grid, block, smem, ...)                   \
    do {                                                                 \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        ++(ctx)->call_launches;                                          \
        ++(ctx)->total_launches;                                         \
        CUDA_CHECK(cudaGetLastError());                                  \
    } while (0)

constexpr size_t kIdleMinBytes = 256u << 10;

// device memory for this context's stream: an idle block of a fitting size (at most 25 % larger) or the pool
inline void *ctx_alloc(pk_ctx *ctx, size_t &bytes)
{
    if (bytes >= kIdleMinBytes) {
        size_t best = ctx->idle_blocks.size();
        for (size_t i = 0; i < ctx->idle_blocks.size(); ++i) {
            const size_t b = ctx->idle_blocks[i].bytes;
            if (b >= bytes && b <= bytes + bytes / 4 && (best == ctx->idle_blocks.size() || b < ctx->idle_blocks[best].bytes)) best = i;
        }
        if (best != ctx->idle_blocks.size()) {
            void *p = ctx->idle_blocks[best].p;
            bytes = ctx->idle_blocks[best].bytes;
            ctx->idle_bytes -= bytes;
            ctx->idle_blocks[best] = ctx->idle_blocks.back();
            ctx->idle_blocks.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    CUDA_CHECK(cudaMallocFromPoolAsync(&p, bytes, ctx->pool, ctx->stream));
    return p;
}

inline void ctx_release(pk_ctx *ctx, void *p, size_t bytes)
{
    if (!p) return;
    // (keep at most a quarter of the device's memory idle, and at most 64 blocks)
    if (bytes >= kIdleMinBytes && ctx->idle_blocks.size() < 64 && ctx->idle_bytes + bytes <= ctx->total_mem / 4) {
        ctx->idle_blocks.push_back({p, bytes});
        ctx->idle_bytes += bytes;
        return;
    }
    cudaFreeAsync(p, ctx->stream);
}

inline void ctx_drop_idle(pk_ctx *ctx)
{
    for (auto &b : ctx->idle_blocks) cudaFreeAsync(b.p, ctx->stream);
    ctx->idle_blocks.clear();
    ctx->idle_bytes = 0;
}

struct DevBuf { // stream-ordered allocation with scope lifetime
    pk_ctx *ctx;
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf(pk_ctx *c, size_t b) : ctx(c), bytes(b)
    {
        if (bytes) p = ctx_alloc(c, bytes);
    }
    ~DevBuf()
    {
        if (p) ctx_release(ctx, p, bytes);
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <typename T>
    T *as() const
    {
        return static_cast<T *>(p);
    }
};

// dev -> pinned host, `bytes` a multiple of 4, without a copy engine (see k_readback); the caller synchronises the stream
inline void readback(pk_ctx *ctx, void *host_pinned, const void *dev, size_t bytes)
{
    PK_LAUNCH(ctx, k_readback, 1, 64, 0, static_cast<u32 *>(host_pinned), static_cast<const u32 *>(dev), (u32)(bytes / 4));
    --ctx->call_launches; // plumbing, not part of the product's kernel count
    --ctx->total_launches;
}

// two read-backs in one launch
inline void readback2(pk_ctx *ctx, void *h1, const void *d1, size_t b1, void *h2, const void *d2, size_t b2)
{
    PK_LAUNCH(ctx, k_readback2, 1, 64, 0, static_cast<u32 *>(h1), static_cast<const u32 *>(d1), (u32)(b1 / 4), static_cast<u32 *>(h2),
              static_cast<const u32 *>(d2), (u32)(b2 / 4));
    --ctx->call_launches;
    --ctx->total_launches;
}

inline u32 grid_for(u64 n, u32 block) { return (u32)std::max<u64>(1, (n + block - 1) / block); }

inline u32 bit_length(unsigned __int128 x)
{
    u32 b = 0;
    while (x) {
        ++b;
        x >>= 1;
    }
    return b;
}

inline u64 gcd64(u64 a, u64 b)
{
    while (b) {
        u64 t = a % b;
        a = b;
        b = t;
    }
    return a;
}

struct Plan {
    u32 n_vars = 0;
    CodecParams cpA{}, cpB{}, cpR{}; // operand decoders and result encoder
    i64 base_code = 0;               // piranha code of the exponent vector (emin_a + emin_b)
    u64 range = 1;                   // number of tight codes
    u32 code_bits = 1;
    u32 LA = 1, LB = 1;
    u32 prod_bits = 0, cnt_bits = 0;
    u32 cbits = 0, nl = 0, out_limbs = 1;
};


inline void free_dpoly_arrays(pk_ctx *ctx, pk_dpoly *p)
{
    // (a polynomial is freed through the context whose stream last wrote it -- its own; views hold no sizes and free nothing)
    if (p->key && p->key_bytes) ctx_release(ctx, p->key, p->key_bytes);
    if (p->planes && p->planes_bytes) ctx_release(ctx, p->planes, p->planes_bytes);
    if (p->sign && p->sign_bytes) ctx_release(ctx, p->sign, p->sign_bytes);
    p->key = nullptr;
    p->planes = nullptr;
    p->sign = nullptr;
}

struct DPolyGuard {
    pk_ctx *ctx;
    pk_dpoly *p;
    ~DPolyGuard()
    {
        if (p) {
            free_dpoly_arrays(ctx, p);
            delete p;
        }
    }
    pk_dpoly *release()
    {
        pk_dpoly *r = p;
        p = nullptr;
        return r;
    }
    void release_now() // free what the guard holds, now
    {
        if (p) {
            free_dpoly_arrays(ctx, p);
            delete p;
            p = nullptr;
        }
    }
};

inline pk_dpoly *new_dpoly(pk_ctx *ctx, u64 n, u32 n_vars, u32 limbs)
{
    std::unique_ptr<pk_dpoly> p(new pk_dpoly);
    p->n = n;
    p->n_vars = n_vars;
    p->limbs = limbs;
    p->alloc_limbs = limbs;
    p->stride = ((n + 1) & ~1ull) + 2; // even, with slack for the 16-byte TMA granularity
    DPolyGuard g{ctx, p.release()};
    g.p->key_bytes = (n + 2) * sizeof(int64_t);
    g.p->key = static_cast<pk::i64 *>(ctx_alloc(ctx, g.p->key_bytes));
    g.p->planes_bytes = g.p->stride * limbs * sizeof(uint64_t);
    g.p->planes = static_cast<pk::u64 *>(ctx_alloc(ctx, g.p->planes_bytes));
    g.p->sign_bytes = n + 2;
    g.p->sign = static_cast<signed char *>(ctx_alloc(ctx, g.p->sign_bytes));
    return g.release();
}

// The staging arrays of the zone / block paths are bounded by min(term products, code range) terms (~1 GB for
// pearce1) although only the result's size is ever written.  They are kept in the context and reused: allocating and
// freeing blocks of that size per product through the stream-ordered pool made the pool re-map memory every few
// products (measured: 16 ms stalls inside the timed region of bench.py).
struct StagingView {
    pk_dpoly *p;
};
inline StagingView staging_acquire(pk_ctx *ctx, u64 n, u32 n_vars, u32 limbs)
{
    pk_dpoly *s = ctx->staging;
    if (!s || ctx->staging_cap < n || s->alloc_limbs < limbs) {
        if (s) {
            free_dpoly_arrays(ctx, s);
            delete s;
            ctx->staging = nullptr;
            ctx->staging_cap = 0;
        }
        const u64 cap = n + n / 8 + 1024;
        ctx->staging = new_dpoly(ctx, cap, n_vars, limbs);
        ctx->staging_cap = cap;
        s = ctx->staging;
    }
    s->n = n;
    s->n_vars = n_vars;
    return StagingView{s};
}

} // namespace pk
