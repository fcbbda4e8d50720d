// Host side of the C ABI (include/piranha_b200.h): context, operand upload, multiplication planning and
// orchestration, result download.  The CUDA kernels live in pk_kernels.cuh (HBM/L2 accumulators) and
// pk_zone.cuh (shared-memory zones).  There is no CPU fallback: every product is computed by the kernels.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../host/kronecker_array.hpp"
#include "pk_host.cuh"
#include "pk_kernels.cuh"
#include "pk_zone.cuh"
#include "pk_block.cuh"
#include "pk_wide.cuh"

using namespace pk;

namespace
{

void fill_piranha_codec(const pk_ctx *ctx, u32 n_vars, CodecParams &cp)
{
    memset(&cp, 0, sizeof(cp));
    cp.n_vars = n_vars;
    cp.h_min = n_vars ? ctx->h_min[n_vars] : 0;
    cp.h_max = n_vars ? ctx->h_max[n_vars] : 0;
    i64 ps = 1;
    for (u32 v = 0; v < n_vars; ++v) {
        const i64 M = ctx->limits[n_vars][v];
        cp.M[v] = M;
        cp.radix[v] = 2 * M + 1;
        cp.pstride[v] = ps;
        ps = (i64)((u64)ps * (u64)(2 * M + 1));
        cp.step[v] = 1;
        cp.tradix[v] = 1;
    }
}

// Exponent bounds, gcd steps and coefficient width of a polynomial whose terms are in the given arrays (any order;
// `aos`: magnitudes row-major as the caller passed them, else limb planes).
void compute_meta(pk_ctx *ctx, pk_dpoly *p, const i64 *key, const u64 *mag, const signed char *sign, u64 stride, bool aos)
{
    const u32 nv = p->n_vars;
    p->emin.assign(nv, 0);
    p->emax.assign(nv, 0);
    p->egcd.assign(nv, 0);
    if (p->n == 0) {
        p->max_bits = 0;
        p->has_meta = true;
        return;
    }
    MetaDev init;
    for (u32 v = 0; v < PK_MAX_VARS; ++v) {
        init.emin[v] = INT64_MAX;
        init.emax[v] = INT64_MIN;
        init.egcd[v] = 0;
    }
    init.max_bits = 0;
    init.any_neg = 0;
    memcpy(ctx->h_scratch, &init, sizeof(init));
    CUDA_CHECK(cudaMemcpyAsync(ctx->d_meta, ctx->h_scratch, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    CodecParams cp;
    fill_piranha_codec(ctx, nv, cp);
    const u32 grid = std::min<u32>(grid_for(p->n, 256), (u32)ctx->sm_count * 8);
    PK_LAUNCH(ctx, k_meta, grid, 256, 0, key, mag, sign, p->n, p->limbs, stride, cp, ctx->d_meta, aos ? 1u : 0u);
    readback(ctx, ctx->h_scratch, ctx->d_meta, sizeof(MetaDev));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const MetaDev *m = static_cast<const MetaDev *>(ctx->h_scratch);
    for (u32 v = 0; v < nv; ++v) {
        p->emin[v] = m->emin[v];
        p->emax[v] = m->emax[v];
        p->egcd[v] = m->egcd[v];
    }
    p->max_bits = m->max_bits;
    if (m->any_neg & kMetaKeyRange) fail(PK_E_ARGS, "a key is not a Kronecker code of this many variables (outside [h_min, h_max])");
    if (m->any_neg & kMetaBadSign) fail(PK_E_ARGS, "a sign is not -1, 0 or +1");
    p->any_neg = (m->any_neg & kMetaNeg) != 0;
    p->any_zero = (m->any_neg & kMetaZero) != 0;
    p->has_meta = true;
}

void ensure_meta(pk_ctx *ctx, pk_dpoly *p)
{
    // results of pk_mul_dev get their exponent bounds on first use as an operand; two threads (contexts) may share the
    // polynomial, so the lazy fill is serialised
    std::lock_guard<std::mutex> lock(p->meta_mutex);
    if (!p->has_meta) compute_meta(ctx, p, p->key, p->planes, p->sign, p->stride, false);
}

template <typename K, typename V>
void radix_sort_pairs(pk_ctx *ctx, const K *kin, K *kout, const V *vin, V *vout, u64 n, int end_bit)
{
    size_t tmp_bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, ctx->stream));
    DevBuf tmp(ctx, tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, ctx->stream));
}

template <typename K>
void radix_sort_keys(pk_ctx *ctx, const K *kin, K *kout, u64 n, int end_bit)
{
    size_t tmp_bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, kin, kout, (int64_t)n, 0, end_bit, ctx->stream));
    DevBuf tmp(ctx, tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, kin, kout, (int64_t)n, 0, end_bit, ctx->stream));
}

pk_dpoly *upload_impl(pk_ctx *ctx, const pk_poly *p)
{
    if (!p) fail(PK_E_ARGS, "null polynomial");
    if (p->n_vars >= ctx->limits.size()) fail(PK_E_ARGS, "too many variables for the Kronecker codification");
    if (p->limbs < 1 || p->limbs > 4) fail(PK_E_CF_WIDTH, "operand coefficients must have 1 to 4 limbs");
    if (p->n >= (1ull << 32)) fail(PK_E_UNSUPPORTED, "operands of 2^32 terms or more are not supported");
    if (p->n && (!p->key || !p->mag || !p->sign)) fail(PK_E_ARGS, "null term arrays");
    const u64 n = p->n;
    DPolyGuard g{ctx, new_dpoly(ctx, n, p->n_vars, p->limbs)};
    if (n == 0) {
        g.p->has_meta = true;
        g.p->emin.assign(p->n_vars, 0);
        g.p->emax.assign(p->n_vars, 0);
        g.p->egcd.assign(p->n_vars, 0);
        return g.release();
    }
    DevBuf raw_key(ctx, n * 8), raw_mag(ctx, n * p->limbs * 8), raw_sign(ctx, n);
    CUDA_CHECK(cudaMemcpyAsync(raw_key.p, p->key, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(raw_mag.p, p->mag, n * p->limbs * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(raw_sign.p, p->sign, n, cudaMemcpyHostToDevice, ctx->stream));
    // exponent bounds first (on the caller's order), then the canonical order: ascending Kronecker code (the role of
    // fill_term_pointers' per-bucket sort, base_series_multiplier.hpp:126-132; hash_set order itself is unspecified).
    // The sort key is the polynomial's own tight code -- monotone in the Kronecker code and only as wide as the
    // exponent spans need (18 bits for fateman1's operands instead of 64), so the radix sort runs 3 passes, not 8.
    compute_meta(ctx, g.p, raw_key.as<i64>(), raw_mag.as<u64>(), raw_sign.as<signed char>(), 0, true);
    CodecParams cp;
    fill_piranha_codec(ctx, p->n_vars, cp);
    unsigned __int128 range = 1;
    for (u32 v = 0; v < p->n_vars; ++v) {
        const u64 gs = g.p->egcd[v] ? g.p->egcd[v] : 1;
        cp.emin[v] = g.p->emin[v];
        cp.step[v] = (i64)gs;
        cp.stride[v] = (u64)range;
        range *= (u64)(g.p->emax[v] - g.p->emin[v]) / gs + 1;
        if (range >= ((unsigned __int128)1 << 62)) fail(PK_E_UNSUPPORTED, "exponent spans too wide to re-encode (code range of 2^62 or more)");
    }
    DevBuf fk(ctx, n * 8), fk2(ctx, n * 8), idx(ctx, n * 4), idx2(ctx, n * 4);
    {
        const PrepareArgs one{raw_key.as<i64>(), raw_sign.as<signed char>(), n, fk.as<u64>(), nullptr, nullptr, nullptr}, none{};
        PK_LAUNCH(ctx, k_prepare, grid_for(n, 256), 256, 0, one, none, cp, cp, grid_for(n, 256));
    }
    PK_LAUNCH(ctx, k_iota32, grid_for(n, 256), 256, 0, idx.as<u32>(), n);
    radix_sort_pairs(ctx, fk.as<u64>(), fk2.as<u64>(), idx.as<u32>(), idx2.as<u32>(), n, (int)std::max<u32>(1, bit_length(range)));
    CUDA_CHECK(cudaMemsetAsync(&ctx->d_meta->any_neg, 0, 4, ctx->stream));
    PK_LAUNCH(ctx, k_gather_terms, grid_for(n, 256), 256, 0, idx2.as<u32>(), raw_key.as<i64>(), raw_mag.as<u64>(),
              raw_sign.as<signed char>(), g.p->key, g.p->planes, g.p->sign, n, p->limbs, g.p->stride, &ctx->d_meta->any_neg);
    // the upload returns with the polynomial complete (other contexts / streams may use it) and its keys checked
    readback(ctx, ctx->h_scratch, &ctx->d_meta->any_neg, 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (*static_cast<const u32 *>(ctx->h_scratch) & kMetaDuplicate) fail(PK_E_ARGS, "two terms with the same key");
    return g.release();
}

PinnedBuf *pinned_acquire(pk_ctx *ctx, size_t bytes)
{
    PinnedBuf *best = nullptr;
    for (auto &b : ctx->pinned)
        if (!b.in_use && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
    if (!best) {
        for (auto &b : ctx->pinned)
            if (!b.in_use && b.ptr) { // recycle the slot of a too-small buffer
                cudaFreeHost(b.ptr);
                b.ptr = nullptr;
                b.bytes = 0;
            }
        ctx->pinned.erase(std::remove_if(ctx->pinned.begin(), ctx->pinned.end(),
                                         [](const PinnedBuf &b) { return !b.in_use && !b.ptr; }),
                          ctx->pinned.end());
        PinnedBuf nb;
        nb.bytes = std::max<size_t>(bytes, 4096);
        nb.bytes += nb.bytes / 8; // headroom so similar-sized results reuse it
        CUDA_CHECK(cudaHostAlloc(&nb.ptr, nb.bytes, cudaHostAllocPortable)); // any device may copy into it
        ctx->pinned.push_back(nb);
        best = &ctx->pinned.back();
    }
    best->in_use = true;
    return best;
}

void download_impl(pk_ctx *ctx, const pk_dpoly *p, pk_result *out)
{
    memset(out, 0, sizeof(*out));
    out->n = p->n;
    out->n_vars = p->n_vars;
    out->limbs = p->limbs;
    if (p->n == 0) return;
    const u64 n = p->n;
    const size_t key_b = (n * 8 + 63) & ~63ull, mag_b = (n * p->limbs * 8 + 63) & ~63ull, sign_b = (n + 63) & ~63ull;
    PinnedBuf *pb = pinned_acquire(ctx, key_b + mag_b + sign_b);
    char *base = static_cast<char *>(pb->ptr);
    out->key = reinterpret_cast<int64_t *>(base);
    out->mag = reinterpret_cast<uint64_t *>(base + key_b);
    out->sign = reinterpret_cast<int8_t *>(base + key_b + mag_b);
    out->owner = pb->ptr;
    try {
        CUDA_CHECK(cudaMemcpyAsync(out->key, p->key, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (p->limbs == 1) {
            CUDA_CHECK(cudaMemcpyAsync(out->mag, p->planes, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            DevBuf aos(ctx, n * p->limbs * 8);
            PK_LAUNCH(ctx, k_planes_to_aos, grid_for(n, 256), 256, 0, p->planes, aos.as<u64>(), n, p->limbs, p->stride, p->limbs);
            CUDA_CHECK(cudaMemcpyAsync(out->mag, aos.p, n * p->limbs * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CUDA_CHECK(cudaMemcpyAsync(out->sign, p->sign, n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        pb->in_use = false;
        memset(out, 0, sizeof(*out));
        throw;
    }
}

// the terms grouped by ((u64)key >> shift) & (2^bits - 1), ascending keys inside a group (pk_download_grouped)
void download_grouped_impl(pk_ctx *ctx, const pk_dpoly *p, u32 shift, u32 bits, pk_result *out, u32 n_chunks, pk_progress_fn fn, void *user,
                           uint64_t *group_offsets)
{
    memset(out, 0, sizeof(*out));
    out->n = p->n;
    out->n_vars = p->n_vars;
    out->limbs = p->limbs;
    if (p->n == 0) {
        if (group_offsets) std::fill(group_offsets, group_offsets + (1u << bits) + 1, 0ull);
        return;
    }
    const u64 n = p->n;
    if (n >= (1ull << 32)) fail(PK_E_UNSUPPORTED, "pk_download_grouped: 2^32 terms or more");
    const u32 L = p->limbs;
    DevBuf g0(ctx, n * 4), g1(ctx, n * 4), i0(ctx, n * 4), i1(ctx, n * 4);
    PK_LAUNCH(ctx, k_group_of, grid_for(n, 256), 256, 0, p->key, n, shift, (1u << bits) - 1u, g0.as<u32>(), i0.as<u32>());
    radix_sort_pairs(ctx, g0.as<u32>(), g1.as<u32>(), i0.as<u32>(), i1.as<u32>(), n, (int)bits); // (stable: keys stay ascending inside a group)
    DevBuf goff(ctx, group_offsets ? ((size_t)(1u << bits) + 1) * 8 : 0);
    if (group_offsets) PK_LAUNCH(ctx, k_group_offsets, grid_for((1u << bits) + 1, 256), 256, 0, g1.as<u32>(), n, 1u << bits, goff.as<u64>());
    DevBuf dk(ctx, n * 8), dm(ctx, n * L * 8), dsg(ctx, n);
    PK_LAUNCH(ctx, k_group_gather, grid_for(n, 256), 256, 0, i1.as<u32>(), p->key, p->planes, p->sign, n, L, p->stride, dk.as<i64>(),
              dm.as<u64>(), dsg.as<signed char>());
    const size_t key_b = (n * 8 + 63) & ~63ull, mag_b = (n * L * 8 + 63) & ~63ull, sign_b = (n + 63) & ~63ull;
    PinnedBuf *pb = pinned_acquire(ctx, key_b + mag_b + sign_b);
    char *base = static_cast<char *>(pb->ptr);
    out->key = reinterpret_cast<int64_t *>(base);
    out->mag = reinterpret_cast<uint64_t *>(base + key_b);
    out->sign = reinterpret_cast<int8_t *>(base + key_b + mag_b);
    out->owner = pb->ptr;
    try {
        n_chunks = (u32)std::min<u64>(std::max<u32>(1, n_chunks), n);
        while (ctx->chunk_events.size() < n_chunks) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); // (not cudaEventBlockingSync: measured ~1.2 ms per wake-up on this platform)
            ctx->chunk_events.push_back(e);
        }
        if (group_offsets) { // (a small copy into the caller's memory: complete when the call returns)
            CUDA_CHECK(cudaMemcpyAsync(group_offsets, goff.p, ((size_t)(1u << bits) + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        // all pieces are queued at once (the copies run back to back); the host then follows the events
        for (u32 c = 0; c < n_chunks; ++c) {
            const u64 lo = n * c / n_chunks, cnt = n * (c + 1) / n_chunks - lo;
            CUDA_CHECK(cudaMemcpyAsync(out->key + lo, dk.as<i64>() + lo, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(out->mag + lo * L, dm.as<u64>() + lo * L, cnt * L * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(out->sign + lo, dsg.as<signed char>() + lo, cnt, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaEventRecord(ctx->chunk_events[c], ctx->stream));
        }
        for (u32 c = 0; c < n_chunks; ++c) {
            CUDA_CHECK(cudaEventSynchronize(ctx->chunk_events[c]));
            if (fn) fn(user, out, n * (c + 1) / n_chunks);
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        cudaStreamSynchronize(ctx->stream); // (no copy may still be writing into the buffer when it is handed back)
        pb->in_use = false;
        memset(out, 0, sizeof(*out));
        throw;
    }
}

// ------------------------------------------------------------------------------------------------ multiply

Plan make_plan(pk_ctx *ctx, const pk_dpoly *A, const pk_dpoly *B, const pk_opts &o)
{
    Plan pl;
    const u32 nv = A->n_vars;
    pl.n_vars = nv;
    fill_piranha_codec(ctx, nv, pl.cpA);
    fill_piranha_codec(ctx, nv, pl.cpB);
    fill_piranha_codec(ctx, nv, pl.cpR);
    pl.cpA.trunc_mode = pl.cpB.trunc_mode = (u32)o.trunc_mode;
    unsigned __int128 range = 1;
    __int128 base = 0;
    for (u32 v = 0; v < nv; ++v) { // check_bounds first, for every variable (polynomial.hpp:1189-1193): its error wins
        const __int128 lo = (__int128)A->emin[v] + B->emin[v], hi = (__int128)A->emax[v] + B->emax[v];
        const i64 M = ctx->limits[nv][v];
        if (lo < -(__int128)M || hi > (__int128)M) fail(PK_E_KEY_RANGE, "Kronecker monomial components are out of bounds");
    }
    for (u32 v = 0; v < nv; ++v) {
        const __int128 lo = (__int128)A->emin[v] + B->emin[v];
        u64 g = gcd64(A->egcd[v], B->egcd[v]);
        if (g == 0) g = 1;
        const u64 span = (u64)(A->emax[v] - A->emin[v]) + (u64)(B->emax[v] - B->emin[v]);
        u64 tr = span / g + 1;
        if (v == 0 && nv >= 2 && ctx->block_tuning.radix0_residue) // (bank-conflict experiment: any radix >= tr is exact)
            tr += (ctx->block_tuning.radix0_residue + 32 - tr % 32) % 32;
        pl.cpA.emin[v] = A->emin[v];
        pl.cpB.emin[v] = B->emin[v];
        pl.cpR.emin[v] = (i64)lo;
        pl.cpA.step[v] = pl.cpB.step[v] = pl.cpR.step[v] = (i64)g;
        pl.cpA.stride[v] = pl.cpB.stride[v] = pl.cpR.stride[v] = (u64)range;
        pl.cpA.tradix[v] = pl.cpB.tradix[v] = pl.cpR.tradix[v] = tr;
        const bool counts = o.trunc_mode == 1 || (o.trunc_mode == 2 && o.degree_mask && o.degree_mask[v]);
        pl.cpA.mask[v] = pl.cpB.mask[v] = counts ? 1 : 0;
        range *= tr;
        if (range >= ((unsigned __int128)1 << 62)) fail(PK_E_UNSUPPORTED, "exponent spans too wide to re-encode (code range of 2^62 or more)");
        base += lo * (__int128)pl.cpR.pstride[v];
    }
    pl.base_code = (i64)base;
    pl.range = (u64)range;
    pl.code_bits = std::max<u32>(1, bit_length(range));
    // coefficient widths: |a_i * b_j| < 2^(bitsA + bitsB); at most min(n1, n2) products land on one key
    pl.LA = std::max<u32>(1, (A->max_bits + 63) / 64);
    pl.LB = std::max<u32>(1, (B->max_bits + 63) / 64);
    if (pl.LA > 2 || pl.LB > 2) fail(PK_E_CF_WIDTH, "operand coefficient wider than 2 limbs");
    pl.prod_bits = std::max<u32>(1, A->max_bits + B->max_bits);
    pl.cnt_bits = bit_length(std::min(A->n, B->n));
    if (pl.cnt_bits > 31) fail(PK_E_UNSUPPORTED, "operands of 2^31 terms or more are not supported");
    pl.cbits = std::min<u32>(62, 63 - pl.cnt_bits);
    pl.nl = (pl.prod_bits + pl.cbits - 1) / pl.cbits;
    pl.out_limbs = (pl.prod_bits + pl.cnt_bits + 63) / 64;
    if (pl.out_limbs > (u32)kMaxAccLimbs - 1) fail(PK_E_CF_WIDTH, "coefficient bound exceeds the accumulator width");
    return pl;
}

template <int MODE, bool FILTER>
void launch_mul_global(pk_ctx *ctx, const Plan &pl, const MulArgs &args)
{
    const u32 R = (pl.LA + pl.LB == 2) ? 8 : 4;
    const u32 TI = (kMulThreads / 32) * R;
    const u64 tiles = ((args.nA + TI - 1) / TI) * ((args.nB + kTileJ - 1) / kTileJ);
    if (tiles >= (1ull << 31)) fail(PK_E_UNSUPPORTED, "product too large for the HBM-accumulator kernels (2^31 tiles or more)");
    const u32 grid = (u32)tiles;
    if (pl.LA == 1 && pl.LB == 1) PK_LAUNCH(ctx, (k_mul_global<MODE, 1, 1, FILTER>), grid, kMulThreads, 0, args);
    else if (pl.LA == 2 && pl.LB == 1) PK_LAUNCH(ctx, (k_mul_global<MODE, 2, 1, FILTER>), grid, kMulThreads, 0, args);
    else if (pl.LA == 1 && pl.LB == 2) PK_LAUNCH(ctx, (k_mul_global<MODE, 1, 2, FILTER>), grid, kMulThreads, 0, args);
    else PK_LAUNCH(ctx, (k_mul_global<MODE, 2, 2, FILTER>), grid, kMulThreads, 0, args);
}

pk_dpoly *wide_multiply(pk_ctx *ctx, pk_dpoly *A, pk_dpoly *B, const pk_opts *opts_in);

pk_dpoly *mul_impl(pk_ctx *ctx, pk_dpoly *a, pk_dpoly *b, const pk_opts *opts_in)
{
    pk_opts o{};
    if (opts_in) {
        if (opts_in->struct_size != sizeof(pk_opts)) fail(PK_E_ARGS, "pk_opts.struct_size mismatch");
        o = *opts_in;
    }
    if (!a || !b) fail(PK_E_ARGS, "null operand");
    // base_series_multiplier.hpp:365-367: "incompatible arguments sets"
    if (a->n_vars != b->n_vars) fail(PK_E_ARGS, "incompatible arguments sets");
    if (o.trunc_mode < 0 || o.trunc_mode > 2 || o.reserved != 0 || o.algo > PK_ALGO_BLOCK) fail(PK_E_ARGS, "bad options");
    if (o.trunc_mode == 2 && a->n_vars && !o.degree_mask) fail(PK_E_ARGS, "partial truncation needs degree_mask");
    u32 P = o.part_count ? o.part_count : 1;
    if (o.part_index >= P || P > 4096) fail(PK_E_ARGS, "part_index / part_count out of range");
    const u32 nv = a->n_vars;

    ctx->call_launches = 0;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_valid = false;
    ctx->aux_valid = false;
    // larger series first, base_series_multiplier.hpp:368-372
    pk_dpoly *A = a, *B = b;
    if (A->n < B->n) std::swap(A, B);
    ctx->stats.n1 = A->n;
    ctx->stats.n2 = B->n;
    if (A->n == 0 || B->n == 0) { // polynomial.hpp:1654-1656: empty series, symbol set preserved
        pk_dpoly *r = new_dpoly(ctx, 0, nv, 1);
        r->has_meta = true;
        r->emin.assign(nv, 0);
        r->emax.assign(nv, 0);
        r->egcd.assign(nv, 0);
        return r;
    }
    ensure_meta(ctx, A);
    ensure_meta(ctx, B);
    // operands of three or four limbs: 128-bit halves through the same kernels, summed with shifts (pk_wide.cuh)
    if (A->max_bits > 128 || B->max_bits > 128) return wide_multiply(ctx, A, B, opts_in);
    Plan pl = make_plan(ctx, A, B, o);
    const bool trunc = o.trunc_mode != 0;
    const u64 nA = A->n, nB = B->n;

    CUDA_CHECK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(MulCounters), ctx->stream));

    // ---- operand preparation
    DevBuf keyA(ctx, (nA + 2) * 8), keyB(ctx, (nB + 2) * 8);
    DevBuf degA(ctx, trunc ? nA * 8 : 0), degB(ctx, trunc ? nB * 8 : 0);
    // the block path's packed operand records are written by the same pass (untruncated products; the truncated
    // form of the record needs the operand's smallest degree first)
    BlockPlan bp = block_plan(ctx, pl, nA, nB);
    const bool prepack = bp.ok && !trunc && (o.algo == PK_ALGO_AUTO || o.algo == PK_ALGO_BLOCK);
    DevBuf packA(ctx, prepack ? (nA + 1) * 16 : 0), packB(ctx, prepack ? (nB + 1) * 16 : 0);
    {
        const PrepareArgs pa{A->key, A->sign, nA, keyA.as<u64>(), degA.as<i64>(), A->planes, packA.as<uint4>()};
        const PrepareArgs pb{B->key, B->sign, nB, keyB.as<u64>(), degB.as<i64>(), B->planes, packB.as<uint4>()};
        const u32 blocksA = grid_for(nA, 256);
        PK_LAUNCH(ctx, k_prepare, blocksA + grid_for(nB, 256), 256, 0, pa, pb, pl.cpA, pl.cpB, blocksA);
    }
    const u64 *kB = keyB.as<u64>();
    const u64 *mB = B->planes;
    u64 strideB = B->stride;
    std::unique_ptr<DevBuf> keyB2, magB2, slbuf, degBs;
    // The HBM-accumulator kernels truncate per row: second operand stable-sorted by degree (polynomial.hpp:1427-1442),
    // then the skip limits.  Only built when those kernels run (the block path prunes per tile instead).
    auto prepare_degree_sorted = [&]() {
        DevBuf fd(ctx, nB * 8), idx(ctx, nB * 4), idx2(ctx, nB * 4);
        degBs.reset(new DevBuf(ctx, nB * 8));
        PK_LAUNCH(ctx, k_flip_deg, grid_for(nB, 256), 256, 0, degB.as<i64>(), fd.as<u64>(), nB);
        PK_LAUNCH(ctx, k_iota32, grid_for(nB, 256), 256, 0, idx.as<u32>(), nB);
        radix_sort_pairs(ctx, fd.as<u64>(), degBs->as<u64>(), idx.as<u32>(), idx2.as<u32>(), nB, 64);
        keyB2.reset(new DevBuf(ctx, (nB + 2) * 8));
        const u64 st2 = ((nB + 1) & ~1ull) + 2;
        magB2.reset(new DevBuf(ctx, st2 * pl.LB * 8));
        PK_LAUNCH(ctx, k_permute_operand, grid_for(nB, 256), 256, 0, idx2.as<u32>(), keyB.as<u64>(), B->planes,
                  keyB2->as<u64>(), magB2->as<u64>(), nB, pl.LB, B->stride, st2);
        kB = keyB2->as<u64>();
        mB = magB2->as<u64>();
        strideB = st2;
        slbuf.reset(new DevBuf(ctx, nA * 4));
        PK_LAUNCH(ctx, k_skip_limits, grid_for(nA, 256), 256, 0, degA.as<i64>(), nA, degBs->as<u64>(), nB, o.max_degree,
                  slbuf->as<u32>(), ctx->d_ctr);
    };

    // ---- strategy
    // a partitioned product uses the block geometry only when every part gets many whole blocks; a function of the operands
    // and P only, so all ranks agree
    if (bp.ok && P > 1 && pl.range / bp.S < 256ull * P) bp.ok = false;
    const unsigned __int128 w_bound = (unsigned __int128)nA * nB;
    const bool filter = trunc || P > 1;
    u32 algo = o.algo;
    pk_dpoly *result = nullptr;
    if ((algo == PK_ALGO_AUTO && w_bound >= (1u << 16) && bp.ok) || algo == PK_ALGO_BLOCK) {
        // the block path cuts the output partition itself, on the device, from its per-unit product counts
        if (bp.ok)
            result = block_multiply(ctx, pl, bp, A, B, keyA.as<u64>(), keyB.as<u64>(), o.part_index, P, nv, trunc ? degA.as<i64>() : nullptr,
                                    trunc ? degB.as<i64>() : nullptr, o.max_degree, prepack ? packA.as<uint4>() : nullptr,
                                    prepack ? packB.as<uint4>() : nullptr);
        algo = result ? PK_ALGO_BLOCK : PK_ALGO_ZONE; // no block geometry for these operands (or tiny groups): generic zones
    }
    // ---- output partition [lo, hi) in tight-code space for the generic paths
    u64 lo = 0, hi = pl.range;
    if (!result && P > 1) {
        // stratified sample of pair sums: 512 x 512 pairs resolve the cuts to ~0.2 % of the work for any P <= 64
        const u32 SA = (u32)std::min<u64>(nA, 512), SB = (u32)std::min<u64>(nB, 512);
        const u64 S = (u64)SA * SB;
        DevBuf sums(ctx, S * 8), sorted(ctx, S * 8), dcuts(ctx, ((size_t)P + 1) * 8);
        const i64 *dB = degB.as<i64>(); // (both operands are still in code order here: keys and degrees match)
        PK_LAUNCH(ctx, k_sample_sums, grid_for(S, 256), 256, 0, keyA.as<u64>(), nA, kB, nB, SA, SB, degA.as<i64>(), dB,
                  o.max_degree, trunc ? 1 : 0, sums.as<u64>(), ctx->d_ctr);
        // (excluded pairs carry the all-ones sentinel; sums are below the code range, so sorting on one bit more than
        // the range has still puts the sentinels last -- 4 radix passes instead of 8 for a 2^30 range)
        radix_sort_keys(ctx, sums.as<u64>(), sorted.as<u64>(), S, (int)std::min<u32>(64, bit_length(pl.range) + 1));
        PK_LAUNCH(ctx, k_pick_cuts, 1, 64, 0, sorted.as<u64>(), ctx->d_ctr, P, dcuts.as<u64>());
        u64 *cuts = reinterpret_cast<u64 *>(static_cast<char *>(ctx->h_scratch) + 1024);
        readback(ctx, cuts, dcuts.p, ((size_t)P + 1) * 8);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cuts[0] = 0;
        cuts[P] = pl.range;
        for (u32 p = 1; p <= P; ++p) cuts[p] = std::max(cuts[p], cuts[p - 1]);
        lo = cuts[o.part_index];
        hi = cuts[o.part_index + 1];
    }
    const u64 sub_range = hi - lo;
    if (!result && algo == PK_ALGO_AUTO) algo = choose_algo(ctx, pl, nA, nB, sub_range, trunc);
    if (trunc && !result) prepare_degree_sorted();
    if (!result && algo == PK_ALGO_ZONE) {
        result = zone_multiply(ctx, pl, A, B, keyA.as<u64>(), kB, mB, strideB, trunc ? slbuf->as<u32>() : nullptr, lo, hi,
                               trunc, nv);
        if (!result) algo = PK_ALGO_AUTO_FALLBACK; // zone path declined (configuration outside its envelope)
    }
    if (!result && algo == PK_ALGO_AUTO_FALLBACK) {
        const unsigned __int128 dense_bytes = (unsigned __int128)sub_range * pl.nl * 8;
        algo = (sub_range <= 8 * w_bound && dense_bytes <= ((unsigned __int128)16 << 30)) ? PK_ALGO_DENSE : PK_ALGO_HASH;
    }
    if (!result) {
        MulArgs m{};
        m.keyA = keyA.as<u64>();
        m.magA = A->planes;
        m.strideA = A->stride;
        m.nA = nA;
        m.keyB = kB;
        m.magB = mB;
        m.strideB = strideB;
        m.nB = nB;
        m.sl = trunc ? slbuf->as<u32>() : nullptr;
        m.lo = lo;
        m.hi = hi;
        m.nl = pl.nl;
        m.cbits = pl.cbits;
        m.keys_sorted = trunc ? 0 : 1;
        m.ctr = ctx->d_ctr;
        FinalArgs f{};
        f.nl = pl.nl;
        f.cbits = pl.cbits;
        f.out_limbs = pl.out_limbs;
        f.cp = pl.cpR;
        f.base_code = pl.base_code;
        f.ctr = ctx->d_ctr;
        const size_t total_b = ctx->total_mem; // allocation failures surface as PK_E_NOMEM from cudaMallocAsync
        MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
        if (algo == PK_ALGO_DENSE) {
            const unsigned __int128 bytes = (unsigned __int128)std::max<u64>(sub_range, 1) * pl.nl * 8;
            if (bytes > (unsigned __int128)total_b) fail(PK_E_NOMEM, "dense accumulator does not fit in device memory");
            DevBuf acc(ctx, (size_t)bytes);
            CUDA_CHECK(cudaMemsetAsync(acc.p, 0, (size_t)bytes, ctx->stream));
            m.acc = acc.as<u64>() - lo * pl.nl; // indexed by absolute code
            CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
            if (filter) launch_mul_global<0, true>(ctx, pl, m);
            else launch_mul_global<0, false>(ctx, pl, m);
            CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
            // normalise + ordered compaction
            f.acc = m.acc;
            f.n_slots = sub_range;
            f.code0 = lo;
            const u32 nblocks = grid_for(sub_range, kFinThreads * kFinItems);
            DevBuf counts(ctx, (size_t)nblocks * 4), offsets(ctx, (size_t)nblocks * 8);
            f.block_counts = counts.as<u32>();
            f.block_offsets = offsets.as<u64>();
            PK_LAUNCH(ctx, k_dense_count, nblocks, kFinThreads, 0, f);
            PK_LAUNCH(ctx, k_scan_counts, 1, 1024, 0, counts.as<u32>(), offsets.as<u64>(), (u64)nblocks, ctx->d_ctr);
            readback(ctx, hc, ctx->d_ctr, sizeof(MulCounters));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            const u64 n_out = hc->n_out;
            DPolyGuard g{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
            if (n_out) {
                f.okey = g.p->key;
                f.oplanes = g.p->planes;
                f.ostride = g.p->stride;
                f.osign = g.p->sign;
                PK_LAUNCH(ctx, k_dense_write, nblocks, kFinThreads, 0, f);
            }
            result = g.release();
            ctx->stats.table_slots = sub_range;
        } else {
            unsigned __int128 distinct = std::min<unsigned __int128>(w_bound, sub_range);
            u64 cap = 1024;
            while ((unsigned __int128)cap < 2 * distinct) cap <<= 1;
            u32 rec_words = 2;
            while (rec_words < pl.nl + 1) rec_words <<= 1;
            const unsigned __int128 bytes = (unsigned __int128)cap * rec_words * 8;
            if (bytes > (unsigned __int128)total_b / 10 * 9) fail(PK_E_NOMEM, "hash table does not fit in device memory");
            if (cap > (1ull << 32)) fail(PK_E_NOMEM, "hash table too large");
            DevBuf tab(ctx, (size_t)bytes);
            PK_LAUNCH(ctx, k_hash_init, grid_for((u64)cap * rec_words, 256), 256, 0, tab.as<u64>(), (u64)cap * rec_words,
                      rec_words);
            m.acc = tab.as<u64>();
            m.cap_mask = cap - 1;
            u32 lg = 0;
            while ((1ull << lg) < cap) ++lg;
            m.hash_shift = 64 - lg;
            m.rec_words = rec_words;
            CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
            if (filter) launch_mul_global<1, true>(ctx, pl, m);
            else launch_mul_global<1, false>(ctx, pl, m);
            CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
            f.acc = tab.as<u64>();
            f.n_slots = cap;
            f.rec_words = rec_words;
            // the number of occupied slots is only known on the device: collect into buffers sized by a first count
            DevBuf codes(ctx, (size_t)std::min<unsigned __int128>(distinct, cap) * 8),
                slots(ctx, (size_t)std::min<unsigned __int128>(distinct, cap) * 4);
            const u32 cgrid = std::min<u32>(grid_for(cap, 256), (u32)ctx->sm_count * 16);
            PK_LAUNCH(ctx, k_hash_collect, cgrid, 256, 0, f, codes.as<u64>(), slots.as<u32>());
            readback(ctx, hc, ctx->d_ctr, sizeof(MulCounters));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            const u64 n_out = hc->n_out;
            DPolyGuard g{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
            if (n_out) {
                DevBuf codes2(ctx, n_out * 8), slots2(ctx, n_out * 4);
                radix_sort_pairs(ctx, codes.as<u64>(), codes2.as<u64>(), slots.as<u32>(), slots2.as<u32>(), n_out,
                                 (int)pl.code_bits);
                f.okey = g.p->key;
                f.oplanes = g.p->planes;
                f.ostride = g.p->stride;
                f.osign = g.p->sign;
                PK_LAUNCH(ctx, k_hash_write, grid_for(n_out, 256), 256, 0, f, codes2.as<u64>(), slots2.as<u32>(), n_out);
            }
            result = g.release();
            ctx->stats.table_slots = cap;
        }
        ctx->stats.acc_lanes = pl.nl;
        ctx->stats.chunk_bits = pl.cbits;
    }
    DPolyGuard rg{ctx, result};
    CUDA_CHECK(cudaEventRecord(ctx->ev_end, ctx->stream));
    MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
    // (the product returns only when the stream is idle: measured, a product that returned with its last copy kernel in
    // flight made the next product's stream-ordered allocations of ~1 GB take 16 ms instead of reusing the freed blocks)
    // (the block path read the counters after its pairing kernel, into the same host words; only the gather follows it)
    if (algo != PK_ALGO_BLOCK) readback(ctx, hc, ctx->d_ctr, sizeof(MulCounters));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->ev_valid = true;
    if (hc->flags & kFlagDegreeOverflow) fail(PK_E_DEGREE, "overflow in the truncation degree arithmetic");
    if (hc->flags & kFlagTableFull) fail(PK_E_NOMEM, "accumulator table overflow");
    if (hc->flags & kFlagWidth) fail(PK_E_CF_WIDTH, "coefficient exceeded the proven bound (internal error)");
    rg.p->max_bits = hc->max_bits;
    // trim limbs that the actual coefficients never reach (planes are independent, so this is free)
    rg.p->limbs = std::max<u32>(1, std::min<u32>(rg.p->limbs, (hc->max_bits + 63) / 64));
    ctx->stats.term_products = filter ? hc->pairs : (u64)w_bound;
    ctx->stats.result_terms = rg.p->n;
    ctx->stats.code_range = pl.range;
    ctx->stats.algo = algo;
    ctx->stats.result_limbs = rg.p->limbs;
    ctx->stats.kernel_launches = ctx->call_launches;
    return rg.release();
}

// a * b for operands of up to four limbs (see pk_wide.cuh)
pk_dpoly *wide_multiply(pk_ctx *ctx, pk_dpoly *A, pk_dpoly *B, const pk_opts *opts_in)
{
    if (A->max_bits > 256 || B->max_bits > 256) fail(PK_E_CF_WIDTH, "operand coefficient wider than 4 limbs");
    const u32 nv = A->n_vars;
    // 128-bit halves as views of the operand's arrays (a half that is all zero contributes nothing and is skipped)
    struct Half {
        pk_dpoly v;
        u32 shift = 0;
        bool used = false;
    };
    Half hA[2], hB[2];
    auto make_views = [&](pk_dpoly *P, Half (&h)[2]) {
        for (u32 k = 0; k < 2; ++k) {
            if (P->limbs <= 2 * k) continue;
            pk_dpoly &v = h[k].v;
            v.n = P->n;
            v.stride = P->stride;
            v.n_vars = P->n_vars;
            v.limbs = std::min<u32>(2, P->limbs - 2 * k);
            v.alloc_limbs = v.limbs;
            v.key = P->key;
            v.planes = P->planes + (u64)2 * k * P->stride;
            v.sign = P->sign;
            v.has_meta = false;
            ensure_meta(ctx, &v);
            h[k].shift = 2 * k;
            h[k].used = v.max_bits != 0;
        }
    };
    make_views(A, hA);
    make_views(B, hB);
    std::vector<pk_dpoly *> parts;
    std::vector<u32> shifts;
    struct PartsGuard {
        pk_ctx *ctx;
        std::vector<pk_dpoly *> &v;
        ~PartsGuard()
        {
            for (pk_dpoly *p : v)
                if (p) {
                    free_dpoly_arrays(ctx, p);
                    delete p;
                }
        }
    } guard{ctx, parts};
    pk_stats first{};
    float kernel_ms = 0.f;
    for (u32 i = 0; i < 2; ++i)
        for (u32 j = 0; j < 2; ++j) {
            if (!hA[i].used || !hB[j].used) continue;
            parts.push_back(mul_impl(ctx, &hA[i].v, &hB[j].v, opts_in));
            shifts.push_back(hA[i].shift + hB[j].shift);
            pk_stats st{};
            ctx->stats.total_launches = ctx->total_launches;
            if (parts.size() == 1) first = ctx->stats;
            float ms = 0.f;
            if (ctx->ev_valid && cudaEventSynchronize(ctx->ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev_mul0, ctx->ev_mul1) == cudaSuccess)
                kernel_ms += ms;
            (void)st;
        }
    // restore the views' borrowed pointers before they go out of scope (they own nothing)
    u64 total = 0;
    for (pk_dpoly *p : parts) total += p->n;
    const u32 cnt_bits = bit_length(std::min(A->n, B->n));
    const u32 out_limbs = std::max<u32>(1, (A->max_bits + B->max_bits + cnt_bits + 63) / 64);
    if (out_limbs + 1 > kWideLimbs) fail(PK_E_CF_WIDTH, "coefficient bound exceeds the accumulator width");
    if (total >= (1ull << 30)) fail(PK_E_NOMEM, "wide product too large for the reduce pass");
    if (total == 0) {
        pk_dpoly *r = new_dpoly(ctx, 0, nv, 1);
        r->has_meta = true;
        r->emin.assign(nv, 0);
        r->emax.assign(nv, 0);
        r->egcd.assign(nv, 0);
        return r;
    }
    WideSrc w{};
    w.n_parts = (u32)parts.size();
    w.h_min = nv ? ctx->h_min[nv] : 0;
    u64 off = 0;
    for (u32 s = 0; s < w.n_parts; ++s) {
        w.key[s] = parts[s]->key;
        w.planes[s] = parts[s]->planes;
        w.sign[s] = parts[s]->sign;
        w.stride[s] = parts[s]->stride;
        w.limbs[s] = parts[s]->limbs;
        w.shift[s] = shifts[s];
        w.off[s] = off;
        off += parts[s]->n;
    }
    for (u32 s = w.n_parts; s <= kWideMaxParts; ++s) w.off[s] = off;
    DevBuf k0(ctx, total * 8), k1(ctx, total * 8), s0(ctx, total * 4), s1(ctx, total * 4), head(ctx, total * 4), upos(ctx, total * 4);
    PK_LAUNCH(ctx, k_wide_collect, grid_for(total, 256), 256, 0, w, k0.as<u64>(), s0.as<u32>());
    const unsigned __int128 span = nv ? (unsigned __int128)((__int128)ctx->h_max[nv] - ctx->h_min[nv]) : 0;
    radix_sort_pairs(ctx, k0.as<u64>(), k1.as<u64>(), s0.as<u32>(), s1.as<u32>(), total, (int)std::max<u32>(1, bit_length(span)));
    PK_LAUNCH(ctx, k_wide_heads, grid_for(total, 256), 256, 0, k1.as<u64>(), total, head.as<u32>());
    auto exclusive_sum = [&](const u32 *in, u32 *out, u64 n) {
        size_t tmp_bytes = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, ctx->stream));
        DevBuf tmp(ctx, tmp_bytes);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int)n, ctx->stream));
    };
    exclusive_sum(head.as<u32>(), upos.as<u32>(), total);
    u32 *h2 = reinterpret_cast<u32 *>(static_cast<char *>(ctx->h_scratch) + 3072);
    auto last_plus = [&](const u32 *scan, const u32 *flag, u64 n) -> u64 { // exclusive scan value + flag of the last element
        readback(ctx, h2, scan + (n - 1), 4);
        readback(ctx, h2 + 1, flag + (n - 1), 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        return (u64)h2[0] + h2[1];
    };
    const u64 n_unique = last_plus(upos.as<u32>(), head.as<u32>(), total);
    DevBuf tkey(ctx, n_unique * 8), tmag(ctx, n_unique * kWideLimbs * 8), tsign(ctx, n_unique), tnz(ctx, n_unique * 4), pos(ctx, n_unique * 4);
    CUDA_CHECK(cudaMemsetAsync(&ctx->d_ctr->max_bits, 0, 4, ctx->stream));
    PK_LAUNCH(ctx, k_wide_reduce, grid_for(total, 256), 256, 0, w, k1.as<u64>(), s1.as<u32>(), head.as<u32>(), upos.as<u32>(), total,
              tkey.as<u64>(), tmag.as<u64>(), tsign.as<signed char>(), tnz.as<u32>(), &ctx->d_ctr->max_bits);
    exclusive_sum(tnz.as<u32>(), pos.as<u32>(), n_unique);
    const u64 n_out = last_plus(pos.as<u32>(), tnz.as<u32>(), n_unique);
    DPolyGuard res{ctx, new_dpoly(ctx, n_out, nv, out_limbs)};
    if (n_out)
        PK_LAUNCH(ctx, k_wide_compact, grid_for(n_unique, 256), 256, 0, tkey.as<u64>(), tmag.as<u64>(), tsign.as<signed char>(), tnz.as<u32>(),
                  pos.as<u32>(), n_unique, w.h_min, res.p->key, res.p->planes, res.p->stride, res.p->sign, out_limbs);
    readback(ctx, h2, &ctx->d_ctr->max_bits, 4);
    CUDA_CHECK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    res.p->max_bits = h2[0];
    res.p->limbs = std::max<u32>(1, std::min<u32>(out_limbs, (h2[0] + 63) / 64));
    // statistics of the whole product: the pair count of one half product (all have the same pairs), times of all
    ctx->stats = first;
    ctx->stats.result_terms = n_out;
    ctx->stats.result_limbs = res.p->limbs;
    ctx->stats.kernel_launches = ctx->call_launches;
    ctx->stats.retries = (u32)parts.size() << 16 | (first.retries & 0xFFFFu);
    ctx->ev_valid = false;
    ctx->stats.mul_kernel_ms = kernel_ms;
    return res.release();
}

} // namespace

#include "pk_multi.cuh"

// ------------------------------------------------------------------------------------------------ C ABI

#define PK_TRY(ctx_expr, ...)                                    \
    pk_ctx *ctx__ = (ctx_expr);                                  \
    try {                                                        \
        __VA_ARGS__;                                             \
        return PK_OK;                                            \
    } catch (const PkError &e) {                                 \
        if (ctx__) ctx__->err = e.msg;                           \
        return e.code;                                           \
    } catch (const std::bad_alloc &) {                           \
        if (ctx__) ctx__->err = "host allocation failure";       \
        return PK_E_NOMEM;                                       \
    } catch (const std::exception &e) {                          \
        if (ctx__) ctx__->err = e.what();                        \
        return PK_E_CUDA;                                        \
    }

extern "C" {

int pk_ctx_create(int device, void *stream, pk_ctx **out)
{
    if (!out) return PK_E_ARGS;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return PK_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PK_E_CUDA;
    if (prop.major < 10) return PK_E_CUDA; // kernels are built for sm_100a only
    std::unique_ptr<pk_ctx> c(new (std::nothrow) pk_ctx);
    if (!c) return PK_E_NOMEM;
    c->device = device;
    DeviceScope ds(device); // the caller's current device is restored on return
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    c->total_mem = prop.totalGlobalMem;
    if (stream) {
        c->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return PK_E_CUDA;
        c->own_stream = true;
    }
    // a private stream-ordered pool that keeps freed blocks cached (no cudaMalloc in steady state); the device's
    // default pool and its attributes belong to the host application and are left alone
    cudaMemPoolProps pp{};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    bool ok = cudaMemPoolCreate(&c->pool, &pp) == cudaSuccess;
    if (ok) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ok = ok && cudaEventCreate(&c->ev_begin) == cudaSuccess && cudaEventCreate(&c->ev_mul0) == cudaSuccess
         && cudaEventCreate(&c->ev_mul1) == cudaSuccess && cudaEventCreate(&c->ev_end) == cudaSuccess
         && cudaEventCreate(&c->ev_aux0) == cudaSuccess && cudaEventCreate(&c->ev_aux1) == cudaSuccess
         && cudaHostAlloc(&c->h_scratch, kScratchBytes, cudaHostAllocPortable) == cudaSuccess
         && cudaMalloc((void **)&c->d_ctr, sizeof(MulCounters)) == cudaSuccess
         && cudaMalloc((void **)&c->d_meta, sizeof(MetaDev)) == cudaSuccess
         && cudaMalloc((void **)&c->d_digest, 8) == cudaSuccess;
    if (!ok) {
        pk_ctx_destroy(c.release());
        return PK_E_CUDA;
    }
    const auto &tab = pb200::kronecker_array<int64_t>::get_limits();
    c->limits.resize(tab.size());
    c->h_min.resize(tab.size());
    c->h_max.resize(tab.size());
    for (size_t m = 0; m < tab.size(); ++m) {
        c->limits[m] = std::get<0>(tab[m]);
        c->h_min[m] = std::get<1>(tab[m]);
        c->h_max[m] = std::get<2>(tab[m]);
    }
    zone_configure(c.get());
    block_configure(c.get());
    c->pipe_parts = env_u32("PK_PIPE_PARTS", c->pipe_parts);
    *out = c.release();
    return PK_OK;
}

int pk_ctx_create_multi(uint32_t n_devices, const int *device_ids, pk_ctx **out)
{
    if (!out) return PK_E_ARGS;
    *out = nullptr;
    if (n_devices == 0 || n_devices > kMaxPeers) return PK_E_ARGS;
    std::unique_ptr<pk_ctx> M(new (std::nothrow) pk_ctx);
    if (!M) return PK_E_NOMEM;
    int rc = PK_OK;
    for (uint32_t i = 0; i < n_devices && rc == PK_OK; ++i) {
        pk_ctx *sub = nullptr;
        rc = pk_ctx_create(device_ids ? device_ids[i] : (int)i, nullptr, &sub);
        if (rc == PK_OK) M->subs.push_back(sub);
    }
    if (rc != PK_OK) {
        for (pk_ctx *sub : M->subs) pk_ctx_destroy(sub);
        return rc;
    }
    M->device = M->subs[0]->device;
    M->limits = M->subs[0]->limits;
    M->h_min = M->subs[0]->h_min;
    M->h_max = M->subs[0]->h_max;
    // direct peer copies between the devices (parts of a device-resident product are replicated over NVLink when the
    // product becomes an operand); without peer access cudaMemcpyPeerAsync stages through the host, still correct
    for (pk_ctx *x : M->subs) {
        DeviceScope ds(x->device);
        for (pk_ctx *y : M->subs) {
            int can = 0;
            if (x != y && cudaDeviceCanAccessPeer(&can, x->device, y->device) == cudaSuccess && can)
                if (cudaDeviceEnablePeerAccess(y->device, 0) == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
    }
    try {
        M->multi = new pk_multi_state;
        M->multi->w.resize(n_devices);
        pk_ctx *Mp = M.get();
        for (uint32_t r = 0; r < n_devices; ++r) M->multi->w[r].th = std::thread(multi_worker_main, Mp, r);
    } catch (...) {
        multi_destroy(M.release());
        return PK_E_NOMEM;
    }
    *out = M.release();
    return PK_OK;
}

int pk_device_count(void)
{
    int count = 0, usable = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    for (int d = 0; d < count; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major >= 10) ++usable;
    }
    return usable;
}

int pk_ctx_synchronize(pk_ctx *ctx)
{
    if (!ctx) return PK_E_ARGS;
    PK_TRY(ctx, {
        if (!ctx->subs.empty()) {
            for (pk_ctx *sub : ctx->subs) CUDA_CHECK(cudaStreamSynchronize(sub->stream));
        } else {
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    })
}

uint32_t pk_ctx_device_count(const pk_ctx *ctx) { return ctx ? (ctx->subs.empty() ? 1u : (uint32_t)ctx->subs.size()) : 0u; }

void pk_ctx_destroy(pk_ctx *ctx)
{
    if (!ctx) return;
    if (!ctx->subs.empty() || ctx->multi) {
        multi_destroy(ctx);
        return;
    }
    DeviceScope ds(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->staging) {
        free_dpoly_arrays(ctx, ctx->staging);
        delete ctx->staging;
    }
    for (auto &b : ctx->pinned)
        if (b.ptr) cudaFreeHost(b.ptr);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->d_meta) cudaFree(ctx->d_meta);
    if (ctx->d_digest) cudaFree(ctx->d_digest);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_mul0) cudaEventDestroy(ctx->ev_mul0);
    if (ctx->ev_mul1) cudaEventDestroy(ctx->ev_mul1);
    for (cudaEvent_t e : ctx->chunk_events) cudaEventDestroy(e);
    if (ctx->ev_aux0) cudaEventDestroy(ctx->ev_aux0);
    if (ctx->ev_aux1) cudaEventDestroy(ctx->ev_aux1);
    if (ctx->tiles_buf) cudaFreeAsync(ctx->tiles_buf, ctx->stream);
    if (ctx->bitmap_buf) cudaFreeAsync(ctx->bitmap_buf, ctx->stream);
    ctx_drop_idle(ctx);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_part) cudaEventDestroy(ctx->ev_part);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool); // (blocks still owned by live polynomials are released with them)
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pk_ctx_trim(pk_ctx *ctx)
{
    if (!ctx) return PK_E_ARGS;
    if (!ctx->subs.empty()) {
        for (pk_ctx *sub : ctx->subs) {
            const int rc = pk_ctx_trim(sub);
            if (rc != PK_OK) {
                ctx->err = sub->err;
                return rc;
            }
        }
        for (auto &b : ctx->pinned)
            if (!b.in_use && b.ptr) {
                cudaFreeHost(b.ptr);
                b.ptr = nullptr;
                b.bytes = 0;
            }
        ctx->pinned.erase(std::remove_if(ctx->pinned.begin(), ctx->pinned.end(), [](const PinnedBuf &b) { return !b.in_use && !b.ptr; }),
                          ctx->pinned.end());
        return PK_OK;
    }
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (ctx->staging) {
            free_dpoly_arrays(ctx, ctx->staging);
            delete ctx->staging;
            ctx->staging = nullptr;
            ctx->staging_cap = 0;
        }
        if (ctx->tiles_buf) {
            cudaFreeAsync(ctx->tiles_buf, ctx->stream);
            ctx->tiles_buf = nullptr;
            ctx->tiles_cap = 0;
        }
        if (ctx->bitmap_buf) {
            cudaFreeAsync(ctx->bitmap_buf, ctx->stream);
            ctx->bitmap_buf = nullptr;
            ctx->bitmap_bytes = 0;
        }
        ctx_drop_idle(ctx);
        for (auto &b : ctx->pinned)
            if (!b.in_use && b.ptr) {
                cudaFreeHost(b.ptr);
                b.ptr = nullptr;
                b.bytes = 0;
            }
        ctx->pinned.erase(std::remove_if(ctx->pinned.begin(), ctx->pinned.end(), [](const PinnedBuf &b) { return !b.in_use && !b.ptr; }),
                          ctx->pinned.end());
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (ctx->pool) cudaMemPoolTrimTo(ctx->pool, 0);
    })
}

const char *pk_last_error(const pk_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pk_ctx_set_limits(pk_ctx *ctx, uint32_t n_vars, const int64_t *limits)
{
    if (!ctx || !limits || n_vars == 0 || n_vars >= PK_MAX_VARS) return PK_E_ARGS;
    // recompute h_min / h_max the way determine_limit does (kronecker_array.hpp:155-157) and validate the table
    __int128 c = 1, hmax = 0;
    for (uint32_t v = 0; v < n_vars; ++v) {
        if (limits[v] <= 0) return PK_E_ARGS;
        hmax += c * limits[v];
        c *= 2 * (__int128)limits[v] + 1;
        if (c > ((__int128)1 << 64)) return PK_E_ARGS;
    }
    if (hmax > (__int128)INT64_MAX || 2 * hmax + 1 > (__int128)INT64_MAX) return PK_E_ARGS;
    if (ctx->limits.size() <= n_vars) {
        ctx->limits.resize(n_vars + 1);
        ctx->h_min.resize(n_vars + 1);
        ctx->h_max.resize(n_vars + 1);
    }
    ctx->limits[n_vars].assign(limits, limits + n_vars);
    ctx->h_min[n_vars] = (int64_t)-hmax;
    ctx->h_max[n_vars] = (int64_t)hmax;
    for (pk_ctx *sub : ctx->subs) pk_ctx_set_limits(sub, n_vars, limits);
    return PK_OK;
}

int pk_ctx_get_limits(const pk_ctx *ctx, uint32_t n_vars, int64_t *limits_out, int64_t *h_min_out, int64_t *h_max_out)
{
    if (!ctx || n_vars >= ctx->limits.size() || (n_vars && ctx->limits[n_vars].size() != n_vars)) return PK_E_ARGS;
    if (limits_out)
        for (uint32_t v = 0; v < n_vars; ++v) limits_out[v] = ctx->limits[n_vars][v];
    if (h_min_out) *h_min_out = ctx->h_min[n_vars];
    if (h_max_out) *h_max_out = ctx->h_max[n_vars];
    return PK_OK;
}

int pk_ctx_set_tuning(pk_ctx *ctx, const char *key, uint64_t value)
{
    if (!ctx || !key) return PK_E_ARGS;
    if (!ctx->subs.empty()) {
        for (pk_ctx *sub : ctx->subs) {
            const int rc = pk_ctx_set_tuning(sub, key, value);
            if (rc != PK_OK) return rc;
        }
        return PK_OK;
    }
    ZoneTuning &t = ctx->zone_tuning;
    const std::string k(key);
    if (k == "zone_target") t.target_zones = (u32)value;
    else if (k == "zone_chunk") t.zones_per_chunk = std::max<u32>(1, (u32)value);
    else if (k == "zone_span_log2") t.span_log2_max = std::min<u32>(20, std::max<u32>(10, (u32)value));
    else if (k == "zone_cap") t.force_cap = (u32)value;
    else if (k == "zone_mode") t.force_mode = (u32)value;
    else if (k == "zone_disable") t.disable = (u32)value;
    else if (k == "zone_cursor_bytes") t.cursor_bytes_max = (u32)value;
    else if (k == "zone_uniform") t.uniform = (u32)value;
    else if (k == "zone_tpb") t.tpb = (value == 256 || value == 512 || value == 1024) ? (u32)value : 0;
    else if (k == "block_disable") ctx->block_tuning.disable = (u32)value;
    else if (k == "block_target") ctx->block_tuning.tile_target = std::max<u32>(32, (u32)value);
    else if (k == "block_zone_work") ctx->block_tuning.zone_work = std::max<u32>(1, (u32)value);
    else if (k == "block_min_tile") ctx->block_tuning.min_tile = (u32)value;
    else if (k == "block_mode") ctx->block_tuning.force_mode = (u32)value;
    else if (k == "block_hz") ctx->block_tuning.force_hz = (u32)value;
    else if (k == "block_natural") ctx->block_tuning.natural = (u32)value;
    else if (k == "block_equal_zones") ctx->block_tuning.equal_zones = (u32)value;
    else if (k == "block_zone_products") ctx->block_tuning.zone_products = (u32)value;
    else if (k == "pipe_parts") ctx->pipe_parts = (u32)std::min<uint64_t>(value, 64);
    else if (k == "pipe_min_products") ctx->pipe_min_products = value;
    else if (k == "pipe_test_small") ctx->pipe_test_small = (u32)value;
    else if (k == "block_wide") ctx->block_tuning.wide = (u32)value;
    else if (k == "block_pair_walk") ctx->block_tuning.pair_walk = (u32)value;
    else if (k == "block_hash") ctx->block_tuning.hash = (u32)value;
    else if (k == "block_permute") ctx->block_tuning.permute = (u32)value;
    else if (k == "block_smem") ctx->block_tuning.smem_limit = (u32)value;
    else if (k == "block_radix0") ctx->block_tuning.radix0_residue = std::min<u32>(32, (u32)value);
    else if (k == "block_unit_codes") ctx->block_tuning.unit_codes = std::max<u32>(32, (u32)value);
    else if (k == "block_merge") ctx->block_tuning.merge = (u32)value;
    else if (k == "block_ctas") ctx->block_tuning.ctas = (u32)value;
    else if (k == "block_size_hint") ctx->block_tuning.size_hint = (u32)value;
    else if (k == "block_stage_rows") ctx->block_tuning.stage_rows = (u32)value;
    else if (k == "block_flatcol") ctx->block_tuning.flatcol = (u32)value;
    else if (k == "block_self_clear") ctx->block_tuning.self_clear = (u32)value;
    else return PK_E_ARGS;
    return PK_OK;
}

int pk_get_stats(pk_ctx *ctx, pk_stats *out)
{
    if (!ctx || !out) return PK_E_ARGS;
    if (!ctx->subs.empty()) { // aggregated by the last product (sums over the parts, times = the slowest device)
        *out = ctx->stats;
        return PK_OK;
    }
    if (ctx->ev_valid) {
        float ms = 0.f;
        cudaEventSynchronize(ctx->ev_end); // the last copy of a product may still be running
        if (cudaEventElapsedTime(&ms, ctx->ev_mul0, ctx->ev_mul1) == cudaSuccess) ctx->stats.mul_kernel_ms = ms;
        if (ctx->aux_valid && cudaEventElapsedTime(&ms, ctx->ev_aux0, ctx->ev_aux1) == cudaSuccess) ctx->stats.mul_kernel_ms += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) ctx->stats.device_ms = ms;
    }
    ctx->stats.total_launches = ctx->total_launches;
    *out = ctx->stats;
    return PK_OK;
}

int pk_upload(pk_ctx *ctx, const pk_poly *p, pk_dpoly **out)
{
    if (!ctx || !out) return PK_E_ARGS;
    *out = nullptr;
    PK_TRY(ctx, {
        if (!ctx->subs.empty()) {
            *out = multi_upload(ctx, p);
        } else {
            DeviceScope ds(ctx->device);
            *out = upload_impl(ctx, p);
        }
    })
}

int pk_mul_dev(pk_ctx *ctx, const pk_dpoly *a, const pk_dpoly *b, const pk_opts *opts, pk_dpoly **out)
{
    if (!ctx || !out) return PK_E_ARGS;
    *out = nullptr;
    PK_TRY(ctx, {
        if (!ctx->subs.empty()) {
            *out = multi_mul_dev(ctx, const_cast<pk_dpoly *>(a), const_cast<pk_dpoly *>(b), opts);
        } else {
            if ((a && a->is_multi) || (b && b->is_multi)) fail(PK_E_ARGS, "operand belongs to a multi-device context");
            DeviceScope ds(ctx->device);
            *out = mul_impl(ctx, const_cast<pk_dpoly *>(a), const_cast<pk_dpoly *>(b), opts);
        }
    })
}

int pk_check_bounds(pk_ctx *ctx, const pk_dpoly *a, const pk_dpoly *b)
{
    if (!ctx || !a || !b) return PK_E_ARGS;
    if (!ctx->subs.empty()) { // the exponent bounds of the replicas on the first device
        PK_TRY(ctx, {
            multi_require(ctx, a);
            multi_require(ctx, b);
            multi_replicate(ctx, const_cast<pk_dpoly *>(a));
            multi_replicate(ctx, const_cast<pk_dpoly *>(b));
            const int rc = pk_check_bounds(ctx->subs[0], a->rep[0], b->rep[0]);
            if (rc != PK_OK) fail(rc, ctx->subs[0]->err);
        })
    }
    PK_TRY(ctx, {
        if (a->is_multi || b->is_multi) fail(PK_E_ARGS, "operand belongs to a multi-device context");
        if (a->n_vars != b->n_vars) fail(PK_E_ARGS, "incompatible arguments sets");
        if (a->n_vars >= ctx->limits.size()) fail(PK_E_ARGS, "too many variables for the Kronecker codification");
        if (a->n && b->n) {
            DeviceScope ds(ctx->device);
            ensure_meta(ctx, const_cast<pk_dpoly *>(a));
            ensure_meta(ctx, const_cast<pk_dpoly *>(b));
            for (u32 v = 0; v < a->n_vars; ++v) {
                const __int128 lo = (__int128)a->emin[v] + b->emin[v], hi = (__int128)a->emax[v] + b->emax[v];
                const i64 M = ctx->limits[a->n_vars][v];
                if (lo < -(__int128)M || hi > (__int128)M) fail(PK_E_KEY_RANGE, "Kronecker monomial components are out of bounds");
            }
        }
    })
}

int pk_download(pk_ctx *ctx, const pk_dpoly *p, pk_result *out)
{
    if (!ctx || !p || !out) return PK_E_ARGS;
    PK_TRY(ctx, {
        if (!ctx->subs.empty()) {
            multi_download(ctx, p, out);
        } else {
            if (p->is_multi) fail(PK_E_ARGS, "polynomial belongs to a multi-device context");
            DeviceScope ds(ctx->device);
            download_impl(ctx, p, out);
        }
    })
}

int pk_download_grouped_progress(pk_ctx *ctx, const pk_dpoly *p, uint32_t shift, uint32_t bits, uint32_t n_chunks, pk_progress_fn fn,
                                 void *user, uint64_t *group_offsets, pk_result *out)
{
    if (!ctx || !p || !out || bits < 1 || bits > 16 || shift > 63 || n_chunks < 1 || n_chunks > 256) return PK_E_ARGS;
    PK_TRY(ctx, {
        if (!ctx->subs.empty() || p->is_multi) fail(PK_E_UNSUPPORTED, "pk_download_grouped: not on a multi-device context (use pk_download)");
        DeviceScope ds(ctx->device);
        download_grouped_impl(ctx, p, shift, bits, out, n_chunks, fn, user, group_offsets);
    })
}

int pk_download_grouped(pk_ctx *ctx, const pk_dpoly *p, uint32_t shift, uint32_t bits, pk_result *out)
{
    return pk_download_grouped_progress(ctx, p, shift, bits, 1, nullptr, nullptr, nullptr, out);
}

void pk_dpoly_free(pk_ctx *ctx, pk_dpoly *p)
{
    if (!ctx || !p) return;
    if (p->is_multi) {
        if (!ctx->subs.empty()) multi_free(ctx, p);
        return;
    }
    DeviceScope ds(ctx->device);
    free_dpoly_arrays(ctx, p);
    delete p;
}

uint64_t pk_dpoly_size(const pk_dpoly *p) { return p ? p->n : 0; }
uint32_t pk_dpoly_limbs(const pk_dpoly *p) { return p ? p->limbs : 0; }
uint32_t pk_dpoly_nvars(const pk_dpoly *p) { return p ? p->n_vars : 0; }

int pk_dpoly_export_padded(pk_ctx *ctx, const pk_dpoly *p, void *d_key, void *d_mag, void *d_sign, uint32_t mag_limbs)
{
    if (!ctx || !p || mag_limbs < p->limbs || p->is_multi || !ctx->subs.empty()) return PK_E_ARGS;
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        if (p->n) {
            if (d_key) CUDA_CHECK(cudaMemcpyAsync(d_key, p->key, p->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            if (d_mag)
                PK_LAUNCH(ctx, k_planes_to_aos, grid_for(p->n, 256), 256, 0, p->planes, static_cast<u64 *>(d_mag), p->n,
                          p->limbs, p->stride, mag_limbs);
            if (d_sign) CUDA_CHECK(cudaMemcpyAsync(d_sign, p->sign, p->n, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    })
}

int pk_dpoly_export(pk_ctx *ctx, const pk_dpoly *p, void *d_key, void *d_mag, void *d_sign)
{
    return pk_dpoly_export_padded(ctx, p, d_key, d_mag, d_sign, p ? p->limbs : 0);
}

namespace
{
bool fill_peer_table(PeerTable &t, uint32_t n_peers, const uint64_t *slots, const uint64_t *keys, const uint64_t *mags,
                     const uint64_t *signs)
{
    if (!n_peers || n_peers > kMaxPeers) return false;
    for (uint32_t j = 0; j < n_peers; ++j) {
        t.slot[j] = slots ? reinterpret_cast<u64 *>(slots[j]) : nullptr;
        t.key[j] = keys ? reinterpret_cast<i64 *>(keys[j]) : nullptr;
        t.mag[j] = mags ? reinterpret_cast<u64 *>(mags[j]) : nullptr;
        t.sign[j] = signs ? reinterpret_cast<signed char *>(signs[j]) : nullptr;
    }
    return true;
}
} // namespace

int pk_dpoly_publish_size(pk_ctx *ctx, const pk_dpoly *p, uint32_t n_peers, const uint64_t *slot_ptrs)
{
    PeerTable t{};
    if (!ctx || !p || p->is_multi || !ctx->subs.empty() || !slot_ptrs || !fill_peer_table(t, n_peers, slot_ptrs, nullptr, nullptr, nullptr))
        return PK_E_ARGS;
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        PK_LAUNCH(ctx, k_peer_publish, 1, 32, 0, t, n_peers, p->n, (u64)std::max<u32>(1, p->limbs));
    })
}

int pk_dpoly_export_peers(pk_ctx *ctx, const pk_dpoly *p, uint32_t n_peers, uint32_t rank, const void *d_sizes,
                          const uint64_t *key_ptrs, const uint64_t *mag_ptrs, const uint64_t *sign_ptrs, uint64_t cap_terms,
                          uint32_t cap_limbs)
{
    PeerTable t{};
    if (!ctx || !p || !d_sizes || !key_ptrs || !mag_ptrs || !sign_ptrs || rank >= n_peers || p->limbs > 8 || cap_limbs > 8
        || p->is_multi || !ctx->subs.empty() || !fill_peer_table(t, n_peers, nullptr, key_ptrs, mag_ptrs, sign_ptrs))
        return PK_E_ARGS;
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        if (p->n)
            PK_LAUNCH(ctx, k_peer_push, grid_for(p->n, 256), 256, 0, t, n_peers, rank, static_cast<const u64 *>(d_sizes), p->key,
                      p->planes, p->stride, p->sign, p->n, p->limbs, cap_terms, cap_limbs);
    })
}

int pk_dpoly_digest(pk_ctx *ctx, const pk_dpoly *p, uint64_t *digest_out)
{
    if (!ctx || !p || !digest_out) return PK_E_ARGS;
    if (!ctx->subs.empty()) {
        PK_TRY(ctx, { *digest_out = multi_digest(ctx, p); })
    }
    if (p->is_multi) return PK_E_ARGS;
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        CUDA_CHECK(cudaMemsetAsync(ctx->d_digest, 0, 8, ctx->stream));
        if (p->n) {
            const u32 grid = std::min<u32>(grid_for(p->n, 256), (u32)ctx->sm_count * 8);
            PK_LAUNCH(ctx, k_digest, grid, 256, 0, p->key, p->planes, p->sign, p->n, p->limbs,
                      p->stride, ctx->d_digest);
        }
        readback(ctx, ctx->h_scratch, ctx->d_digest, 8);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *digest_out = *static_cast<uint64_t *>(ctx->h_scratch);
    })
}

namespace
{

// A product whose result is large next to its work (sparse output range) is bound by the device-to-host copy of the
// result.  Such a product is computed in `pipe_parts` consecutive output ranges (the same partition mechanism the
// ranks of a multi-GPU product use), and the terms of part p travel to the host on a second stream while part p+1 is
// being computed.  Parts are ordered ranges, so they land back to back in the host arrays and the result is sorted.
bool should_pipeline(pk_ctx *ctx, const pk_dpoly *a, const pk_dpoly *b, const pk_opts *opts)
{
    if (ctx->pipe_parts < 2 || !a->n || !b->n || a->n_vars != b->n_vars) return false;
    if (opts && (opts->part_count > 1 || opts->trunc_mode != 0)) return false;
    const unsigned __int128 W = (unsigned __int128)a->n * b->n;
    if (W < ctx->pipe_min_products) return false;
    unsigned __int128 range = 1; // tight code range of the product (as in make_plan)
    for (u32 v = 0; v < a->n_vars; ++v) {
        u64 g = gcd64(a->egcd[v], b->egcd[v]);
        if (!g) g = 1;
        range *= ((u64)(a->emax[v] - a->emin[v]) + (u64)(b->emax[v] - b->emin[v])) / g + 1;
        if (range >= ((unsigned __int128)1 << 62)) break;
    }
    return range > 3 * W; // same test that selects bitmap zones: the result has about as many terms as products allow
}

void collect_part_stats(pk_ctx *ctx, pk_stats &acc)
{
    float ms = 0.f;
    if (ctx->ev_valid) {
        cudaEventSynchronize(ctx->ev_end);
        if (cudaEventElapsedTime(&ms, ctx->ev_mul0, ctx->ev_mul1) == cudaSuccess) acc.mul_kernel_ms += ms;
        if (ctx->aux_valid && cudaEventElapsedTime(&ms, ctx->ev_aux0, ctx->ev_aux1) == cudaSuccess) acc.mul_kernel_ms += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) acc.device_ms += ms;
    }
    const pk_stats &s = ctx->stats;
    acc.n1 = s.n1;
    acc.n2 = s.n2;
    acc.term_products += s.term_products;
    acc.result_terms += s.result_terms;
    acc.code_range = s.code_range;
    acc.table_slots = s.table_slots;
    acc.algo = s.algo;
    acc.acc_lanes = s.acc_lanes;
    acc.chunk_bits = s.chunk_bits;
    acc.result_limbs = std::max(acc.result_limbs, s.result_limbs);
    acc.kernel_launches += s.kernel_launches;
    acc.retries = s.retries;
}

void mul_pipelined(pk_ctx *ctx, pk_dpoly *da, pk_dpoly *db, const pk_opts *opts, pk_result *out)
{
    const u32 P = ctx->pipe_parts;
    if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->ev_part) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_part, cudaEventDisableTiming));
    struct Parts {
        pk_ctx *ctx;
        std::vector<pk_dpoly *> v;
        std::vector<void *> aos; // magnitude rows staged for the copies; released only after the copies are done (a
                                 // stream-ordered free on the copy stream would make the next part's allocations wait
                                 // for the copy)
        ~Parts()
        {
            cudaStreamSynchronize(ctx->copy_stream); // copies may still read the parts
            for (void *q : aos) cudaFreeAsync(q, ctx->stream);
            for (pk_dpoly *p : v)
                if (p) {
                    free_dpoly_arrays(ctx, p);
                    delete p;
                }
        }
    } parts{ctx, {}};
    pk_stats acc{};
    PinnedBuf *pb = nullptr;
    u64 cap = 0, off = 0;
    u32 L = 1;
    bool overflow = false;
    char *base = nullptr;
    auto copy_part = [&](const pk_dpoly *p, u64 at, cudaStream_t after) {
        if (!p->n) return;
        void *aos = nullptr;
        CUDA_CHECK(cudaMallocFromPoolAsync(&aos, p->n * L * 8, ctx->pool, ctx->stream));
        parts.aos.push_back(aos);
        PK_LAUNCH(ctx, k_planes_to_aos, grid_for(p->n, 256), 256, 0, p->planes, static_cast<u64 *>(aos), p->n, p->limbs, p->stride, L);
        CUDA_CHECK(cudaEventRecord(ctx->ev_part, ctx->stream));
        CUDA_CHECK(cudaStreamWaitEvent(after, ctx->ev_part, 0));
        CUDA_CHECK(cudaMemcpyAsync(base + at * 8, p->key, p->n * 8, cudaMemcpyDeviceToHost, after));
        CUDA_CHECK(cudaMemcpyAsync(base + cap * 8 + at * L * 8, aos, p->n * L * 8, cudaMemcpyDeviceToHost, after));
        CUDA_CHECK(cudaMemcpyAsync(base + cap * 8 + cap * L * 8 + at, p->sign, p->n, cudaMemcpyDeviceToHost, after));
    };
    try {
        for (u32 p = 0; p < P; ++p) {
            pk_opts o{};
            if (opts) o = *opts;
            o.struct_size = sizeof(pk_opts);
            o.part_index = p;
            o.part_count = P;
            parts.v.push_back(mul_impl(ctx, da, db, &o));
            pk_dpoly *part = parts.v.back();
            collect_part_stats(ctx, acc);
            if (p == 0) { // size the host arrays from the first part (parts are balanced by work, not by terms)
                L = part->alloc_limbs;
                cap = std::max<u64>(4096, (part->n * P) + (part->n * P) / 4 + 4096);
                if (ctx->pipe_test_small) cap = std::max<u64>(1, part->n);
                cap = (cap + 63) & ~63ull;
                pb = pinned_acquire(ctx, cap * (9 + 8 * (size_t)L));
                base = static_cast<char *>(pb->ptr);
            }
            if (off + part->n > cap) overflow = true;
            if (!overflow) copy_part(part, off, ctx->copy_stream);
            off += part->n;
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
        if (overflow) { // the estimate was too small: exact allocation, copy everything again
            pb->in_use = false;
            cap = (off + 63) & ~63ull;
            pb = pinned_acquire(ctx, cap * (9 + 8 * (size_t)L));
            base = static_cast<char *>(pb->ptr);
            u64 at = 0;
            for (pk_dpoly *part : parts.v) {
                copy_part(part, at, ctx->stream);
                at += part->n;
            }
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    } catch (...) {
        if (pb) pb->in_use = false;
        throw;
    }
    out->n = off;
    out->n_vars = da->n_vars;
    out->limbs = L;
    if (off) {
        out->key = reinterpret_cast<int64_t *>(base);
        out->mag = reinterpret_cast<uint64_t *>(base + cap * 8);
        out->sign = reinterpret_cast<int8_t *>(base + cap * 8 + cap * L * 8);
        out->owner = pb->ptr;
    } else if (pb) {
        pb->in_use = false;
    }
    ctx->stats = acc;
    ctx->ev_valid = false; // the per-part event times were summed above
}

} // namespace

int pk_mul(pk_ctx *ctx, const pk_poly *a, const pk_poly *b, const pk_opts *opts, pk_result *out)
{
    if (!ctx || !out) return PK_E_ARGS;
    memset(out, 0, sizeof(*out));
    if (!ctx->subs.empty()) {
        PK_TRY(ctx, { multi_mul(ctx, a, b, opts, out); })
    }
    PK_TRY(ctx, {
        DeviceScope ds(ctx->device);
        DPolyGuard da{ctx, upload_impl(ctx, a)};
        DPolyGuard db{ctx, (a == b) ? nullptr : upload_impl(ctx, b)};
        pk_dpoly *pa = da.p;
        pk_dpoly *pb2 = db.p ? db.p : da.p;
        if (should_pipeline(ctx, pa, pb2, opts)) {
            mul_pipelined(ctx, pa, pb2, opts, out);
        } else {
            DPolyGuard dr{ctx, mul_impl(ctx, pa, pb2, opts)};
            download_impl(ctx, dr.p, out);
        }
    })
}

void pk_result_free(pk_ctx *ctx, pk_result *r)
{
    if (!ctx || !r) return;
    for (auto &b : ctx->pinned)
        if (b.ptr && b.ptr == r->owner) b.in_use = false;
    memset(r, 0, sizeof(*r));
}

} // extern "C"
