// Shared-memory zone path (PK_ALGO_ZONE): the GPU form of piranha's zone scheme (polynomial.hpp:1783-1966).
//
// The output range of tight codes [lo, hi) is cut into zones of `span` consecutive codes.  A zone is owned by ONE
// thread block for its whole life, so every accumulator of the zone lives in that SM's shared memory and all
// multiply-accumulates are shared-memory atomics (measured on B200: ~10 ATOMS/clk/SM with the carry chain, 13x the
// rate of L2-resident global RED.64 and 60x the rate out of HBM -- tools/microbench.cu, profiles/).
//
//   * Both operands are sorted by code, so for a row a_i the terms b_j landing in [zlo, zhi) are one contiguous
//     span of B: the role of l_bound() in polynomial.hpp:1800-1826.  Consecutive zones are processed by the same
//     block ("chunk"), and each row keeps a cursor in shared memory, so the span of the next zone starts where the
//     previous one stopped and only one binary search per row per chunk is needed.
//   * DENSE zones: accumulator slot = code - zlo (compact ranges: the Fateman family, monagan1-5).
//     BITMAP zones: a first walk marks the codes that occur in a shared-memory bitmap; a popcount prefix turns the
//     bitmap into a perfect, order-preserving hash (rank), and the second walk accumulates at acc[rank(code)]
//     (sparse ranges: pearce1).  Either way the entries of a zone come out in ascending code order, so no sort.
//   * Accumulators are NW 32-bit words, two's complement, stored as word planes (bank-conflict-free for code
//     runs).  A product is added word by word with ATOMS.ADD; the returned old value gives the carry (borrow) into
//     the next word.  Each atomic is exact modulo 2^32 and reports its own carry, so the value is exact under any
//     interleaving.  NW is chosen on the host from a proven bound, so the top word never wraps.
//   * A finished zone claims space in a staging area with one atomicAdd (no inter-block waiting of any kind) and
//     records (offset, count); an exclusive scan over the zone counts and one gather pass then lay the zones out in
//     zone order = ascending code order.  Zero coefficients are dropped while emitting: sanitise_series
//     (base_series_multiplier.hpp:855-949) fused in.
#pragma once

#include <cstdlib>

#include "pk_host.cuh"

namespace pk
{

constexpr int kZoneThreads = 1024;

struct ZoneArgs {
    const ulonglong2 *A; // {tight code | sign bit, magnitude limb 0}, ascending code
    u64 nA;
    const ulonglong2 *B;
    u64 nB;
    u64 lo, hi; // output partition of this call
    u64 span;   // codes per zone
    u32 n_zones, zones_per_chunk, n_chunks;
    u32 mode; // 0 dense, 1 bitmap
    u32 cap;  // accumulator entries per zone pass
    u32 use_cursors;
    u32 pw; // 32-bit words of a product magnitude
    u32 bm_words;
    u32 off_cursor, off_acc, off_bm, off_pre; // byte offsets into dynamic shared memory
    i64 *okey;
    u64 *oplanes;
    u64 ostride;
    signed char *osign;
    u32 out_limbs;
    u64 out_cap;
    CodecParams cp;
    i64 base_code;
    u32 *chunk_counter;
    u32 *zone_count; // result terms of zone z
    u64 *zone_off;   // where zone z starts in the staging arrays
    u64 *bump;       // staging allocation cursor
    MulCounters *ctr;
};

__global__ void k_pack_operand(const u64 *__restrict__ key, const u64 *__restrict__ m0, ulonglong2 *__restrict__ out, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_ulonglong2(key[i], m0[i]);
}

// first index with (T[idx].x & mask) >= x
__device__ __forceinline__ u32 zone_lower_bound(const ulonglong2 *__restrict__ T, u32 n, u64 x)
{
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if ((__ldg(&T[mid].x) & kKeyMask) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// acc[w * cap + slot] += / -= the 128-bit magnitude (p0..p3), eager carry through the returned old values.
template <int NW>
__device__ __forceinline__ void zone_accumulate(u32 *acc, u32 cap, u32 slot, u64 plo, u64 phi, u32 pw, bool neg)
{
    const u32 p[4] = {(u32)plo, (u32)(plo >> 32), (u32)phi, (u32)(phi >> 32)};
    u32 c = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if ((u32)w >= pw && c == 0) break;
        const u64 t = (u64)((w < 4) ? p[w < 4 ? w : 0] : 0u) + c;
        if (t == 0) {
            c = 0;
        } else if (t >> 32) {
            c = 1; // exactly 2^32: this word is unchanged, the carry (borrow) moves on
        } else {
            const u32 x = (u32)t;
            if (!neg) {
                const u32 old = atomicAdd(&acc[w * cap + slot], x);
                c = (old + x) < x;
            } else {
                const u32 old = atomicAdd(&acc[w * cap + slot], 0u - x);
                c = old < x;
            }
        }
    }
}

__device__ __forceinline__ i64 zone_code_to_piranha(u64 code, const CodecParams &cp, i64 base_code)
{
    i64 out = base_code;
    if (code <= 0xFFFFFFFFull) { // 32-bit divisions are ~5x cheaper than 64-bit ones
        u32 c = (u32)code;
        for (u32 v = 0; v < cp.n_vars; ++v) {
            const u32 r = (u32)cp.tradix[v];
            const u32 q = c / r, d = c - q * r;
            c = q;
            out += (i64)d * cp.step[v] * cp.pstride[v];
        }
        return out;
    }
    return tight_to_piranha(code, cp, base_code);
}

// Two's complement NW words -> sign, magnitude; writes the term at `pos`.
template <int NW>
__device__ __forceinline__ void zone_emit(const ZoneArgs &a, u64 pos, u64 code, const u32 (&wd)[NW], u32 &maxbits, u32 &flags)
{
    u32 m[NW];
    const bool neg = (wd[NW - 1] >> 31) != 0;
    if (neg) {
        u32 carry = 1;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const u32 t = ~wd[w] + carry;
            carry = (carry && t == 0) ? 1u : 0u;
            m[w] = t;
        }
    } else {
#pragma unroll
        for (int w = 0; w < NW; ++w) m[w] = wd[w];
    }
    if (pos >= a.out_cap) {
        flags |= kFlagTableFull;
        return;
    }
    u32 bits = 0;
#pragma unroll
    for (int l = 0; l < (NW + 1) / 2; ++l) {
        const u64 limb = (u64)m[2 * l] | ((2 * l + 1 < NW) ? ((u64)m[(2 * l + 1 < NW) ? 2 * l + 1 : 0] << 32) : 0ull);
        if ((u32)l < a.out_limbs) a.oplanes[(u64)l * a.ostride + pos] = limb;
        else if (limb) flags |= kFlagWidth;
        if (limb) bits = 64u * l + bitlen64(limb);
    }
    for (u32 l = (NW + 1) / 2; l < a.out_limbs; ++l) a.oplanes[(u64)l * a.ostride + pos] = 0;
    a.okey[pos] = zone_code_to_piranha(code, a.cp, a.base_code);
    a.osign[pos] = neg ? (signed char)-1 : (signed char)1;
    maxbits = max(maxbits, bits);
}

template <int NW>
__global__ void __launch_bounds__(kZoneThreads, 1) k_zone(const ZoneArgs a)
{
    extern __shared__ __align__(16) unsigned char zsmem[];
    u32 *cursor = reinterpret_cast<u32 *>(zsmem + a.off_cursor);
    u32 *acc = reinterpret_cast<u32 *>(zsmem + a.off_acc);
    u32 *bm = reinterpret_cast<u32 *>(zsmem + a.off_bm);
    u32 *pre = reinterpret_cast<u32 *>(zsmem + a.off_pre);
    __shared__ u32 s_ws[40];
    __shared__ u64 s_b64[2];
    __shared__ u32 s_b32[4];

    const u32 tid = threadIdx.x;
    const u32 nA = (u32)a.nA, nB = (u32)a.nB;
    const u64 kb0 = __ldg(&a.B[0].x) & kKeyMask, kbmax = __ldg(&a.B[nB - 1].x) & kKeyMask;
    u64 my_pairs = 0;
    u32 my_maxbits = 0, my_flags = 0;

    // accumulators (and bitmap) start cleared; every finalisation leaves them cleared again
    for (u32 i = tid; i < a.cap * NW; i += kZoneThreads) acc[i] = 0;
    if (a.mode == 1)
        for (u32 i = tid; i < a.bm_words; i += kZoneThreads) bm[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_b32[0] = atomicAdd(a.chunk_counter, 1u);
        __syncthreads();
        const u32 chunk = s_b32[0];
        __syncthreads();
        if (chunk >= a.n_chunks) break;
        const u32 z0 = chunk * a.zones_per_chunk, z1 = min(a.n_zones, z0 + a.zones_per_chunk);
        if (a.use_cursors) { // one binary search per row per chunk
            const u64 clo = a.lo + (u64)z0 * a.span;
            for (u32 i = tid; i < nA; i += kZoneThreads) {
                const u64 ka = __ldg(&a.A[i].x) & kKeyMask;
                cursor[i] = (ka >= clo) ? 0u : zone_lower_bound(a.B, nB, clo - ka);
            }
            __syncthreads();
        }
        for (u32 z = z0; z < z1; ++z) {
            const u64 zlo = a.lo + (u64)z * a.span;
            const u64 zhi = min(a.hi, zlo + a.span);
            // rows that can reach the zone: a_i + b_max >= zlo and a_i + b_0 < zhi
            const u32 i_beg = (zlo > kbmax) ? zone_lower_bound(a.A, nA, zlo - kbmax) : 0u;
            const u32 i_end = (zhi > kb0) ? zone_lower_bound(a.A, nA, zhi - kb0) : 0u;

            // ---------------- walk 1 (bitmap zones): mark the codes that occur
            u32 zone_count = 0; // distinct codes of the zone (bitmap mode)
            if (a.mode == 1) {
                for (u32 i = i_beg + tid; i < i_end; i += kZoneThreads) {
                    const u64 ka = __ldg(&a.A[i].x) & kKeyMask;
                    u32 j = a.use_cursors ? cursor[i] : ((ka >= zlo) ? 0u : zone_lower_bound(a.B, nB, zlo - ka));
                    while (j < nB) {
                        const u64 code = ka + (__ldg(&a.B[j].x) & kKeyMask);
                        if (code >= zhi) break;
                        const u32 slot = (u32)(code - zlo);
                        atomicOr(&bm[slot >> 5], 1u << (slot & 31u));
                        ++j;
                    }
                }
                __syncthreads();
                // exclusive popcount prefix per bitmap word: rank(code) = pre[word] + popc(bits below)
                const u32 wpt = (a.bm_words + kZoneThreads - 1) / kZoneThreads;
                const u32 w0 = min(a.bm_words, tid * wpt), w1 = min(a.bm_words, w0 + wpt);
                u32 local = 0;
                for (u32 w = w0; w < w1; ++w) local += __popc(bm[w]);
                u32 total;
                u32 run = block_excl_scan(local, s_ws, total);
                for (u32 w = w0; w < w1; ++w) {
                    pre[w] = run;
                    run += __popc(bm[w]);
                }
                zone_count = total;
                __syncthreads();
            }

            // ---------------- accumulate + emit, in passes of at most `cap` entries
            const u32 n_entries = (a.mode == 1) ? zone_count : (u32)(zhi - zlo);
            const bool single = n_entries <= a.cap;
            u64 zone_base = 0, zone_written = 0;
            bool have_base = false;
            u32 r0 = 0;      // first rank (bitmap) / slot (dense) of this pass
            u64 c0 = zlo;    // first code of this pass
            do {
                const u32 r1 = min(n_entries, r0 + a.cap);
                u64 c1 = zhi;
                if (r1 < n_entries) { // code of rank r1 bounds this pass (bitmap zones with more entries than cap)
                    if (tid == 0) {
                        u32 lo_w = 0, hi_w = a.bm_words; // last word with pre[w] <= r1
                        while (hi_w - lo_w > 1) {
                            const u32 mid = (lo_w + hi_w) >> 1;
                            if (pre[mid] <= r1) lo_w = mid;
                            else hi_w = mid;
                        }
                        while (__popc(bm[lo_w]) <= (int)(r1 - pre[lo_w])) ++lo_w; // skip empty words
                        const u32 bit = __fns(bm[lo_w], 0, (int)(r1 - pre[lo_w]) + 1);
                        s_b64[0] = zlo + (u64)lo_w * 32u + bit;
                    }
                    __syncthreads();
                    c1 = s_b64[0];
                    __syncthreads();
                }
                for (u32 i = i_beg + tid; i < i_end; i += kZoneThreads) {
                    const ulonglong2 av = __ldg(&a.A[i]);
                    const u64 ka = av.x & kKeyMask;
                    u32 j = a.use_cursors ? cursor[i] : ((ka >= c0) ? 0u : zone_lower_bound(a.B, nB, c0 - ka));
                    while (j < nB) {
                        const ulonglong2 bv = __ldg(&a.B[j]);
                        const u64 code = ka + (bv.x & kKeyMask);
                        if (code >= c1) break;
                        const u32 slot = (u32)(code - zlo);
                        u32 e;
                        if (a.mode == 1) e = pre[slot >> 5] + __popc(bm[slot >> 5] & ((1u << (slot & 31u)) - 1u)) - r0;
                        else e = slot - r0;
                        const u64 plo = av.y * bv.y, phi = __umul64hi(av.y, bv.y);
                        zone_accumulate<NW>(acc, a.cap, e, plo, phi, a.pw, ((av.x ^ bv.x) >> 63) != 0);
                        ++j;
                        ++my_pairs;
                    }
                    if (a.use_cursors) cursor[i] = j;
                }
                __syncthreads();

                // ---- emit the entries [r0, r1) in ascending code order; threads own contiguous pieces
                const u32 n_pass = r1 - r0;
                u32 cnt = 0;
                u32 u0, u1; // unit range of this thread: slots (dense) or bitmap words (bitmap)
                if (a.mode == 1) {
                    const u32 wpt = (a.bm_words + kZoneThreads - 1) / kZoneThreads;
                    u0 = min(a.bm_words, tid * wpt);
                    u1 = min(a.bm_words, u0 + wpt);
                    for (u32 w = u0; w < u1; ++w) {
                        u32 bits = bm[w], rank = pre[w];
                        while (bits) {
                            bits &= bits - 1;
                            if (rank >= r0 && rank < r1) {
                                u32 any = 0;
#pragma unroll
                                for (int q = 0; q < NW; ++q) any |= acc[q * a.cap + (rank - r0)];
                                cnt += any != 0;
                            }
                            ++rank;
                        }
                    }
                } else {
                    const u32 spt = (n_pass + kZoneThreads - 1) / kZoneThreads;
                    u0 = min(n_pass, tid * spt);
                    u1 = min(n_pass, u0 + spt);
                    for (u32 s = u0; s < u1; ++s) {
                        u32 any = 0;
#pragma unroll
                        for (int q = 0; q < NW; ++q) any |= acc[q * a.cap + s];
                        cnt += any != 0;
                    }
                }
                u32 total;
                const u32 excl = block_excl_scan(cnt, s_ws, total);
                if (!have_base) { // claim staging space: exact for a single pass, the distinct count otherwise
                    if (tid == 0) s_b64[1] = atomicAdd(a.bump, single ? (u64)total : (u64)n_entries);
                    __syncthreads();
                    zone_base = s_b64[1];
                    have_base = true;
                    __syncthreads();
                }
                u64 pos = zone_base + zone_written + excl;
                if (a.mode == 1) {
                    for (u32 w = u0; w < u1; ++w) {
                        u32 bits = bm[w], rank = pre[w];
                        while (bits) {
                            const u32 b = __ffs(bits) - 1;
                            bits &= bits - 1;
                            if (rank >= r0 && rank < r1) {
                                u32 wd[NW], any = 0;
#pragma unroll
                                for (int q = 0; q < NW; ++q) {
                                    wd[q] = acc[q * a.cap + (rank - r0)];
                                    acc[q * a.cap + (rank - r0)] = 0;
                                    any |= wd[q];
                                }
                                if (any) {
                                    zone_emit<NW>(a, pos, zlo + (u64)w * 32u + b, wd, my_maxbits, my_flags);
                                    ++pos;
                                }
                            }
                            ++rank;
                        }
                    }
                } else {
                    for (u32 s = u0; s < u1; ++s) {
                        u32 wd[NW], any = 0;
#pragma unroll
                        for (int q = 0; q < NW; ++q) {
                            wd[q] = acc[q * a.cap + s];
                            acc[q * a.cap + s] = 0;
                            any |= wd[q];
                        }
                        if (any) {
                            zone_emit<NW>(a, pos, zlo + r0 + s, wd, my_maxbits, my_flags);
                            ++pos;
                        }
                    }
                }
                zone_written += total;
                r0 = r1;
                c0 = c1;
                __syncthreads();
            } while (r0 < n_entries);
            if (tid == 0) {
                a.zone_count[z] = (u32)zone_written;
                a.zone_off[z] = zone_base;
            }
            if (a.mode == 1) { // clear the bitmap for the next zone
                for (u32 i = tid; i < a.bm_words; i += kZoneThreads) bm[i] = 0;
            }
            __syncthreads();
        }
    }
    // per-thread statistics
    for (int o = 16; o > 0; o >>= 1) {
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
        my_maxbits = max(my_maxbits, __shfl_xor_sync(0xffffffffu, my_maxbits, o));
        my_flags |= __shfl_xor_sync(0xffffffffu, my_flags, o);
    }
    if (lane_id() == 0) {
        if (my_pairs) atomicAdd(&a.ctr->pairs, my_pairs);
        if (my_maxbits) atomicMax(&a.ctr->max_bits, my_maxbits);
        if (my_flags) atomicOr(&a.ctr->flags, my_flags);
    }
}

// Lay the zones out in zone order (= ascending code order): one block per zone copies its staged terms to
// prefix[z] in the final arrays.
__global__ void __launch_bounds__(256) k_zone_gather(const u32 *__restrict__ zone_count, const u64 *__restrict__ zone_off,
                                                     const u64 *__restrict__ zone_prefix, u32 n_zones,
                                                     const i64 *__restrict__ skey, const u64 *__restrict__ splanes,
                                                     u64 sstride, const signed char *__restrict__ ssign,
                                                     i64 *__restrict__ okey, u64 *__restrict__ oplanes, u64 ostride,
                                                     signed char *__restrict__ osign, u32 limbs)
{
    for (u32 z = blockIdx.x; z < n_zones; z += gridDim.x) {
        const u32 n = zone_count[z];
        const u64 src = zone_off[z], dst = zone_prefix[z];
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
            okey[dst + i] = skey[src + i];
            osign[dst + i] = ssign[src + i];
            for (u32 l = 0; l < limbs; ++l) oplanes[(u64)l * ostride + dst + i] = splanes[(u64)l * sstride + src + i];
        }
    }
}

// exclusive scan of 32-bit counts into 64-bit offsets (single block); total -> *total_out
__global__ void __launch_bounds__(1024) k_zone_scan(const u32 *__restrict__ counts, u64 *__restrict__ offsets, u32 n,
                                                    u64 *total_out)
{
    __shared__ u32 s_ws[40];
    __shared__ u64 s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < n ? counts[i] : 0;
        u32 total;
        const u32 ex = block_excl_scan(v, s_ws, total);
        if (i < n) offsets[i] = s_run + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_run += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_run;
}

// ------------------------------------------------------------------------------------------------ host side

inline u32 env_u32(const char *name, u32 dflt)
{
    const char *v = getenv(name);
    return v ? (u32)strtoul(v, nullptr, 10) : dflt;
}

template <int NW>
void zone_set_smem(size_t bytes)
{
    cudaFuncSetAttribute(k_zone<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline void zone_configure(pk_ctx *ctx)
{
    ZoneTuning &t = ctx->zone_tuning;
    t.target_zones = env_u32("PK_ZONE_TARGET", t.target_zones);
    t.zones_per_chunk = std::max<u32>(1, env_u32("PK_ZONE_CHUNK", t.zones_per_chunk));
    t.span_log2_max = std::min<u32>(20, std::max<u32>(10, env_u32("PK_ZONE_SPAN_LOG2", t.span_log2_max)));
    t.force_cap = env_u32("PK_ZONE_CAP", t.force_cap);
    t.force_mode = env_u32("PK_ZONE_MODE", t.force_mode);
    t.disable = env_u32("PK_ZONE_DISABLE", t.disable);
    const size_t dyn = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
    zone_set_smem<2>(dyn);
    zone_set_smem<3>(dyn);
    zone_set_smem<4>(dyn);
    zone_set_smem<5>(dyn);
    zone_set_smem<6>(dyn);
}

struct ZonePlan {
    bool ok = false;
    u32 NW = 0, mode = 0, cap = 0, use_cursors = 0, bm_words = 0, n_zones = 0, zones_per_chunk = 1, n_chunks = 0;
    u64 span = 0;
    u32 off_cursor = 0, off_acc = 0, off_bm = 0, off_pre = 0;
    size_t smem = 0;
};

inline ZonePlan zone_plan(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB, u64 sub_range)
{
    const ZoneTuning &t = ctx->zone_tuning;
    ZonePlan zp;
    if (t.disable || pl.LA != 1 || pl.LB != 1 || sub_range == 0) return zp;
    u32 NW = (pl.prod_bits + pl.cnt_bits + 1 + 31) / 32;
    NW = std::max<u32>(NW, 2);
    if (NW > 6) return zp;
    zp.NW = NW;
    const size_t budget = (ctx->smem_optin > 4096 ? ctx->smem_optin - 2048 : 0);
    if (budget < (64u << 10)) return zp;
    const size_t cursor_bytes = ((nA * 4 + 15) & ~15ull);
    zp.use_cursors = cursor_bytes <= t.cursor_bytes_max ? 1 : 0;
    const size_t avail = budget - (zp.use_cursors ? cursor_bytes : 0);
    const u32 target = t.target_zones ? t.target_zones : (u32)ctx->sm_count * 16;
    const u64 dense_cap_max = (avail / (NW * 4)) & ~31ull;
    // dense zones sized for the target zone count, but never beyond what shared memory holds
    u64 span_d = std::max<u64>(64, (sub_range + target - 1) / target);
    span_d = (span_d + 31) & ~31ull;
    if (span_d > dense_cap_max) span_d = dense_cap_max;
    const u64 zones_d = (sub_range + span_d - 1) / span_d;
    // bitmap zones
    u64 span_b = 1ull << t.span_log2_max;
    while (span_b > 1024 && (sub_range + span_b - 1) / span_b < target) span_b >>= 1;
    const size_t bm_bytes = span_b / 8, pre_bytes = span_b / 8;
    const u64 zones_b = (sub_range + span_b - 1) / span_b;
    const bool bitmap_fits = avail > bm_bytes + pre_bytes + (size_t)NW * 4 * 256;
    // cost model: a row visit costs about half a product; bitmap zones walk twice
    const double W = (double)nA * (double)nB;
    const double cost_d = (double)zones_d * (double)nA * 0.5 + W;
    const double cost_b = bitmap_fits ? (double)zones_b * (double)nA * 1.0 + W * 1.6 : 1e300;
    u32 mode = cost_d <= cost_b ? 0 : 1;
    if (t.force_mode == 1) mode = 0;
    if (t.force_mode == 2 && bitmap_fits) mode = 1;
    zp.mode = mode;
    size_t off = 0;
    zp.off_cursor = 0;
    off += zp.use_cursors ? cursor_bytes : 0;
    if (mode == 0) {
        zp.span = span_d;
        zp.cap = (u32)span_d;
        zp.n_zones = (u32)zones_d;
        if (zones_d > t.max_zones) return zp;
        zp.off_acc = (u32)off;
        off += (size_t)zp.cap * NW * 4;
    } else {
        zp.span = span_b;
        zp.bm_words = (u32)(span_b / 32);
        zp.n_zones = (u32)zones_b;
        if (zones_b > t.max_zones) return zp;
        zp.off_bm = (u32)off;
        off += bm_bytes;
        zp.off_pre = (u32)off;
        off += pre_bytes;
        zp.off_acc = (u32)off;
        u64 cap = ((budget - off) / (NW * 4)) & ~31ull;
        cap = std::min<u64>(cap, span_b);
        zp.cap = (u32)cap;
        off += (size_t)zp.cap * NW * 4;
    }
    if (t.force_cap) {
        if (mode == 0) { // dense: shrinking the capacity means shrinking the zone
            zp.cap = std::max<u32>(32, t.force_cap & ~31u);
            zp.span = zp.cap;
            zp.n_zones = (u32)((sub_range + zp.span - 1) / zp.span);
        } else {
            zp.cap = std::max<u32>(1, std::min<u32>(zp.cap, t.force_cap));
        }
    }
    zp.smem = off;
    zp.zones_per_chunk = zp.use_cursors ? t.zones_per_chunk : 1;
    zp.n_chunks = (zp.n_zones + zp.zones_per_chunk - 1) / zp.zones_per_chunk;
    zp.ok = true;
    return zp;
}

inline u32 choose_algo(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB, u64 sub_range, bool trunc)
{
    if (trunc) return PK_ALGO_AUTO_FALLBACK;
    if ((unsigned __int128)nA * nB < (1u << 16)) return PK_ALGO_AUTO_FALLBACK; // launch-bound anyway
    return zone_plan(ctx, pl, nA, nB, sub_range).ok ? PK_ALGO_ZONE : PK_ALGO_AUTO_FALLBACK;
}

template <int NW>
void zone_launch(pk_ctx *ctx, const ZoneArgs &za, u32 grid, size_t smem)
{
    PK_LAUNCH(ctx, k_zone<NW>, grid, kZoneThreads, smem, za);
}

// Returns the product restricted to [lo, hi), or nullptr when the configuration is outside the zone path's envelope.
inline pk_dpoly *zone_multiply(pk_ctx *ctx, const Plan &pl, const pk_dpoly *A, const pk_dpoly *B, const u64 *keyA,
                               const u64 *keyB, const u64 *magB, u64 strideB, const u32 *sl, u64 lo, u64 hi, bool trunc,
                               u32 nv)
{
    (void)strideB;
    if (trunc || sl) return nullptr;
    const u64 nA = A->n, nB = B->n;
    const ZonePlan zp = zone_plan(ctx, pl, nA, nB, hi - lo);
    if (!zp.ok) return nullptr;
    const unsigned __int128 w_bound = (unsigned __int128)nA * nB;
    const u64 out_cap = (u64)std::min<unsigned __int128>(w_bound, hi - lo);
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    if ((unsigned __int128)out_cap * (9 + 8 * pl.out_limbs) > (unsigned __int128)total_b / 2) return nullptr;

    DevBuf pa(ctx, (nA + 1) * 16), pb(ctx, (nB + 1) * 16);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nA, 256), 256, 0, keyA, A->planes, pa.as<ulonglong2>(), nA);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nB, 256), 256, 0, keyB, magB, pb.as<ulonglong2>(), nB);
    // per-zone bookkeeping: offsets (u64), prefix (u64), counts (u32), then the bump cursor and the chunk counter
    const size_t nz = zp.n_zones;
    DevBuf book(ctx, nz * 20 + 64);
    u64 *zone_off = book.as<u64>();
    u64 *zone_prefix = zone_off + nz;
    u64 *bump = zone_prefix + nz;
    u32 *chunk_counter = reinterpret_cast<u32 *>(bump + 1);
    u32 *zone_count = chunk_counter + 2;
    CUDA_CHECK(cudaMemsetAsync(bump, 0, 16, ctx->stream));

    DPolyGuard g{ctx, new_dpoly(ctx, out_cap, nv, pl.out_limbs)};
    ZoneArgs za{};
    za.A = pa.as<ulonglong2>();
    za.nA = nA;
    za.B = pb.as<ulonglong2>();
    za.nB = nB;
    za.lo = lo;
    za.hi = hi;
    za.span = zp.span;
    za.n_zones = zp.n_zones;
    za.zones_per_chunk = zp.zones_per_chunk;
    za.n_chunks = zp.n_chunks;
    za.mode = zp.mode;
    za.cap = zp.cap;
    za.use_cursors = zp.use_cursors;
    za.pw = (pl.prod_bits + 31) / 32;
    za.bm_words = zp.bm_words;
    za.off_cursor = zp.off_cursor;
    za.off_acc = zp.off_acc;
    za.off_bm = zp.off_bm;
    za.off_pre = zp.off_pre;
    za.okey = g.p->key;
    za.oplanes = g.p->planes;
    za.ostride = g.p->stride;
    za.osign = g.p->sign;
    za.out_limbs = pl.out_limbs;
    za.out_cap = out_cap;
    za.cp = pl.cpR;
    za.base_code = pl.base_code;
    za.chunk_counter = chunk_counter;
    za.zone_count = zone_count;
    za.zone_off = zone_off;
    za.bump = bump;
    za.ctr = ctx->d_ctr;
    const u32 grid = std::min<u32>((u32)ctx->sm_count, zp.n_chunks);
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
    switch (zp.NW) {
    case 2: zone_launch<2>(ctx, za, grid, zp.smem); break;
    case 3: zone_launch<3>(ctx, za, grid, zp.smem); break;
    case 4: zone_launch<4>(ctx, za, grid, zp.smem); break;
    case 5: zone_launch<5>(ctx, za, grid, zp.smem); break;
    default: zone_launch<6>(ctx, za, grid, zp.smem); break;
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
    PK_LAUNCH(ctx, k_zone_scan, 1, 1024, 0, zone_count, zone_prefix, zp.n_zones, &ctx->d_ctr->n_out);
    MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
    CUDA_CHECK(cudaMemcpyAsync(hc, ctx->d_ctr, sizeof(MulCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const u64 n_out = hc->n_out;
    if (n_out > out_cap || (hc->flags & kFlagTableFull)) fail(PK_E_CUDA, "zone path: result count exceeds its bound (internal error)");
    ctx->stats.table_slots = (u64)zp.cap;
    ctx->stats.acc_lanes = zp.NW;
    ctx->stats.chunk_bits = 0;
    ctx->stats.retries = zp.mode; // 0 dense zones, 1 bitmap zones (reported through the spare counter)
    DPolyGuard e{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
    if (n_out) {
        const u32 ggrid = std::min<u32>(zp.n_zones, (u32)ctx->sm_count * 16);
        PK_LAUNCH(ctx, k_zone_gather, ggrid, 256, 0, zone_count, zone_off, zone_prefix, zp.n_zones, g.p->key, g.p->planes,
                  g.p->stride, g.p->sign, e.p->key, e.p->planes, e.p->stride, e.p->sign, pl.out_limbs);
    }
    return e.release();
}

} // namespace pk
