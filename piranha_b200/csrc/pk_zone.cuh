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


// Result key codec, slimmed for the kernel parameter space: tight code -> piranha code by repeated division with
// precomputed 64-bit magic multipliers (exact for 32-bit codes), digit * (step * piranha stride) summed up.
struct ZoneCodec {
    u32 n_vars;
    u32 pad;
    i64 base_code;
    u32 tradix[PK_MAX_VARS];
    u64 magic[PK_MAX_VARS];     // floor(2^64 / tradix) + 1
    i64 spstride[PK_MAX_VARS];  // step * piranha stride
};

struct ZoneArgs {
    const uint4 *A; // {tight code (32 bit), sign flag, magnitude lo, magnitude hi}, ascending code
    u32 nA;
    const uint4 *B;
    u32 nB;
    const uint2 *A1, *B1; // block path, two-limb operands: the second limb of every magnitude (else unused)
    u32 lo, hi; // output partition of this call (tight codes, < 2^32)
    const u32 *cuts;        // zone z covers codes [cuts[z], cuts[z+1]): equal-work cuts built on the device
    const u32 *n_zones_dev; // number of zones (device-side value, decided by k_zone_cuts)
    u32 zones_per_chunk;
    u32 cap; // accumulator entries per zone pass
    u32 use_cursors;
    u32 pw; // 32-bit words of a product magnitude
    u32 bm_words;
    u32 *cursors;      // use_cursors: nA row cursors per block, in global memory (L1/L2 resident, coalesced per batch)
    u32 cursor_stride; // words per block
    u32 off_acc, off_bm, off_pre, off_list; // byte offsets into dynamic shared memory
    i64 *okey;
    u64 *oplanes;
    u64 ostride;
    signed char *osign;
    u32 out_limbs;
    u64 out_cap;
    ZoneCodec cd;
    u32 *chunk_counter;
    u64 *trace;             // optional: per chunk {start ns, end ns, smid, products} (PK_ZONE_TRACE)
    u32 *zone_count; // result terms of zone z
    u64 *zone_off;   // where zone z starts in the staging arrays
    u64 *bump;       // staging allocation cursor
    MulCounters *ctr;
};

// prepared key (tight | sign bit) + magnitude -> packed 16-byte operand record
__global__ void k_pack_operand(const u64 *__restrict__ key, const u64 *__restrict__ m0, uint4 *__restrict__ out, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 k = key[i], m = m0[i];
        out[i] = make_uint4((u32)(k & 0xFFFFFFFFull), (u32)(k >> 63), (u32)m, (u32)(m >> 32));
    }
}

// first index with T[idx].x >= x
__device__ __forceinline__ u32 zone_lower_bound(const uint4 *__restrict__ T, u32 n, u32 x)
{
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (__ldg(&T[mid].x) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ u32 atoms_add(u32 saddr, u32 v)
{
    u32 old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add(u32 saddr, u32 v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_or(u32 saddr, u32 v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// Accumulator word w of entry e lives at saddr0 + w * wstride (word planes).  The 128-bit magnitude p0..p3 is
// added (subtracted when neg) word by word; the value returned by each ATOMS gives the carry (borrow) into the
// next word.  Words below pw always issue their atomic (branch-free); above pw only a pending carry travels on.
template <int NW, bool SIGNED>
__device__ __forceinline__ void zone_accumulate(u32 saddr0, u32 wstride, u32 p0, u32 p1, u32 p2, u32 p3, u32 pw, bool neg)
{
    const u32 p[4] = {p0, p1, p2, p3};
    u32 c = 0;
    u32 addr = saddr0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const u32 pv = (w < 4) ? p[w < 4 ? w : 0] : 0u;
        if ((u32)w >= pw && c == 0) break;
        const u32 x = pv + c;
        const u32 ovf = (x < c) ? 1u : 0u; // pv == 0xffffffff and c == 1: the word is unchanged, the carry moves on
        if (w == NW - 1) {
            reds_add(addr, (SIGNED && neg) ? 0u - x : x); // top word: its carry-out is the proven-impossible overflow
        } else if (SIGNED && neg) {
            const u32 old = atoms_add(addr, 0u - x);
            c = ((old < x) ? 1u : 0u) | ovf;
        } else {
            const u32 old = atoms_add(addr, x);
            c = ((old + x < x) ? 1u : 0u) | ovf;
        }
        addr += wstride;
    }
}

__device__ __forceinline__ i64 zone_code_to_piranha(u32 code, const ZoneCodec &cd)
{
    i64 out = cd.base_code;
    u32 c = code;
    for (u32 v = 0; v < cd.n_vars; ++v) {
        const u32 r = cd.tradix[v];
        if (r > 1) {
            const u32 q = (u32)__umul64hi((u64)c, cd.magic[v]);
            const u32 d = c - q * r;
            c = q;
            out += (i64)d * cd.spstride[v];
        }
    }
    return out;
}

// Two's complement NW words -> sign, magnitude; writes the term with Kronecker code `key` at `pos`.  SIGNED = false: the
// accumulator cannot be negative (operands without negative coefficients).
template <int NW, bool SIGNED = true>
__device__ __forceinline__ void zone_emit_key(const ZoneArgs &a, u64 pos, i64 key, const u32 (&wd)[NW], u32 (&orw)[NW], u32 &flags)
{
    u32 m[NW];
    const bool neg = SIGNED && (wd[NW - 1] >> 31) != 0;
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
#pragma unroll
    for (int l = 0; l < (NW + 1) / 2; ++l) {
        const u64 limb = (u64)m[2 * l] | ((2 * l + 1 < NW) ? ((u64)m[(2 * l + 1 < NW) ? 2 * l + 1 : 0] << 32) : 0ull);
        if ((u32)l < a.out_limbs) a.oplanes[(u64)l * a.ostride + pos] = limb;
        else if (limb) flags |= kFlagWidth;
    }
    for (u32 l = (NW + 1) / 2; l < a.out_limbs; ++l) a.oplanes[(u64)l * a.ostride + pos] = 0;
    a.okey[pos] = key;
    a.osign[pos] = neg ? (signed char)-1 : (signed char)1;
    // the largest bit length over all terms = the position of the highest bit set in ANY magnitude: OR the words per
    // position (orw), the caller turns the ORs into a bit length once at the end (zone_maxbits)
#pragma unroll
    for (int w = 0; w < NW; ++w) orw[w] |= m[w];
}

template <int NW>
__device__ __forceinline__ u32 zone_maxbits(const u32 (&orw)[NW])
{
    u32 bits = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        if (orw[w]) bits = 32u * (u32)w + (32u - (u32)__clz((int)orw[w]));
    return bits;
}

template <int NW, bool SIGNED = true>
__device__ __forceinline__ void zone_emit(const ZoneArgs &a, u64 pos, u32 code, const u32 (&wd)[NW], u32 &maxbits, u32 &flags)
{
    u32 orw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) orw[w] = 0;
    zone_emit_key<NW, SIGNED>(a, pos, zone_code_to_piranha(code, a.cd), wd, orw, flags);
    maxbits = max(maxbits, zone_maxbits<NW>(orw));
}

// Tight code -> Kronecker code for ASCENDING codes: the digits of the previous code are kept (first kKeyWalkVars digits in
// registers, the rest folded into a base) and a step adds the difference to digit 0 and carries on; most steps touch
// one or two digits instead of all of them.
constexpr int kKeyWalkVars = 8;
struct KeyWalker {
    u32 d[kKeyWalkVars];
    u32 code;
    i64 key;
    __device__ __forceinline__ void init(u32 c0, const ZoneCodec &cd)
    {
        code = c0;
        key = cd.base_code;
        u32 c = c0;
#pragma unroll
        for (int v = 0; v < kKeyWalkVars; ++v) {
            d[v] = 0;
            if ((u32)v < cd.n_vars) {
                const u32 r = cd.tradix[v];
                if (r > 1) {
                    const u32 q = (u32)__umul64hi((u64)c, cd.magic[v]);
                    d[v] = c - q * r;
                    c = q;
                    key += (i64)d[v] * cd.spstride[v];
                }
            }
        }
        for (u32 v = kKeyWalkVars; v < cd.n_vars; ++v) { // (more variables than the walker keeps: their digits never change within its reach)
            const u32 r = cd.tradix[v];
            if (r > 1) {
                const u32 q = (u32)__umul64hi((u64)c, cd.magic[v]);
                key += (i64)(c - q * r) * cd.spstride[v];
                c = q;
            }
        }
    }
    // advance to c1 >= code: most steps stay inside digit 0 (one add, one 32 x 64 multiply); a step that carries out of it
    // decodes from scratch
    __device__ __forceinline__ void step(u32 c1, const ZoneCodec &cd)
    {
        const u32 delta = c1 - code;
        const u32 x = d[0] + delta;
        if (cd.tradix[0] > 1u && delta < cd.tradix[0] && x < cd.tradix[0]) {
            key += (i64)delta * cd.spstride[0];
            d[0] = x;
            code = c1;
        } else {
            init(c1, cd);
        }
    }
};

// MODE 0: dense zones (slot = code - zlo).  MODE 1: bitmap zones (slot = rank of the code among the codes that occur).
// TPB threads per block, 1024 / TPB blocks per SM: while one block sits in a barrier phase (prefix, emit, walk tail)
// the others keep the issue slots busy.
template <int NW, int MODE, bool SIGNED, int TPB>
__global__ void __launch_bounds__(TPB, 1024 / TPB) k_zone(const ZoneArgs a)
{
    constexpr u32 kZoneThreads = TPB;
    extern __shared__ __align__(16) unsigned char zsmem[];
    u32 *cursor = a.cursors + (size_t)blockIdx.x * a.cursor_stride;
    u32 *acc = reinterpret_cast<u32 *>(zsmem + a.off_acc);
    u32 *bm = reinterpret_cast<u32 *>(zsmem + a.off_bm);
    u32 *pre = reinterpret_cast<u32 *>(zsmem + a.off_pre);
    u32 *list = reinterpret_cast<u32 *>(zsmem + a.off_list);
    __shared__ u32 s_ws[40];
    __shared__ u64 s_b64[2];
    __shared__ u32 s_b32[4];
    __shared__ u32 s_next; // dynamic row-batch counter

    const u32 tid = threadIdx.x, lane = tid & 31u;
    const u32 nA = a.nA, nB = a.nB;
    const u32 kb0 = __ldg(&a.B[0].x), kbmax = __ldg(&a.B[nB - 1].x);
    const u32 acc_s = smem_u32(acc), bm_s = smem_u32(bm);
    const u32 wstride = a.cap * 4u;
    u64 my_pairs = 0;
    u32 my_maxbits = 0, my_flags = 0;
    const u32 n_zones = __ldg(a.n_zones_dev);
    const u32 n_chunks = (n_zones + a.zones_per_chunk - 1) / a.zones_per_chunk;

    // accumulators (and bitmap) start cleared; every pass leaves them cleared again
    for (u32 i = tid; i < a.cap * NW; i += kZoneThreads) acc[i] = 0;
    if (MODE == 1)
        for (u32 i = tid; i < a.bm_words; i += kZoneThreads) bm[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_b32[0] = atomicAdd(a.chunk_counter, 1u);
        __syncthreads();
        const u32 claimed = s_b32[0];
        __syncthreads();
        if (claimed >= n_chunks) break;
        const u32 chunk = claimed;
        u64 t_start = 0;
        if (a.trace && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
        const u64 pairs_before = my_pairs;
        const u32 z0 = chunk * a.zones_per_chunk, z1 = min(n_zones, z0 + a.zones_per_chunk);
        if (a.use_cursors) { // one binary search per row per chunk
            const u32 clo = __ldg(&a.cuts[z0]);
            for (u32 i = tid; i < nA; i += kZoneThreads) {
                const u32 ka = __ldg(&a.A[i].x);
                cursor[i] = (ka >= clo) ? 0u : zone_lower_bound(a.B, nB, clo - ka);
            }
            __syncthreads();
        }
        for (u32 z = z0; z < z1; ++z) {
            const u32 zlo = __ldg(&a.cuts[z]), zhi = __ldg(&a.cuts[z + 1]);
            const u32 zwords = (MODE == 1) ? min(a.bm_words, ((zhi - zlo + 127u) >> 7) << 2) : 0u; // bitmap words in use
            // rows that can reach the zone: a_i + b_max >= zlo and a_i + b_0 < zhi
            const u32 i_beg = (zlo > kbmax) ? zone_lower_bound(a.A, nA, zlo - kbmax) : 0u;
            const u32 i_end = (zhi > kb0) ? zone_lower_bound(a.A, nA, zhi - kb0) : 0u;

            // ---------------- walk 1 (bitmap zones): mark the codes that occur
            u32 zone_count = 0;
            if (MODE == 1) {
                if (tid == 0) s_next = i_beg;
                __syncthreads();
                for (;;) {
                    u32 r = 0;
                    if (lane == 0) r = atomicAdd(&s_next, 32u);
                    r = __shfl_sync(0xffffffffu, r, 0);
                    if (r >= i_end) break;
                    const u32 i = r + lane;
                    if (i < i_end) {
                        const u32 ka = __ldg(&a.A[i].x);
                        u32 j = a.use_cursors ? cursor[i] : ((ka >= zlo) ? 0u : zone_lower_bound(a.B, nB, zlo - ka));
                        while (j < nB) {
                            const u32 code = ka + __ldg(&a.B[j].x);
                            if (code >= zhi) break;
                            const u32 slot = code - zlo;
                            reds_or(bm_s + (slot >> 5) * 4u, 1u << (slot & 31u));
                            ++j;
                        }
                    }
                }
                __syncthreads();
                // exclusive popcount prefix per group of 4 bitmap words (128 codes):
                // rank(code) = pre[group] + popc(words of the group below) + popc(bits below in its word)
                const u32 n_groups = zwords >> 2;
                const u32 gpt = (n_groups + kZoneThreads - 1) / kZoneThreads;
                const u32 g0 = min(n_groups, tid * gpt), g1 = min(n_groups, g0 + gpt);
                u32 local = 0;
                for (u32 g = g0; g < g1; ++g) {
                    const uint4 w4 = reinterpret_cast<const uint4 *>(bm)[g];
                    local += __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
                }
                u32 total;
                u32 run = block_excl_scan(local, s_ws, total);
                for (u32 g = g0; g < g1; ++g) {
                    const uint4 w4 = reinterpret_cast<const uint4 *>(bm)[g];
                    pre[g] = run;
                    run += __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
                }
                zone_count = total;
                __syncthreads();
            }

            // ---------------- accumulate + emit, in passes of at most `cap` entries
            const u32 n_entries = (MODE == 1) ? zone_count : (zhi - zlo);
            const bool single = n_entries <= a.cap;
            u64 zone_base = 0, zone_written = 0;
            bool have_base = false;
            u32 r0 = 0;   // first rank (bitmap) / slot (dense) of this pass
            u32 c0 = zlo; // first code of this pass
            do {
                const u32 r1 = min(n_entries, r0 + a.cap);
                u32 c1 = zhi;
                if (MODE == 1) {
                    const u32 n_groups = zwords >> 2;
                    if (r1 < n_entries) { // the code of rank r1 bounds this pass (zone with more entries than cap)
                        if (tid == 0) {
                            u32 lo_g = 0, hi_g = n_groups; // last group with pre[g] <= r1
                            while (hi_g - lo_g > 1) {
                                const u32 mid = (lo_g + hi_g) >> 1;
                                if (pre[mid] <= r1) lo_g = mid;
                                else hi_g = mid;
                            }
                            u32 rem = r1 - pre[lo_g], w = lo_g * 4;
                            while (rem >= (u32)__popc(bm[w])) {
                                rem -= __popc(bm[w]);
                                ++w;
                            }
                            s_b32[1] = zlo + w * 32u + __fns(bm[w], 0, (int)rem + 1);
                        }
                        __syncthreads();
                        c1 = s_b32[1];
                    }
                    // rank -> slot list of this pass, so that emission runs with full warps
                    const u32 gpt = (n_groups + kZoneThreads - 1) / kZoneThreads;
                    const u32 g0 = min(n_groups, tid * gpt), g1 = min(n_groups, g0 + gpt);
                    for (u32 g = g0; g < g1; ++g) {
                        u32 rank = pre[g];
                        for (u32 q = 0; q < 4; ++q) {
                            u32 bits = bm[g * 4 + q];
                            while (bits) {
                                const u32 b = __ffs(bits) - 1;
                                bits &= bits - 1;
                                if (rank >= r0 && rank < r1) list[rank - r0] = (g * 4 + q) * 32u + b;
                                ++rank;
                            }
                        }
                    }
                }
                if (tid == 0) s_next = i_beg;
                __syncthreads();
                for (;;) {
                    u32 r = 0;
                    if (lane == 0) r = atomicAdd(&s_next, 32u);
                    r = __shfl_sync(0xffffffffu, r, 0);
                    if (r >= i_end) break;
                    const u32 i = r + lane;
                    if (i < i_end) {
                        const uint4 av = __ldg(&a.A[i]);
                        const u32 ka = av.x;
                        const u64 ma = (u64)av.z | ((u64)av.w << 32);
                        u32 j = a.use_cursors ? cursor[i] : ((ka >= c0) ? 0u : zone_lower_bound(a.B, nB, c0 - ka));
                        u32 done = 0;
                        while (j < nB) {
                            const uint4 bv = __ldg(&a.B[j]);
                            const u32 code = ka + bv.x;
                            if (code >= c1) break;
                            const u32 slot = code - zlo;
                            u32 e;
                            if (MODE == 1) {
                                const u32 g = slot >> 7, q = (slot >> 5) & 3u;
                                const uint4 w4 = reinterpret_cast<const uint4 *>(bm)[g];
                                const u32 wq = q == 0 ? w4.x : (q == 1 ? w4.y : (q == 2 ? w4.z : w4.w));
                                u32 rk = pre[g] + __popc(wq & ((1u << (slot & 31u)) - 1u));
                                rk += (q > 0 ? __popc(w4.x) : 0) + (q > 1 ? __popc(w4.y) : 0) + (q > 2 ? __popc(w4.z) : 0);
                                e = rk - r0;
                            } else {
                                e = slot - r0;
                            }
                            const u64 mb = (u64)bv.z | ((u64)bv.w << 32);
                            const u64 plo = ma * mb, phi = __umul64hi(ma, mb);
                            zone_accumulate<NW, SIGNED>(acc_s + e * 4u, wstride, (u32)plo, (u32)(plo >> 32), (u32)phi,
                                                        (u32)(phi >> 32), a.pw, SIGNED && ((av.y ^ bv.y) != 0));
                            ++j;
                            ++done;
                        }
                        my_pairs += done;
                        if (a.use_cursors) cursor[i] = j;
                    }
                }
                __syncthreads();

                // ---- emit the entries [0, n_pass) of this pass in ascending code order, 1024 at a time
                const u32 n_pass = r1 - r0;
                if (!have_base) { // claim staging space: the exact count of a single-pass zone, else the entry count
                    u32 nz = 0;
                    if (single) {
                        for (u32 e = tid; e < n_pass; e += kZoneThreads) {
                            u32 any = 0;
#pragma unroll
                            for (int q = 0; q < NW; ++q) any |= acc[q * a.cap + e];
                            nz += any != 0;
                        }
                    }
                    u32 total_nz;
                    block_excl_scan(nz, s_ws, total_nz);
                    if (tid == 0) s_b64[1] = atomicAdd(a.bump, single ? (u64)total_nz : (u64)n_entries);
                    __syncthreads();
                    zone_base = s_b64[1];
                    have_base = true;
                }
                for (u32 base = 0; base < n_pass; base += kZoneThreads) {
                    const u32 e = base + tid;
                    u32 wd[NW], any = 0;
                    if (e < n_pass) {
#pragma unroll
                        for (int q = 0; q < NW; ++q) {
                            wd[q] = acc[q * a.cap + e];
                            acc[q * a.cap + e] = 0;
                            any |= wd[q];
                        }
                    }
                    u32 total;
                    const u32 excl = block_excl_scan(any != 0 ? 1u : 0u, s_ws, total);
                    if (any) {
                        const u32 slot = (MODE == 1) ? list[e] : (r0 + e);
                        zone_emit<NW>(a, zone_base + zone_written + excl, zlo + slot, wd, my_maxbits, my_flags);
                    }
                    zone_written += total;
                }
                r0 = r1;
                c0 = c1;
                __syncthreads();
            } while (r0 < n_entries);
            if (tid == 0) {
                a.zone_count[z] = (u32)zone_written;
                a.zone_off[z] = zone_base;
            }
            if (MODE == 1) { // clear the bitmap for the next zone
                for (u32 i = tid; i < zwords; i += kZoneThreads) bm[i] = 0;
            }
            __syncthreads();
        }
        if (a.trace) {
            if (tid == 0) {
                u64 t_end;
                u32 smid;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
                asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
                a.trace[4 * (size_t)chunk + 0] = t_start;
                a.trace[4 * (size_t)chunk + 1] = t_end;
                a.trace[4 * (size_t)chunk + 2] = ((u64)claimed << 32) | smid;
            }
            if (tid == 32) a.trace[4 * (size_t)chunk + 3] = my_pairs - pairs_before; // one thread's share: relative weight
        }
    }
    // per-thread statistics
    for (int o = 16; o > 0; o >>= 1) {
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
        my_maxbits = max(my_maxbits, __shfl_xor_sync(0xffffffffu, my_maxbits, o));
        my_flags |= __shfl_xor_sync(0xffffffffu, my_flags, o);
    }
    if (lane == 0) {
        if (my_pairs) atomicAdd(&a.ctr->pairs, my_pairs);
        if (my_maxbits) atomicMax(&a.ctr->max_bits, my_maxbits);
        if (my_flags) atomicOr(&a.ctr->flags, my_flags);
    }
}

// Zone work is very uneven over the code range (multinomial coefficients pile up products in the middle: measured
// max/median chunk time 13x with equal-span zones, profiles/), so zones are cut at WORK quantiles: a hashed sample
// of pair sums is radix-sorted and every (S/Z)-th sample becomes a zone boundary; intervals wider than what the
// accumulator holds are split evenly.  This is the role of piranha's zone table (polynomial.hpp:1788-1867), with
// zones equalised by pair count instead of bucket count.
__global__ void k_zone_sample(const uint4 *__restrict__ A, u32 nA, const uint4 *__restrict__ B, u32 nB, u32 S, u32 lo, u32 hi,
                              u32 *__restrict__ out)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    const u64 h = mix64(0x9E3779B97F4A7C15ull * (t + 1));
    const u32 ia = (u32)((h & 0xFFFFFFFFull) * nA >> 32), jb = (u32)((h >> 32) * nB >> 32);
    const u32 code = __ldg(&A[ia].x) + __ldg(&B[jb].x);
    out[t] = (code >= lo && code < hi) ? code : 0xFFFFFFFFu;
}

// One block: boundaries b_k (k = 0..Z) = lo, the work quantiles, hi; interval k is split into ceil(len / max_span)
// zones.  Writes cuts[0..n] and n.  `uniform` != 0 replaces the quantiles by equal spans (tests, comparison).
__global__ void __launch_bounds__(1024) k_zone_cuts(const u32 *__restrict__ sorted, u32 S, u32 lo, u32 hi, u32 Z, u32 max_span,
                                                    u32 uniform, u32 max_zones, u32 *__restrict__ cuts, u32 *__restrict__ n_zones)
{
    __shared__ u32 s_ws[40];
    __shared__ u32 s_valid;
    if (threadIdx.x == 0) { // number of in-range samples (the rest sorted to the end as 0xFFFFFFFF)
        u32 l = 0, h = S;
        while (l < h) {
            const u32 m = (l + h) >> 1;
            if (sorted[m] < 0xFFFFFFFFu) l = m + 1;
            else h = m;
        }
        s_valid = l;
    }
    __syncthreads();
    const u32 V = s_valid;
    const u32 ustep = (u32)(((u64)(hi - lo) + Z - 1) / Z);
    auto boundary = [&](u32 k) -> u32 {
        if (k == 0) return lo;
        if (k >= Z) return hi;
        if (uniform || V == 0) return (u32)min((u64)hi, (u64)lo + (u64)k * ustep);
        return min(hi, max(lo, sorted[(u32)((u64)k * V / Z)]));
    };
    u32 run = 0;
    for (u32 base = 0; base < Z; base += 1024) {
        const u32 k = base + threadIdx.x;
        u32 b0 = 0, len = 0, pieces = 0;
        if (k < Z) {
            b0 = boundary(k);
            const u32 b1 = boundary(k + 1);
            len = b1 > b0 ? b1 - b0 : 0;
            pieces = len ? (len + max_span - 1) / max_span : 0;
        }
        u32 total;
        const u32 excl = block_excl_scan(pieces, s_ws, total);
        if (pieces) {
            const u32 step = (len + pieces - 1) / pieces;
            for (u32 q = 0; q < pieces; ++q)
                if (run + excl + q < max_zones) cuts[run + excl + q] = b0 + q * step;
        }
        run += total;
    }
    if (threadIdx.x == 0) {
        const u32 n = min(run, max_zones);
        cuts[n] = hi;
        *n_zones = n;
    }
}

// Lay the zones out in zone order (= ascending code order).  One thread per OUTPUT term (so every write is coalesced
// and the grid covers the whole result whatever the zone sizes): a block finds the zone of its first term by binary
// search over the scanned zone offsets, then every thread walks forward from there.
constexpr u32 kGatherPerBlock = 256; // (one term per thread: a 135 k-term gather is 530 blocks; with 1024 per block it was 133 blocks and 39 us of latency chains)
__global__ void __launch_bounds__(256) k_zone_gather(const u32 *__restrict__ zone_count, const u64 *__restrict__ zone_off,
                                                     const u64 *__restrict__ zone_prefix, const u32 *__restrict__ n_zones_dev,
                                                     const i64 *__restrict__ skey, const u64 *__restrict__ splanes,
                                                     u64 sstride, const signed char *__restrict__ ssign,
                                                     i64 *__restrict__ okey, u64 *__restrict__ oplanes, u64 ostride,
                                                     signed char *__restrict__ osign, u32 limbs, u64 n_out,
                                                     const u64 *__restrict__ n_out_dev = nullptr)
{
    (void)zone_count;
    // (n_out_dev: the destination was sized from a hint; nothing is copied when the device counted a different size)
    if (n_out_dev && *n_out_dev != n_out) return;
    const u32 n_zones = __ldg(n_zones_dev);
    const u64 base = (u64)blockIdx.x * kGatherPerBlock;
    if (base >= n_out || n_zones == 0) return;
    constexpr u32 kPer = kGatherPerBlock / 256;
    u64 src[kPer];
#pragma unroll
    for (u32 k = 0; k < kPer; ++k) { // source positions first, then all loads of the thread in flight at once
        const u64 i = base + k * 256u + threadIdx.x;
        src[k] = ~0ull;
        if (i < n_out) {
            // the zone of output i: the last zone whose first output position is <= i (a search per thread: the upper levels
            // are the same addresses for the whole block and hit L1; walking zone by zone from the block's first zone was a
            // chain of dependent L2 loads, most of the 39 us this kernel took for fateman1's 135 k terms)
            u32 lo = 0, hi = n_zones;
            while (hi - lo > 1) {
                const u32 mid = (lo + hi) >> 1;
                if (__ldg(&zone_prefix[mid]) <= i) lo = mid;
                else hi = mid;
            }
            src[k] = __ldg(&zone_off[lo]) + (i - __ldg(&zone_prefix[lo]));
        }
    }
    i64 kk[kPer];
    signed char sg[kPer];
#pragma unroll
    for (u32 k = 0; k < kPer; ++k)
        if (src[k] != ~0ull) {
            kk[k] = __ldcs(&skey[src[k]]);
            sg[k] = __ldcs(&ssign[src[k]]);
        }
#pragma unroll
    for (u32 k = 0; k < kPer; ++k)
        if (src[k] != ~0ull) {
            const u64 i = base + k * 256u + threadIdx.x;
            okey[i] = kk[k];
            osign[i] = sg[k];
        }
    for (u32 l = 0; l < limbs; ++l) {
        u64 m[kPer];
#pragma unroll
        for (u32 k = 0; k < kPer; ++k)
            if (src[k] != ~0ull) m[k] = __ldcs(&splanes[(u64)l * sstride + src[k]]);
#pragma unroll
        for (u32 k = 0; k < kPer; ++k)
            if (src[k] != ~0ull) oplanes[(u64)l * ostride + base + k * 256u + threadIdx.x] = m[k];
    }
}

// exclusive scan of 32-bit counts into 64-bit offsets (single block); total -> *total_out
__global__ void __launch_bounds__(1024) k_zone_scan(const u32 *__restrict__ counts, u64 *__restrict__ offsets,
                                                    const u32 *__restrict__ n_dev, u64 *total_out)
{
    const u32 n = __ldg(n_dev);
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

template <int NW, int MODE, bool SIGNED, int TPB>
struct ZoneKernel {
    static void set_smem(size_t bytes)
    {
        cudaFuncSetAttribute(k_zone<NW, MODE, SIGNED, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    static void launch(pk_ctx *ctx, const ZoneArgs &za, u32 grid, size_t smem)
    {
        PK_LAUNCH(ctx, (k_zone<NW, MODE, SIGNED, TPB>), grid, TPB, smem, za);
    }
};

template <int NW, int TPB>
void zone_set_smem_nt(size_t b)
{
    ZoneKernel<NW, 0, false, TPB>::set_smem(b);
    ZoneKernel<NW, 0, true, TPB>::set_smem(b);
    ZoneKernel<NW, 1, false, TPB>::set_smem(b);
    ZoneKernel<NW, 1, true, TPB>::set_smem(b);
}

template <int NW>
void zone_set_smem_nw(size_t optin)
{
    zone_set_smem_nt<NW, 1024>(optin); // (256- and 512-thread variants were measured slower and are no longer instantiated)
}

template <int NW, int TPB>
void zone_launch_nt(pk_ctx *ctx, const ZoneArgs &za, u32 grid, size_t smem, u32 mode, bool is_signed)
{
    if (mode == 0) {
        if (is_signed) ZoneKernel<NW, 0, true, TPB>::launch(ctx, za, grid, smem);
        else ZoneKernel<NW, 0, false, TPB>::launch(ctx, za, grid, smem);
    } else {
        if (is_signed) ZoneKernel<NW, 1, true, TPB>::launch(ctx, za, grid, smem);
        else ZoneKernel<NW, 1, false, TPB>::launch(ctx, za, grid, smem);
    }
}

template <int NW>
void zone_launch_nw(pk_ctx *ctx, const ZoneArgs &za, u32 grid, size_t smem, u32 mode, bool is_signed, u32 tpb)
{
    (void)tpb;
    zone_launch_nt<NW, 1024>(ctx, za, grid, smem, mode, is_signed);
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
    t.tpb = env_u32("PK_ZONE_TPB", t.tpb);
    t.uniform = env_u32("PK_ZONE_UNIFORM", t.uniform);
    const size_t dyn = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
    zone_set_smem_nw<2>(dyn);
    zone_set_smem_nw<3>(dyn);
    zone_set_smem_nw<4>(dyn);
    zone_set_smem_nw<5>(dyn);
    zone_set_smem_nw<6>(dyn);
}

struct ZonePlan {
    bool ok = false;
    u32 NW = 0, mode = 0, cap = 0, use_cursors = 0, bm_words = 0, zones_per_chunk = 1;
    u32 max_span = 0;   // widest zone (dense: = cap; bitmap: codes covered by the bitmap)
    u32 target = 0;     // number of work quantiles
    u32 max_zones = 0;  // upper bound on the zone count (device decides the exact number)
    u32 tpb = 1024, grid = 0;
    u32 off_acc = 0, off_bm = 0, off_pre = 0, off_list = 0;
    size_t smem = 0;
};

inline u64 pow2_ceil(u64 x)
{
    u64 p = 1;
    while (p < x) p <<= 1;
    return p;
}

inline ZonePlan zone_plan(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB, u64 sub_range)
{
    const ZoneTuning &t = ctx->zone_tuning;
    ZonePlan zp;
    if (t.disable || pl.LA != 1 || pl.LB != 1 || sub_range == 0) return zp;
    if (pl.range > 0xFFFFFFFEull) return zp; // the kernel works on 32-bit tight codes
    u32 NW = (pl.prod_bits + pl.cnt_bits + 1 + 31) / 32;
    NW = std::max<u32>(NW, 2);
    if (NW > 6) return zp;
    const u32 use_cursors = (nA * 4 <= (u64)t.cursor_bytes_max) ? 1 : 0;
    const u32 target = std::max<u32>(1, t.target_zones ? t.target_zones : (u32)ctx->sm_count * 12);
    const double W = (double)nA * (double)nB;
    const u64 avg_span = (sub_range + target - 1) / target;
    // cost of one candidate (mode, threads per block): a row visit costs about half a product, bitmap zones walk twice
    double best_cost = 1e300;
    for (u32 tpb : {1024u}) { // (one block of 1024 threads per SM: the smaller variants lost every sweep)
        const u32 ctas = 1024 / tpb;
        size_t budget = (ctx->smem_per_sm > ctas * 1024 ? (ctx->smem_per_sm - ctas * 1024) / ctas : 0);
        budget = std::min(budget, ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0);
        budget = budget > 1024 ? (budget - 512) & ~255ull : 0; // static shared memory + alignment slack
        if (budget < (16u << 10)) continue;
        for (u32 mode = 0; mode < 2; ++mode) {
            if (t.force_mode == 1 && mode != 0) continue;
            if (t.force_mode == 2 && mode != 1) continue;
            ZonePlan c;
            c.NW = NW;
            c.use_cursors = use_cursors;
            c.mode = mode;
            c.tpb = tpb;
            c.target = target;
            size_t off = 0;
            double cost;
            if (mode == 0) {
                // light regions may use zones up to 4x the average span, bounded by what shared memory holds
                const u64 cap_max = (budget / (NW * 4)) & ~31ull;
                if (cap_max < 32) continue;
                u64 cap = std::max<u64>(64, (4 * avg_span + 31) & ~31ull);
                cap = std::min(cap, cap_max);
                if (t.force_cap) cap = std::max<u64>(32, std::min<u64>(cap, t.force_cap & ~31u));
                const u64 nz_max = (u64)target + (sub_range + cap - 1) / cap + 2;
                if (nz_max > t.max_zones) continue;
                c.max_span = (u32)cap;
                c.cap = (u32)cap;
                c.max_zones = (u32)nz_max;
                c.off_acc = 0;
                off = (size_t)c.cap * NW * 4;
                const double nz_est = std::max<double>((double)target, (double)sub_range / (double)cap);
                cost = nz_est * (double)nA * 0.5 + W;
            } else {
                u64 span = std::min<u64>(1ull << t.span_log2_max, std::max<u64>(1024, pow2_ceil(4 * avg_span)));
                // bitmap (span / 8 bytes) + prefix per 128 codes (span / 32 bytes) must leave room for entries
                while (span > 1024 && span / 8 + span / 32 + (size_t)(NW + 1) * 4 * 512 > budget) span >>= 1;
                const size_t bm_bytes = span / 8, pre_bytes = span / 32;
                if (bm_bytes + pre_bytes + (size_t)(NW + 1) * 4 * 64 > budget) continue;
                const u64 nz_max = (u64)target + (sub_range + span - 1) / span + 2;
                if (nz_max > t.max_zones) continue;
                c.max_span = (u32)span;
                c.bm_words = (u32)(span / 32);
                c.max_zones = (u32)nz_max;
                c.off_bm = 0;
                off = bm_bytes;
                c.off_pre = (u32)off;
                off += (pre_bytes + 15) & ~15ull;
                u64 cap = ((budget - off - 16) / ((NW + 1) * 4)) & ~31ull; // NW accumulator words + one list word
                cap = std::min<u64>(cap, span);
                if (t.force_cap) cap = std::max<u64>(1, std::min<u64>(cap, t.force_cap));
                c.cap = (u32)cap;
                c.off_list = (u32)off;
                off += ((size_t)c.cap * 4 + 15) & ~15ull;
                c.off_acc = (u32)off;
                off += (size_t)c.cap * NW * 4;
                const double nz_est = std::max<double>((double)target, (double)sub_range / (double)span);
                const double dens = std::min(1.0, W / (double)sub_range);
                const double passes = std::max(1.0, dens * (double)sub_range / nz_est / (double)cap);
                cost = nz_est * (double)nA * (0.5 + 0.5 * passes) + W * 1.6;
            }
            cost *= (tpb == 1024 ? 1.0 : (tpb == 512 ? 1.15 : 1.4)); // measured preference (profiles/sweeps)
            c.smem = off;
            if (cost < best_cost) {
                best_cost = cost;
                zp = c;
                zp.ok = true;
            }
        }
    }
    if (!zp.ok) return zp;
    zp.zones_per_chunk = zp.use_cursors ? t.zones_per_chunk : 1;
    zp.grid = (u32)ctx->sm_count * (1024 / zp.tpb);
    return zp;
}

inline void zone_fill_codec(const Plan &pl, ZoneCodec &cd)
{
    cd.n_vars = pl.cpR.n_vars;
    cd.base_code = pl.base_code;
    for (u32 v = 0; v < pl.cpR.n_vars; ++v) {
        const u64 r = pl.cpR.tradix[v];
        cd.tradix[v] = (u32)r;
        cd.magic[v] = r > 1 ? (u64)((((unsigned __int128)1) << 64) / r) + 1 : 0;
        cd.spstride[v] = (i64)((u64)pl.cpR.step[v] * (u64)pl.cpR.pstride[v]);
    }
}

inline u32 choose_algo(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB, u64 sub_range, bool trunc)
{
    if (trunc) return PK_ALGO_AUTO_FALLBACK;
    if ((unsigned __int128)nA * nB < (1u << 16)) return PK_ALGO_AUTO_FALLBACK; // launch-bound anyway
    return zone_plan(ctx, pl, nA, nB, sub_range).ok ? PK_ALGO_ZONE : PK_ALGO_AUTO_FALLBACK;
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
    // (cudaMemGetInfo is a multi-millisecond host call on this driver: the device size is cached at context creation)
    if ((unsigned __int128)out_cap * (9 + 8 * pl.out_limbs) > (unsigned __int128)ctx->total_mem / 2) return nullptr;

    DevBuf pa(ctx, (nA + 1) * 16), pb(ctx, (nB + 1) * 16);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nA, 256), 256, 0, keyA, A->planes, pa.as<uint4>(), nA);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nB, 256), 256, 0, keyB, magB, pb.as<uint4>(), nB);

    // ---- zone cuts at work quantiles (sample -> sort -> cut builder), all on the device
    const u32 S = (u32)std::min<unsigned __int128>(w_bound, 1u << 18);
    const size_t nz = zp.max_zones;
    // bookkeeping: offsets (u64), prefix (u64) per zone; bump cursor; chunk counter; zone count; counts + cuts (u32)
    DevBuf book(ctx, nz * 24 + 128);
    u64 *zone_off = book.as<u64>();
    u64 *zone_prefix = zone_off + nz;
    u64 *bump = zone_prefix + nz;
    u32 *chunk_counter = reinterpret_cast<u32 *>(bump + 1);
    u32 *n_zones_dev = chunk_counter + 1;
    u32 *zone_count = chunk_counter + 2;
    u32 *cuts = zone_count + nz; // nz + 1 entries
    CUDA_CHECK(cudaMemsetAsync(bump, 0, 16, ctx->stream));
    {
        DevBuf samp(ctx, (size_t)S * 4), sorted(ctx, (size_t)S * 4);
        PK_LAUNCH(ctx, k_zone_sample, grid_for(S, 256), 256, 0, pa.as<uint4>(), (u32)nA, pb.as<uint4>(), (u32)nB, S, (u32)lo,
                  (u32)hi, samp.as<u32>());
        size_t tmp_bytes = 0;
        CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, samp.as<u32>(), sorted.as<u32>(), (int)S, 0, 32, ctx->stream));
        DevBuf tmp(ctx, tmp_bytes);
        CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, samp.as<u32>(), sorted.as<u32>(), (int)S, 0, 32, ctx->stream));
        PK_LAUNCH(ctx, k_zone_cuts, 1, 1024, 0, sorted.as<u32>(), S, (u32)lo, (u32)hi, zp.target, zp.max_span,
                  ctx->zone_tuning.uniform, zp.max_zones, cuts, n_zones_dev);
    }

    const u32 cursor_stride = (u32)((nA + 31) & ~31ull);
    DevBuf cursors(ctx, zp.use_cursors ? (size_t)zp.grid * cursor_stride * 4 : 0);

    const StagingView g = staging_acquire(ctx, out_cap, nv, pl.out_limbs);
    ZoneArgs za{};
    za.A = pa.as<uint4>();
    za.nA = (u32)nA;
    za.B = pb.as<uint4>();
    za.nB = (u32)nB;
    za.lo = (u32)lo;
    za.hi = (u32)hi;
    za.cuts = cuts;
    za.n_zones_dev = n_zones_dev;
    za.zones_per_chunk = zp.zones_per_chunk;
    za.cap = zp.cap;
    za.use_cursors = zp.use_cursors;
    za.pw = std::min<u32>(4, (pl.prod_bits + 31) / 32);
    za.bm_words = zp.bm_words;
    za.cursors = cursors.as<u32>();
    za.cursor_stride = cursor_stride;
    za.off_acc = zp.off_acc;
    za.off_bm = zp.off_bm;
    za.off_pre = zp.off_pre;
    za.off_list = zp.off_list;
    za.okey = g.p->key;
    za.oplanes = g.p->planes;
    za.ostride = g.p->stride;
    za.osign = g.p->sign;
    za.out_limbs = pl.out_limbs;
    za.out_cap = out_cap;
    zone_fill_codec(pl, za.cd);
    za.chunk_counter = chunk_counter;
    za.zone_count = zone_count;
    za.zone_off = zone_off;
    za.bump = bump;
    za.ctr = ctx->d_ctr;
    const char *trace_path = getenv("PK_ZONE_TRACE");
    const size_t max_chunks = (nz + zp.zones_per_chunk - 1) / zp.zones_per_chunk;
    DevBuf trace(ctx, trace_path ? max_chunks * 32 : 0);
    if (trace_path) CUDA_CHECK(cudaMemsetAsync(trace.p, 0, max_chunks * 32, ctx->stream));
    za.trace = trace.as<u64>();
    const bool is_signed = A->any_neg || B->any_neg;
    const u32 grid = zp.grid;
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
    switch (zp.NW) {
    case 2: zone_launch_nw<2>(ctx, za, grid, zp.smem, zp.mode, is_signed, zp.tpb); break;
    case 3: zone_launch_nw<3>(ctx, za, grid, zp.smem, zp.mode, is_signed, zp.tpb); break;
    case 4: zone_launch_nw<4>(ctx, za, grid, zp.smem, zp.mode, is_signed, zp.tpb); break;
    case 5: zone_launch_nw<5>(ctx, za, grid, zp.smem, zp.mode, is_signed, zp.tpb); break;
    default: zone_launch_nw<6>(ctx, za, grid, zp.smem, zp.mode, is_signed, zp.tpb); break;
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
    PK_LAUNCH(ctx, k_zone_scan, 1, 1024, 0, zone_count, zone_prefix, n_zones_dev, &ctx->d_ctr->n_out);
    MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
    readback(ctx, hc, ctx->d_ctr, sizeof(MulCounters));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (trace_path) {
        std::vector<u64> tr(max_chunks * 4);
        CUDA_CHECK(cudaMemcpy(tr.data(), trace.p, tr.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trace_path, "wb")) {
            fwrite(tr.data(), 8, tr.size(), f);
            fclose(f);
        }
    }
    const u64 n_out = hc->n_out;
    if (n_out > out_cap || (hc->flags & kFlagTableFull)) fail(PK_E_CUDA, "zone path: result count exceeds its bound (internal error)");
    ctx->stats.table_slots = (u64)zp.cap;
    ctx->stats.acc_lanes = zp.NW;
    ctx->stats.chunk_bits = 0;
    ctx->stats.retries = zp.mode | (zp.tpb << 8); // bit 0: bitmap zones; bits 8..: threads per block (spare counter)
    DPolyGuard e{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
    if (n_out) {
        PK_LAUNCH(ctx, k_zone_gather, grid_for(n_out, kGatherPerBlock), 256, 0, zone_count, zone_off, zone_prefix, n_zones_dev,
                  g.p->key, g.p->planes, g.p->stride, g.p->sign, e.p->key, e.p->planes, e.p->stride, e.p->sign, pl.out_limbs, n_out);
    }
    return e.release();
}

} // namespace pk
