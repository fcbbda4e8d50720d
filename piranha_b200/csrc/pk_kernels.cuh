// Kernels of the global-memory (HBM / L2) accumulation paths and of the operand / result plumbing.
//
//   k_meta            per-variable exponent min/max/gcd + coefficient width  (check_bounds operand scan,
//                                                                            polynomial.hpp:1142-1156)
//   k_prepare         Kronecker code -> tight code | sign, degree            (unpack + key_degree_impl,
//                                                                            kronecker_monomial.hpp:1083-1119)
//   k_skip_limits     sl[i] = upper_bound(deg2, max_degree - deg1[i])        (_get_skip_limits, polynomial.hpp:1495-1509)
//   k_mul_global      pairing + multiply-accumulate, dense or hash           (task_consume, polynomial.hpp:1723-1759;
//                                                                            blocked_multiplication, base_series_multiplier.hpp:449-504)
//   k_dense_* / k_hash_*  carry normalisation, zero removal, compaction      (sanitise_series, base_series_multiplier.hpp:855-949)
#pragma once

#include "pk_device.cuh"

namespace pk
{

// ------------------------------------------------------------------------------------------------ plumbing

// limb planes -> AoS rows of out_limbs >= limbs words (missing high limbs are zero)
__global__ void k_planes_to_aos(const u64 *__restrict__ planes, u64 *__restrict__ aos, u64 n, u32 limbs, u64 stride,
                                u32 out_limbs)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (u32 l = 0; l < out_limbs; ++l) aos[i * out_limbs + l] = l < limbs ? planes[l * stride + i] : 0ull;
}

// pk_download_grouped: group number of every term, then the terms gathered in the sorted order (magnitudes row-major, as
// the host result holds them)
__global__ void k_group_of(const i64 *__restrict__ key, u64 n, u32 shift, u32 mask, u32 *__restrict__ group, u32 *__restrict__ idx)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    group[i] = (u32)((u64)key[i] >> shift) & mask;
    idx[i] = (u32)i;
}
// first position of every group in the sorted group numbers (lower bound), and n for the entry past the last group
__global__ void k_group_offsets(const u32 *__restrict__ sorted_group, u64 n, u32 n_groups, u64 *__restrict__ off)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (sorted_group[mid] < g) lo = mid + 1;
        else hi = mid;
    }
    off[g] = lo;
}
__global__ void k_group_gather(const u32 *__restrict__ perm, const i64 *__restrict__ key, const u64 *__restrict__ planes,
                               const signed char *__restrict__ sign, u64 n, u32 limbs, u64 stride, i64 *__restrict__ okey,
                               u64 *__restrict__ oaos, signed char *__restrict__ osign)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 s = perm[i];
    okey[i] = key[s];
    for (u32 l = 0; l < limbs; ++l) oaos[i * limbs + l] = planes[l * stride + s];
    osign[i] = sign[s];
}

// ---- multi-GPU term exchange over peer-mapped memory (NVLink / NVSwitch): every rank of a partitioned product pushes
// its terms straight into the result arrays of ALL ranks with plain stores.
constexpr u32 kMaxPeers = 16;
struct PeerTable {
    u64 *slot[kMaxPeers];          // where this rank's {n, limbs} goes on peer j
    i64 *key[kMaxPeers];           // result arrays of peer j
    u64 *mag[kMaxPeers];
    signed char *sign[kMaxPeers];
};

// publish {n, limbs} of this rank's part in every peer's size table
__global__ void k_peer_publish(const PeerTable t, u32 n_peers, u64 n, u64 limbs)
{
    const u32 j = threadIdx.x;
    if (j < n_peers) {
        t.slot[j][0] = n;
        t.slot[j][1] = limbs;
    }
}

// sizes: {n, limbs} of every rank (local copy of the size table, complete after the barrier that follows
// k_peer_publish).  Rank `rank` writes its terms at term offset sum(n of lower ranks), magnitude rows of
// max(limbs of all ranks) words.  A result that does not fit (cap_terms, cap_limbs) is not written at all: every rank
// takes the same decision from the same table and the host grows the buffers.
__global__ void __launch_bounds__(256) k_peer_push(const PeerTable t, u32 n_peers, u32 rank, const u64 *__restrict__ sizes,
                                                  const i64 *__restrict__ key, const u64 *__restrict__ planes, u64 stride,
                                                  const signed char *__restrict__ sign, u64 n, u32 limbs, u64 cap_terms,
                                                  u32 cap_limbs)
{
    u64 off = 0, total = 0;
    u32 out_limbs = 1;
    for (u32 r = 0; r < n_peers; ++r) {
        const u64 nr = __ldg(&sizes[2 * r]);
        if (r < rank) off += nr;
        total += nr;
        if (nr) out_limbs = max(out_limbs, (u32)__ldg(&sizes[2 * r + 1]));
    }
    if (total > cap_terms || out_limbs > cap_limbs) return;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const i64 k = key[i];
    const signed char sg = sign[i];
    u64 m[8];
#pragma unroll
    for (u32 l = 0; l < 8; ++l) m[l] = (l < limbs) ? planes[l * stride + i] : 0ull;
    const u64 o = off + i;
    for (u32 j = 0; j < n_peers; ++j) {
        const u32 p = (rank + j) % n_peers; // start with the own copy, spread the peers over the links
        t.key[p][o] = k;
        t.sign[p][o] = sg;
        u64 *row = t.mag[p] + o * out_limbs;
#pragma unroll
        for (u32 l = 0; l < 8; ++l)
            if (l < out_limbs) row[l] = m[l];
    }
}

__global__ void k_iota32(u32 *out, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)i;
}

// gather terms through a permutation (after sorting by key)
__global__ void k_gather_terms(const u32 *__restrict__ perm, const i64 *__restrict__ raw_key,
                               const u64 *__restrict__ mag_aos, const signed char *__restrict__ sign, i64 *__restrict__ okey,
                               u64 *__restrict__ oplanes, signed char *__restrict__ osign, u64 n, u32 limbs, u64 stride, u32 *flags)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 s = perm[i];
    okey[i] = raw_key[s];
    // the terms are in key order now: equal neighbours are duplicate keys (a series holds every key once, series.hpp:1312-1320;
    // the width proof "at most min(n1, n2) products per code" relies on it)
    if (i && raw_key[perm[i - 1]] == raw_key[s]) atomicOr(flags, kMetaDuplicate);
    for (u32 l = 0; l < limbs; ++l) oplanes[l * stride + i] = mag_aos[(u64)s * limbs + l];
    osign[i] = sign[s];
}

__device__ __forceinline__ void atomic_gcd(u64 *p, u64 v)
{
    if (!v) return;
    u64 old = *((volatile u64 *)p);
    for (;;) {
        const u64 g = gcd_u64(old, v);
        if (g == old) return;
        const u64 prev = atomicCAS(p, old, g);
        if (prev == old) return;
        old = prev;
    }
}

// One pass over a polynomial: decode every key, reduce per-variable min / max / gcd(e - e_first) and the widest
// coefficient.  meta must be initialised (emin=+inf, emax=-inf, egcd=0, max_bits=0).
__global__ void k_meta(const i64 *__restrict__ key, const u64 *__restrict__ planes, const signed char *__restrict__ sign,
                       u64 n, u32 limbs, u64 stride, CodecParams cp, MetaDev *meta, u32 aos)
{
    __shared__ i64 s_min[PK_MAX_VARS], s_max[PK_MAX_VARS];
    __shared__ u64 s_gcd[PK_MAX_VARS];
    __shared__ u32 s_bits;
    for (u32 v = threadIdx.x; v < cp.n_vars; v += blockDim.x) {
        s_min[v] = INT64_MAX;
        s_max[v] = INT64_MIN;
        s_gcd[v] = 0;
    }
    if (threadIdx.x == 0) s_bits = 0;
    __syncthreads();
    const i64 first = key[0] - cp.h_min;
    volatile i64 *v_min = s_min, *v_max = s_max;
    volatile u64 *v_gcd = s_gcd;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        i64 code = key[i] - cp.h_min, fcode = first;
        for (u32 v = 0; v < cp.n_vars; ++v) {
            const i64 r = cp.radix[v];
            const i64 e = code % r - cp.M[v], e0 = fcode % r - cp.M[v];
            code /= r;
            fcode /= r;
            // plain reads first: the shared values only ever improve, so a stale read can only cause a
            // redundant atomic, never a missed update
            if (e < v_min[v]) atomicMin(&s_min[v], e);
            if (e > v_max[v]) atomicMax(&s_max[v], e);
            const u64 d = (u64)(e > e0 ? e - e0 : e0 - e);
            const u64 g = v_gcd[v];
            if (d && g != 1 && (g == 0 || d % g != 0)) atomic_gcd(&s_gcd[v], d);
        }
        u32 bits = 0;
        for (u32 l = 0; l < limbs; ++l) {
            const u64 m = aos ? planes[i * limbs + l] : planes[l * stride + i]; // (aos: the caller's row-major limbs)
            if (m) bits = 64u * l + bitlen64(m);
        }
        if (bits > s_bits) atomicMax(&s_bits, bits);
        // bit 0: some coefficient is negative; bit 1: some coefficient is zero (sign 0 or zero magnitude)
        u32 flag = (bits == 0u || sign[i] == 0) ? kMetaZero : (sign[i] < 0 ? kMetaNeg : 0u);
        if (key[i] < cp.h_min || key[i] > cp.h_max) flag |= kMetaKeyRange; // (the reference's decode throws for such a code)
        if (sign[i] < -1 || sign[i] > 1) flag |= kMetaBadSign;
        if (flag && (*((volatile u32 *)&meta->any_neg) & flag) != flag) atomicOr(&meta->any_neg, flag);
    }
    __syncthreads();
    for (u32 v = threadIdx.x; v < cp.n_vars; v += blockDim.x) {
        atomicMin(&meta->emin[v], s_min[v]);
        atomicMax(&meta->emax[v], s_max[v]);
        atomic_gcd(&meta->egcd[v], s_gcd[v]);
    }
    if (threadIdx.x == 0) atomicMax(&meta->max_bits, s_bits);
}

// Per-multiplication operand preparation: tight code | sign bit, degree (total or partial), coefficient limbs
// are read straight from the operand's planes by the multiply kernel.
struct PrepareArgs {
    const i64 *key;
    const signed char *sign;
    u64 n;
    u64 *okey;
    i64 *odeg;
    const u64 *m0;
    uint4 *opack;
};
__device__ __forceinline__ void prepare_term(const PrepareArgs &p, const CodecParams &cp, u64 i)
{
    if (i >= p.n) return;
    i64 code = p.key[i] - cp.h_min;
    u64 tight = 0;
    i64 deg = 0;
    for (u32 v = 0; v < cp.n_vars; ++v) {
        const i64 r = cp.radix[v];
        const i64 e = code % r - cp.M[v];
        code /= r;
        tight += (u64)((e - cp.emin[v]) / cp.step[v]) * cp.stride[v];
        if (cp.mask[v]) deg += e;
    }
    const signed char s = p.sign[i];
    // a zero coefficient contributes nothing: its magnitude limbs are zero, the sign bit is irrelevant
    p.okey[i] = tight | (s < 0 ? kSignBit : 0ull);
    if (p.odeg) p.odeg[i] = deg;
    if (p.opack) { // packed 16-byte record of the shared-memory paths: {tight code (32 bit), sign, magnitude (one limb)}
        const u64 m = p.m0[i];
        p.opack[i] = make_uint4((u32)tight, s < 0 ? 1u : 0u, (u32)m, (u32)(m >> 32));
    }
}
// both operands in one launch: blocks [0, blocksA) take the terms of a, the rest those of b
__global__ void k_prepare(const PrepareArgs a, const PrepareArgs b, const CodecParams cpa, const CodecParams cpb, u32 blocksA)
{
    if (blockIdx.x < blocksA) prepare_term(a, cpa, (u64)blockIdx.x * blockDim.x + threadIdx.x);
    else prepare_term(b, cpb, (u64)(blockIdx.x - blocksA) * blockDim.x + threadIdx.x);
}

// permute prepared operand arrays (after sorting the second operand by degree, polynomial.hpp:1427-1442)
__global__ void k_permute_operand(const u32 *__restrict__ perm, const u64 *__restrict__ ikey, const u64 *__restrict__ planes,
                                  u64 *__restrict__ okey, u64 *__restrict__ oplanes, u64 n, u32 limbs, u64 stride,
                                  u64 ostride)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 s = perm[i];
    okey[i] = ikey[s];
    for (u32 l = 0; l < limbs; ++l) oplanes[l * ostride + i] = planes[l * stride + s];
}

__global__ void k_flip_deg(const i64 *__restrict__ in, u64 *__restrict__ out, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u64)in[i] ^ kSignBit;
}

// Skip limits: sl[i] = number of terms of the (degree-sorted) second operand with deg2 <= max_degree - deg1[i].
// The subtraction is checked like safe_int_sub (polynomial.hpp:1248-1252): overflow raises the degree flag.
__global__ void k_skip_limits(const i64 *__restrict__ deg1, u64 n1, const u64 *__restrict__ deg2_sorted_flipped, u64 n2,
                              i64 max_degree, u32 *__restrict__ sl, MulCounters *ctr)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) {
        const i64 d1 = deg1[i];
        const i64 c = (i64)((u64)max_degree - (u64)d1);
        const bool ovf = ((max_degree ^ d1) & (max_degree ^ c)) < 0;
        if (ovf) {
            atomicOr(&ctr->flags, kFlagDegreeOverflow);
            sl[i] = 0;
        } else {
            const u64 cf = (u64)c ^ kSignBit;
            u64 lo = 0, hi = n2;
            while (lo < hi) {
                const u64 mid = (lo + hi) >> 1;
                if (cf < deg2_sorted_flipped[mid]) hi = mid;
                else lo = mid + 1;
            }
            sl[i] = (u32)lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------ multiply

struct MulArgs {
    const u64 *keyA; // prepared keys (tight | sign)
    const u64 *magA; // planes
    u64 strideA;
    u64 nA;
    const u64 *keyB;
    const u64 *magB;
    u64 strideB;
    u64 nB;
    const u32 *sl; // skip limits per A row or nullptr
    u64 lo, hi;    // output partition [lo, hi) in tight-code space (FILTER)
    u64 *acc;      // dense: lanes[code * nl + l]; hash: records of rec_words u64 {key, lanes...}
    u64 cap_mask;  // hash capacity - 1
    u32 hash_shift; // 64 - log2(capacity)
    u32 rec_words;
    u32 nl;
    u32 cbits;
    u32 keys_sorted; // both operands sorted by tight code (enables range tile skipping)
    MulCounters *ctr;
};

constexpr int kMulThreads = 256;
constexpr int kTileJ = 512;

__device__ __forceinline__ u64 hash_slot(u64 *tab, u64 key, const MulArgs &a)
{
    u64 h = (key * 0x9E3779B97F4A7C15ull) >> a.hash_shift;
    for (;;) {
        u64 *rec = tab + h * a.rec_words;
        u64 cur = *((volatile u64 *)rec);
        if (cur == key) return h;
        if (cur == kEmptyKey) {
            cur = atomicCAS(rec, kEmptyKey, key);
            if (cur == kEmptyKey || cur == key) return h;
        }
        h = (h + 1) & a.cap_mask;
    }
}

// One block = TI rows of A (R per warp, held in registers) x one tile of kTileJ terms of B staged in shared memory
// by TMA bulk copies.  Within one warp instruction the 32 lanes hold 32 different B terms against the same A term,
// so their 32 destination codes are distinct by construction (keys of a series are unique): no intra-warp conflicts
// and no need for match/aggregate before the atomics.
template <int MODE /*0 dense, 1 hash*/, int LA, int LB, bool FILTER>
__global__ void __launch_bounds__(kMulThreads) k_mul_global(MulArgs a)
{
    constexpr int R = (LA + LB == 2) ? 8 : 4;
    constexpr int TI = (kMulThreads / 32) * R;
    constexpr int LP = LA + LB;
    __shared__ __align__(16) u64 s_key[kTileJ];
    __shared__ __align__(16) u64 s_mag[LB][kTileJ];
    __shared__ __align__(8) u64 s_bar;

    const u32 tiles_j = (u32)((a.nB + kTileJ - 1) / kTileJ);
    const u32 bj = blockIdx.x % tiles_j, bi = blockIdx.x / tiles_j;
    const u64 j0 = (u64)bj * kTileJ;
    const u32 cnt_j = (u32)min((u64)kTileJ, a.nB - j0);
    const u64 i_blk = (u64)bi * TI;
    const u64 i_end = min(a.nA, i_blk + TI);

    if (FILTER) {
        bool skip = false;
        if (a.sl) { // rows are not degree-sorted: take the block maximum
            u32 m = 0;
            for (u64 i = i_blk + threadIdx.x; i < i_end; i += kMulThreads) m = max(m, a.sl[i]);
            skip = __syncthreads_or(m > j0) == 0;
        } else if (a.keys_sorted) {
            const u64 smin = (a.keyA[i_blk] & kKeyMask) + (a.keyB[j0] & kKeyMask);
            const u64 smax = (a.keyA[i_end - 1] & kKeyMask) + (a.keyB[j0 + cnt_j - 1] & kKeyMask);
            skip = smax < a.lo || smin >= a.hi;
        }
        if (skip) return;
    }

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        const u32 bytes = ((cnt_j + 1u) & ~1u) * 8u; // operand arrays are padded, 16-byte granularity
        mbar_expect_tx(&s_bar, bytes * (1 + LB));
        tma_load_1d(s_key, a.keyB + j0, bytes, &s_bar);
#pragma unroll
        for (int l = 0; l < LB; ++l) tma_load_1d(&s_mag[l][0], a.magB + l * a.strideB + j0, bytes, &s_bar);
    }
    __syncthreads();

    // A rows of this warp, in registers
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    u64 ka[R], ma[R][LA];
    u32 lim[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const u64 i = i_blk + warp * R + r;
        const bool ok = i < a.nA;
        ka[r] = ok ? a.keyA[i] : 0;
#pragma unroll
        for (int l = 0; l < LA; ++l) ma[r][l] = ok ? a.magA[l * a.strideA + i] : 0;
        u32 lm = ok ? cnt_j : 0;
        if (FILTER && a.sl && ok) {
            const u32 s = a.sl[i];
            lm = s > j0 ? (u32)min((u64)cnt_j, (u64)s - j0) : 0;
        }
        lim[r] = lm;
    }
    mbar_wait(&s_bar, 0);

    u32 pairs = 0;
    for (u32 j = lane; j < cnt_j; j += 32) {
        const u64 kb = s_key[j];
        u64 mb[LB];
#pragma unroll
        for (int l = 0; l < LB; ++l) mb[l] = s_mag[l][j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (j < lim[r]) {
                const u64 code = (ka[r] & kKeyMask) + (kb & kKeyMask);
                if (FILTER && (code < a.lo || code >= a.hi)) continue;
                const bool neg = ((ka[r] ^ kb) >> 63) != 0;
                u64 p[LP];
                mag_mul<LA, LB>(ma[r], mb, p);
                u64 *lanes;
                if (MODE == 0) {
                    lanes = a.acc + code * a.nl;
                } else {
                    const u64 slot = hash_slot(a.acc, code, a);
                    lanes = a.acc + slot * a.rec_words + 1;
                }
                accumulate_lanes<LP>(lanes, p, a.nl, a.cbits, neg);
                ++pairs;
            }
        }
    }
    if (FILTER) {
        for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        if (lane == 0 && pairs) atomicAdd(&a.ctr->pairs, (u64)pairs);
    }
}

// hash table records {key, lanes...}: key = empty sentinel, lanes = 0 (one pass over the table)
__global__ void k_hash_init(u64 *__restrict__ tab, u64 words, u32 rec_words)
{
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < words) tab[w] = (w % rec_words == 0) ? kEmptyKey : 0ull;
}

// ------------------------------------------------------------------------------------------------ finalisation

struct FinalArgs {
    const u64 *acc;
    u64 n_slots; // dense: number of codes in [code0, code0 + n_slots); hash: capacity
    u64 code0;   // dense: first code of the scanned range
    u32 rec_words;
    u32 nl;
    u32 cbits;
    u32 out_limbs;
    CodecParams cp; // result codec (tight -> piranha)
    i64 base_code;
    u32 *block_counts;
    u64 *block_offsets;
    i64 *okey;
    u64 *oplanes;
    u64 ostride;
    signed char *osign;
    MulCounters *ctr;
};

constexpr int kFinThreads = 256;
constexpr int kFinItems = 8; // consecutive slots per thread

// Dense pass 1: count non-zero normalised slots per block.
__global__ void __launch_bounds__(kFinThreads) k_dense_count(FinalArgs f)
{
    __shared__ u32 s_ws[40];
    const u64 base = ((u64)blockIdx.x * kFinThreads + threadIdx.x) * kFinItems;
    u32 cnt = 0;
#pragma unroll
    for (int it = 0; it < kFinItems; ++it) {
        const u64 s = base + it;
        if (s < f.n_slots) {
            const u64 *lanes = f.acc + (f.code0 + s) * f.nl;
            u64 any = 0;
            for (u32 l = 0; l < f.nl; ++l) any |= lanes[l];
            if (any) {
                u64 mag[kMaxAccLimbs];
                if (normalise_lanes(lanes, f.nl, f.cbits, mag) != 0) ++cnt;
            }
        }
    }
    u32 total;
    block_excl_scan(cnt, s_ws, total);
    if (threadIdx.x == 0) f.block_counts[blockIdx.x] = total;
}

// Exclusive scan of the block counts (single block) -> 64-bit offsets, total into ctr->n_out.
__global__ void __launch_bounds__(1024) k_scan_counts(const u32 *__restrict__ counts, u64 *__restrict__ offsets, u64 n,
                                                      MulCounters *ctr)
{
    __shared__ u32 s_ws[40];
    __shared__ u64 s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += 1024) {
        const u64 i = base + threadIdx.x;
        const u32 v = i < n ? counts[i] : 0;
        u32 total;
        const u32 ex = block_excl_scan(v, s_ws, total);
        if (i < n) offsets[i] = s_run + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_run += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) ctr->n_out = s_run;
}

__device__ __forceinline__ void write_term(const FinalArgs &f, u64 pos, u64 code, int sgn, const u64 (&mag)[kMaxAccLimbs],
                                           u32 &maxbits, u32 &flags)
{
    f.okey[pos] = tight_to_piranha(code, f.cp, f.base_code);
#pragma unroll
    for (int l = 0; l < kMaxAccLimbs; ++l) {
        if ((u32)l < f.out_limbs) f.oplanes[(u64)l * f.ostride + pos] = mag[l];
        else if (mag[l]) flags |= kFlagWidth; // cannot happen: out_limbs comes from a proven bound
    }
    f.osign[pos] = (signed char)sgn;
    maxbits = max(maxbits, mag_bits(mag));
}

// Dense pass 2: ordered compaction (codes ascend, so the output is sorted by key).
__global__ void __launch_bounds__(kFinThreads) k_dense_write(FinalArgs f)
{
    __shared__ u32 s_ws[40];
    const u64 base = ((u64)blockIdx.x * kFinThreads + threadIdx.x) * kFinItems;
    u64 mags[kFinItems][kMaxAccLimbs];
    int sg[kFinItems];
    u32 cnt = 0;
#pragma unroll
    for (int it = 0; it < kFinItems; ++it) {
        const u64 s = base + it;
        sg[it] = 0;
        if (s < f.n_slots) {
            const u64 *lanes = f.acc + (f.code0 + s) * f.nl;
            u64 any = 0;
            for (u32 l = 0; l < f.nl; ++l) any |= lanes[l];
            if (any) sg[it] = normalise_lanes(lanes, f.nl, f.cbits, mags[it]);
        }
        cnt += sg[it] != 0;
    }
    u32 total;
    const u32 ex = block_excl_scan(cnt, s_ws, total);
    u64 pos = f.block_offsets[blockIdx.x] + ex;
    u32 maxbits = 0, flags = 0;
#pragma unroll
    for (int it = 0; it < kFinItems; ++it) {
        if (sg[it] != 0) {
            write_term(f, pos, f.code0 + base + it, sg[it], mags[it], maxbits, flags);
            ++pos;
        }
    }
    for (int o = 16; o > 0; o >>= 1) maxbits = max(maxbits, __shfl_xor_sync(0xffffffffu, maxbits, o));
    if (lane_id() == 0 && maxbits) atomicMax(&f.ctr->max_bits, maxbits);
    if (flags) atomicOr(&f.ctr->flags, flags);
}

// Hash pass 1: every occupied slot with a non-zero normalised value emits (tight code, slot) through a
// warp-aggregated cursor (one atomicAdd per warp: ballot + popc + shuffle).
__global__ void __launch_bounds__(256) k_hash_collect(FinalArgs f, u64 *__restrict__ out_code, u32 *__restrict__ out_slot)
{
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < ((f.n_slots + 31) & ~31ull);
         s += (u64)gridDim.x * blockDim.x) {
        bool keep = false;
        u64 code = 0;
        if (s < f.n_slots) {
            const u64 *rec = f.acc + s * f.rec_words;
            code = rec[0];
            if (code != kEmptyKey) {
                u64 mag[kMaxAccLimbs];
                keep = normalise_lanes(rec + 1, f.nl, f.cbits, mag) != 0;
            }
        }
        const u32 ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            u64 basepos = 0;
            if (lane_id() == 0) basepos = atomicAdd(&f.ctr->n_out, (u64)__popc(ballot));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (keep) {
                const u64 pos = basepos + __popc(ballot & ((1u << lane_id()) - 1u));
                out_code[pos] = code;
                out_slot[pos] = (u32)s;
            }
        }
    }
}

// Hash pass 2 (after sorting the (code, slot) pairs by code): normalise again and write the final terms in order.
__global__ void __launch_bounds__(256) k_hash_write(FinalArgs f, const u64 *__restrict__ codes, const u32 *__restrict__ slots,
                                                    u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 maxbits = 0, flags = 0;
    if (i < n) {
        const u64 *rec = f.acc + (u64)slots[i] * f.rec_words;
        u64 mag[kMaxAccLimbs];
        const int sg = normalise_lanes(rec + 1, f.nl, f.cbits, mag);
        write_term(f, i, codes[i], sg, mag, maxbits, flags);
    }
    for (int o = 16; o > 0; o >>= 1) maxbits = max(maxbits, __shfl_xor_sync(0xffffffffu, maxbits, o));
    if (lane_id() == 0 && maxbits) atomicMax(&f.ctr->max_bits, maxbits);
    if (flags) atomicOr(&f.ctr->flags, flags);
}

// ------------------------------------------------------------------------------------------------ partition cuts

// Stratified sample of pair sums (every (nA/SA)-th row against every (nB/SB)-th column); pairs excluded by the
// truncation get the sentinel so they sort last.
__global__ void k_sample_sums(const u64 *__restrict__ keyA, u64 nA, const u64 *__restrict__ keyB, u64 nB, u32 SA, u32 SB,
                              const i64 *__restrict__ degA, const i64 *__restrict__ degB, i64 max_degree, int trunc,
                              u64 *__restrict__ out, MulCounters *ctr)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (u64)SA * SB) return;
    const u64 ia = (u64)(t / SB) * nA / SA, jb = (u64)(t % SB) * nB / SB;
    u64 v = (keyA[ia] & kKeyMask) + (keyB[jb] & kKeyMask);
    if (trunc) {
        const __int128 d = (__int128)degA[ia] + (__int128)degB[jb];
        if (d > (__int128)max_degree) v = kEmptyKey;
    }
    if (v != kEmptyKey) atomicAdd(&ctr->aux, 1ull);
    out[t] = v;
}

// cuts[p] = the (valid * p / P)-th smallest sampled pair sum, p = 1 .. P-1 (valid = number of admissible samples)
__global__ void k_pick_cuts(const u64 *__restrict__ sorted, const MulCounters *__restrict__ ctr, u32 P, u64 *__restrict__ cuts)
{
    const u64 valid = ctr->aux;
    for (u32 p = threadIdx.x; p <= P; p += blockDim.x)
        cuts[p] = (p == 0 || p == P || valid == 0) ? 0ull : sorted[(u64)((unsigned __int128)valid * p / P)];
}

// ------------------------------------------------------------------------------------------------ digest

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// Order-independent digest: sum over terms of mix(key, sign, limbs) mod 2^64.
__global__ void k_digest(const i64 *__restrict__ key, const u64 *__restrict__ planes, const signed char *__restrict__ sign,
                         u64 n, u32 limbs, u64 stride, u64 *out)
{
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 h = mix64((u64)key[i] + 0x9E3779B97F4A7C15ull);
        h = mix64(h ^ (u64)(unsigned char)sign[i]);
        for (u32 l = 0; l < limbs; ++l) {
            const u64 m = planes[l * stride + i];
            if (m) h = mix64(h ^ (m + 0x632BE59BD9B4E019ull * (l + 1)));
        }
        acc += h;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane_id() == 0 && acc) atomicAdd(out, acc);
}

} // namespace pk
