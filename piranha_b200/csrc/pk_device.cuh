// Device-side building blocks shared by the kernels of the sparse polynomial multiplier (sm_100a).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/piranha_b200.h"

namespace pk
{

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

// Prepared operand keys: bit 63 carries the coefficient sign, bits 0..62 the re-encoded (tight) code.
constexpr u64 kSignBit = 1ull << 63;
constexpr u64 kKeyMask = ~kSignBit;
constexpr u64 kEmptyKey = ~0ull; // hash-table sentinel; real codes are < 2^62

constexpr int kMaxAccLimbs = 6; // widest normalised accumulator (64-bit limbs)

// Re-encoding of one operand's Kronecker codes into the tight mixed-radix code of this product.
// Decoding restates kronecker_array::decode (kronecker_array.hpp:353-364): code - h_min, then mod/div by 2*M+1.
struct CodecParams {
    u32 n_vars;
    u32 trunc_mode;
    i64 h_min, h_max;        // smallest / largest Kronecker code of the codification
    i64 radix[PK_MAX_VARS];  // piranha radix 2*M_k+1
    i64 M[PK_MAX_VARS];      // piranha limit M_k
    i64 emin[PK_MAX_VARS];   // operand (or result) minimum exponent
    i64 step[PK_MAX_VARS];   // common exponent step g_k >= 1 (gcd compaction)
    u64 stride[PK_MAX_VARS]; // tight stride of variable k
    u64 tradix[PK_MAX_VARS]; // tight radix of variable k
    i64 pstride[PK_MAX_VARS]; // piranha stride c_k = prod_{u<k} (2*M_u+1)
    unsigned char mask[PK_MAX_VARS];
};

// Per-polynomial metadata reduced on the device (the operand half of check_bounds, polynomial.hpp:1142-1156).
struct MetaDev {
    i64 emin[PK_MAX_VARS];
    i64 emax[PK_MAX_VARS];
    u64 egcd[PK_MAX_VARS]; // gcd over terms of (e - e_first)
    u32 max_bits;          // bit length of the largest coefficient magnitude
    u32 any_neg;           // bit 0: some coefficient is negative; 1: some coefficient is zero; 2: a key outside [h_min, h_max];
                           // 3: a sign outside {-1, 0, +1}; 4: two terms with the same key
};
constexpr u32 kMetaNeg = 1u, kMetaZero = 2u, kMetaKeyRange = 4u, kMetaBadSign = 8u, kMetaDuplicate = 16u;

// Device-side counters of one multiplication (read back once at the end).
struct MulCounters {
    u64 pairs;      // term products performed (after truncation / partition filtering)
    u64 n_out;      // result terms
    u32 max_bits;   // bit length of the largest result magnitude
    u32 flags;      // error bits
    u64 aux;        // scratch (compaction cursor)
};
constexpr u32 kFlagDegreeOverflow = 1u;
constexpr u32 kFlagTableFull = 2u;
constexpr u32 kFlagWidth = 4u;

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ u32 bitlen64(u64 x) { return 64u - (u32)__clzll((long long)x); }

// ---- 64-bit limb arithmetic ---------------------------------------------------------------------------------

// (la x lb)-limb schoolbook product of magnitudes; IMAD.WIDE.U32 chains after ptxas lowers __umul64hi.
template <int LA, int LB>
__device__ __forceinline__ void mag_mul(const u64 (&a)[LA], const u64 (&b)[LB], u64 (&p)[LA + LB])
{
#pragma unroll
    for (int i = 0; i < LA + LB; ++i) p[i] = 0;
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        u64 carry = 0;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const u64 lo = a[i] * b[j], hi = __umul64hi(a[i], b[j]);
            u64 s = p[i + j] + lo;
            u64 c1 = s < lo;
            u64 s2 = s + carry;
            c1 += s2 < carry;
            p[i + j] = s2;
            carry = hi + c1; // cannot overflow: hi <= 2^64-2
        }
        p[i + LB] = carry;
    }
}

// Shift an LP-limb magnitude right by s bits, 1 <= s <= 63.
template <int LP>
__device__ __forceinline__ void shr_limbs(u64 (&p)[LP], u32 s)
{
#pragma unroll
    for (int i = 0; i < LP - 1; ++i) p[i] = (p[i] >> s) | (p[i + 1] << (64u - s));
    p[LP - 1] >>= s;
}

// Deferred-carry accumulation: the product magnitude is cut into `nl` chunks of `cbits` bits; chunk l is added
// (or subtracted, for a negative product) to 64-bit lane l with one fire-and-forget RED.ADD.64.  Lanes are signed
// counters with 63 - cbits bits of headroom, so up to 2^(63-cbits) products can land on one slot before a lane
// could wrap; the host picks cbits from min(n1, n2).  This replaces mppp::addmul (integer.hpp:83).
template <int LP>
__device__ __forceinline__ void accumulate_lanes(u64 *lanes, u64 (&p)[LP], u32 nl, u32 cbits, bool neg)
{
    const u64 cmask = (1ull << cbits) - 1ull;
#pragma unroll
    for (int l = 0; l < 2 * LP; ++l) {
        if ((u32)l < nl) {
            u64 chunk = p[0] & cmask;
            shr_limbs<LP>(p, cbits);
            if (neg) chunk = 0ull - chunk;
            if (chunk) atomicAdd(lanes + l, chunk); // result unused -> REDG.E.ADD.64
        }
    }
}

// Carry-normalise `nl` signed lanes into a two's complement integer of kMaxAccLimbs limbs, then to sign+magnitude.
// Returns the sign (-1, 0, +1); mag[] receives the magnitude.
__device__ __forceinline__ int normalise_lanes(const u64 *lanes, u32 nl, u32 cbits, u64 (&mag)[kMaxAccLimbs])
{
    u64 acc[kMaxAccLimbs];
#pragma unroll
    for (int i = 0; i < kMaxAccLimbs; ++i) acc[i] = 0;
    for (u32 l = 0; l < nl; ++l) {
        const i64 v = (i64)lanes[l];
        const u32 sh = l * cbits, w = sh >> 6, b = sh & 63u;
        // sign-extended (v << b) occupies limbs w, w+1 and the extension above
        const u64 lo = (u64)v << b;
        const u64 hi = b ? (u64)(v >> (64u - b)) : (u64)(v >> 63);
        const u64 ext = (u64)(v >> 63);
        u64 carry = 0;
#pragma unroll
        for (int i = 0; i < kMaxAccLimbs; ++i) {
            if ((u32)i >= w) {
                const u64 add = ((u32)i == w) ? lo : (((u32)i == w + 1) ? hi : ext);
                const u64 s = acc[i] + add;
                const u64 c1 = s < add;
                const u64 s2 = s + carry;
                carry = c1 + (s2 < carry);
                acc[i] = s2;
            }
        }
    }
    const bool neg = (acc[kMaxAccLimbs - 1] >> 63) != 0;
    u64 nz = 0;
    if (neg) {
        u64 carry = 1;
#pragma unroll
        for (int i = 0; i < kMaxAccLimbs; ++i) {
            const u64 t = ~acc[i] + carry;
            carry = (carry && t == 0) ? 1 : 0;
            mag[i] = t;
            nz |= t;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kMaxAccLimbs; ++i) {
            mag[i] = acc[i];
            nz |= acc[i];
        }
    }
    return nz ? (neg ? -1 : 1) : 0;
}

__device__ __forceinline__ u32 mag_bits(const u64 (&mag)[kMaxAccLimbs])
{
    u32 bits = 0;
#pragma unroll
    for (int i = 0; i < kMaxAccLimbs; ++i)
        if (mag[i]) bits = 64u * i + bitlen64(mag[i]);
    return bits;
}

// Tight code -> piranha Kronecker code of the result term (linear in the digits, kronecker_array.hpp:298-306).
__device__ __forceinline__ i64 tight_to_piranha(u64 code, const CodecParams &cp, i64 base_code)
{
    i64 out = base_code;
    for (u32 v = 0; v < cp.n_vars; ++v) {
        const u64 d = code % cp.tradix[v];
        code /= cp.tradix[v];
        out += (i64)d * cp.step[v] * cp.pstride[v];
    }
    return out;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) -----------------------------------------------------

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Small read-backs (counters, cut points) are written to mapped pinned host memory by a kernel instead of going through
// a copy engine: a DMA read-back queues behind any bulk device-to-host copy in flight (the pipelined pk_mul keeps one
// running on a second stream), which stalled every such read-back for the length of that copy.
__global__ void k_readback(u32 *__restrict__ host_dst, const u32 *__restrict__ dev_src, u32 n_words)
{
    for (u32 i = threadIdx.x; i < n_words; i += blockDim.x) host_dst[i] = dev_src[i];
    __threadfence_system();
}

__global__ void k_readback2(u32 *__restrict__ h1, const u32 *__restrict__ d1, u32 n1, u32 *__restrict__ h2, const u32 *__restrict__ d2, u32 n2)
{
    for (u32 i = threadIdx.x; i < n1; i += blockDim.x) h1[i] = d1[i];
    for (u32 i = threadIdx.x; i < n2; i += blockDim.x) h2[i] = d2[i];
    __threadfence_system();
}

// ---- warp scans ---------------------------------------------------------------------------------------------

__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32)o) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread over a block of up to 1024 threads; returns the block total in `total`.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *warp_sums /* >= 33 words of smem */, u32 &total)
{
    const u32 incl = warp_incl_scan(v);
    const u32 w = threadIdx.x >> 5;
    if (lane_id() == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        const u32 nw = (blockDim.x + 31) >> 5;
        u32 s = lane_id() < nw ? warp_sums[lane_id()] : 0;
        const u32 si = warp_incl_scan(s);
        warp_sums[lane_id()] = si - s;
        if (lane_id() == 31) warp_sums[32] = si;
    }
    __syncthreads();
    total = warp_sums[32];
    const u32 r = warp_sums[w] + incl - v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ u64 gcd_u64(u64 a, u64 b)
{
    while (b) {
        const u64 t = a % b;
        a = b;
        b = t;
    }
    return a;
}

} // namespace pk
