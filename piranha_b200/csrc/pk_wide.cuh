// Wide coefficients: operands of 3 or 4 limbs (the pairing kernels take one or two).
//
// The reference never overflows: mppp::integer keeps N limbs in place and promotes beyond (integer.hpp:67-85), and its
// *_dynamic benchmarks multiply such operands (benchmarks/fateman1_dynamic.cpp).  Here the width protocol of the C ABI
// is: operands of up to two limbs and results of up to five run on the pairing kernels directly; operands of up to FOUR
// limbs are split in 128-bit halves  a = a0 + 2^128 a1,  b = b0 + 2^128 b1  (views of the same term arrays, no copy), the
// up to four half products  a_i * b_j  run through the same kernels (same keys, so the same output partition), and
// their terms are added up with the shifts 0 / 128 / 128 / 256 bits by a sort-and-reduce pass on the device (results of
// up to nine limbs).  Anything wider returns PK_E_CF_WIDTH: it is reported, never wrapped and never computed on the CPU.
//
// Included by pk_abi.cu (uses mul_impl, ensure_meta, radix_sort_pairs).
#pragma once

namespace pk
{

constexpr u32 kWideMaxParts = 4;
constexpr u32 kWideLimbs = 10; // accumulator of the reduce pass: 9 result limbs + sign headroom

struct WideSrc {
    const i64 *key[kWideMaxParts];
    const u64 *planes[kWideMaxParts];
    const signed char *sign[kWideMaxParts];
    u64 stride[kWideMaxParts];
    u64 off[kWideMaxParts + 1]; // first position of part s in the concatenation
    u32 limbs[kWideMaxParts];
    u32 shift[kWideMaxParts]; // in limbs
    u32 n_parts;
    i64 h_min;
};

// concatenate the keys of all parts (biased to unsigned so that the radix sort orders them like the signed codes)
__global__ void k_wide_collect(const WideSrc w, u64 *__restrict__ okey, u32 *__restrict__ osrc)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.off[w.n_parts]) return;
    u32 s = 0;
    while (s + 1 < w.n_parts && i >= w.off[s + 1]) ++s;
    const u64 t = i - w.off[s];
    okey[i] = (u64)(w.key[s][t] - w.h_min);
    osrc[i] = (s << 30) | (u32)t;
}

__global__ void k_wide_heads(const u64 *__restrict__ key, u64 n, u32 *__restrict__ head)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// one thread per distinct key: add the (at most four) shifted terms, write sign / magnitude and whether it is non-zero
__global__ void k_wide_reduce(const WideSrc w, const u64 *__restrict__ key, const u32 *__restrict__ src, const u32 *__restrict__ head,
                              const u32 *__restrict__ upos, u64 n, u64 *__restrict__ tkey, u64 *__restrict__ tmag /* [unique][kWideLimbs] */,
                              signed char *__restrict__ tsign, u32 *__restrict__ tnz, u32 *__restrict__ max_bits)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    u64 acc[kWideLimbs];
#pragma unroll
    for (u32 l = 0; l < kWideLimbs; ++l) acc[l] = 0;
    for (u64 j = i; j < n && (j == i || !head[j]); ++j) {
        const u32 s = src[j] >> 30, t = src[j] & 0x3FFFFFFFu;
        const bool neg = w.sign[s][t] < 0;
        u64 carry = 0; // carry (add) or borrow (subtract) into the next limb
        for (u32 l = w.shift[s]; l < kWideLimbs; ++l) {
            const u32 q = l - w.shift[s];
            const u64 m = q < w.limbs[s] ? w.planes[s][(u64)q * w.stride[s] + t] : 0ull;
            if (!neg) {
                const u64 x = acc[l] + m, c1 = x < m;
                const u64 y = x + carry, c2 = y < carry;
                acc[l] = y;
                carry = c1 | c2;
            } else {
                const u64 x = acc[l] - m, b1 = acc[l] < m;
                const u64 y = x - carry, b2 = x < carry;
                acc[l] = y;
                carry = b1 | b2;
            }
        }
    }
    const bool neg = (acc[kWideLimbs - 1] >> 63) != 0;
    if (neg) {
        u64 c = 1;
        for (u32 l = 0; l < kWideLimbs; ++l) {
            const u64 x = ~acc[l] + c;
            c = (c && x == 0) ? 1 : 0;
            acc[l] = x;
        }
    }
    u32 bits = 0;
    for (u32 l = 0; l < kWideLimbs; ++l)
        if (acc[l]) bits = 64u * l + bitlen64(acc[l]);
    const u32 u = upos[i];
    tkey[u] = key[i];
    for (u32 l = 0; l < kWideLimbs; ++l) tmag[(u64)u * kWideLimbs + l] = acc[l];
    tsign[u] = neg ? (signed char)-1 : (signed char)1;
    tnz[u] = bits ? 1u : 0u;
    if (bits) atomicMax(max_bits, bits);
}

__global__ void k_wide_compact(const u64 *__restrict__ tkey, const u64 *__restrict__ tmag, const signed char *__restrict__ tsign,
                               const u32 *__restrict__ tnz, const u32 *__restrict__ pos, u64 n_unique, i64 h_min, i64 *__restrict__ okey,
                               u64 *__restrict__ oplanes, u64 ostride, signed char *__restrict__ osign, u32 out_limbs)
{
    const u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique || !tnz[u]) return;
    const u32 p = pos[u];
    okey[p] = (i64)tkey[u] + h_min;
    for (u32 l = 0; l < out_limbs; ++l) oplanes[(u64)l * ostride + p] = tmag[u * kWideLimbs + l];
    osign[p] = tsign[u];
}

} // namespace pk
