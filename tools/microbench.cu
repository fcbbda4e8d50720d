// Microbenchmarks for the accumulation primitives the multiplier is built on (B200, sm_100a):
//   - shared-memory 32-bit ATOMS.ADD (with / without use of the returned value, random / lane-consecutive addresses)
//   - global RED.ADD.64 on L2-resident and HBM-resident tables (random / run-clustered addresses)
//   - CAS-claim + RED (open addressing probe)
//   - random 16-byte LDG gathers (operand fetch)
//   - the INT32 multiply pipe: IMAD (32 x 32 + 32), IMAD.WIDE (32 x 32 + 64 -> 64) and the 64 x 64 -> 128 coefficient
//     product the pairing kernels issue (4 IMAD.WIDE-class operations), as independent chains per thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
// Output: one line per experiment, ops/s over the whole chip.  Numbers feed DESIGN.md's roofline section.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ u32 lcg(u32 &s)
{
    s = s * 1664525u + 1013904223u;
    return s;
}
__device__ __forceinline__ u64 mix(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    return x;
}

// MODE 0: random word, result unused. 1: random word, carry chain on the returned value (3 dependent atomics).
// 2: lane-consecutive words (row-run pattern), result unused. 3: lane-consecutive, carry chain.
template <int MODE>
__global__ void k_smem_atomics(u32 *out, int iters, u32 words_mask)
{
    extern __shared__ u32 s[];
    for (u32 i = threadIdx.x; i <= words_mask; i += blockDim.x) s[i] = 0;
    __syncthreads();
    u32 seed = blockIdx.x * 9781u + threadIdx.x * 7u + 1u;
    const u32 lane = threadIdx.x & 31u;
    u32 sink = 0;
    for (int it = 0; it < iters; ++it) {
        u32 r = lcg(seed);
        u32 base;
        if (MODE >= 2) base = (__shfl_sync(0xffffffffu, r, 0) >> 8) + lane;
        else base = r >> 8;
        const u32 v = r | 1u;
        if (MODE == 0 || MODE == 2) {
            atomicAdd(&s[base & words_mask], v);
        } else {
            // three-word accumulator with eager carry: old + v overflow -> +1 into the next word
            const u32 a0 = (base * 1u) & words_mask;
            u32 old = atomicAdd(&s[a0], v);
            u32 c = (old + v) < v;
            const u32 a1 = (a0 + (words_mask + 1) / 4) & words_mask;
            old = atomicAdd(&s[a1], (v >> 3) + c);
            c = (old + (v >> 3) + c) < old;
            const u32 a2 = (a1 + (words_mask + 1) / 4) & words_mask;
            old = atomicAdd(&s[a2], (v >> 7) + c);
            sink += old;
        }
    }
    __syncthreads();
    if (sink == 0x12345678u) out[0] = s[threadIdx.x & words_mask];
    if (threadIdx.x == 0) out[blockIdx.x + 1] = s[0];
}

// MODE 0: RED.64 random slot, `lanes` consecutive words per slot.  1: runs of 8 consecutive slots per warp-octet.
template <int MODE>
__global__ void k_global_red(u64 *tab, u64 slot_mask, int lanes, int iters)
{
    u32 seed = blockIdx.x * 9781u + threadIdx.x * 7u + 1u;
    const u32 lane = threadIdx.x & 31u;
    for (int it = 0; it < iters; ++it) {
        u64 r = mix(((u64)lcg(seed) << 32) | lcg(seed));
        if (MODE == 1) r = __shfl_sync(0xffffffffu, r, lane & ~7u) + (lane & 7u);
        u64 *p = tab + (r & slot_mask) * lanes;
        for (int l = 0; l < lanes; ++l) atomicAdd(p + l, r | 1ull);
    }
}

// open addressing: probe key word (volatile load), CAS if empty, then `lanes` REDs in the same 32-byte record
__global__ void k_hash_probe(u64 *tab, u64 slot_mask, u64 key_space, int lanes, int iters)
{
    u32 seed = blockIdx.x * 9781u + threadIdx.x * 7u + 1u;
    for (int it = 0; it < iters; ++it) {
        const u64 key = mix(((u64)lcg(seed) << 32) | lcg(seed)) % key_space;
        u64 h = (key * 0x9E3779B97F4A7C15ull) & slot_mask;
        for (;;) {
            u64 *rec = tab + h * 4;
            u64 cur = *((volatile u64 *)rec);
            if (cur == key) break;
            if (cur == ~0ull) {
                cur = atomicCAS(rec, ~0ull, key);
                if (cur == ~0ull || cur == key) break;
            }
            h = (h + 1) & slot_mask;
        }
        for (int l = 0; l < lanes; ++l) atomicAdd(tab + h * 4 + 1 + l, key | 1ull);
    }
}

__global__ void k_gather16(const ulonglong2 *src, u64 mask, int iters, u64 *out)
{
    u32 seed = blockIdx.x * 9781u + threadIdx.x * 7u + 1u;
    u64 acc = 0;
    for (int it = 0; it < iters; ++it) {
        const u64 r = mix(((u64)lcg(seed) << 32) | lcg(seed));
        const ulonglong2 v = src[r & mask];
        acc += v.x ^ v.y;
    }
    if (acc == 0x1234567ull) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f(); // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

// 64-bit shared-memory atomics: MODE 0 fire-and-forget, MODE 1 two-word (128-bit) accumulator with the carry taken from the
// returned value -- the alternative to three / four 32-bit words per accumulator.
template <int MODE>
__global__ void k_smem_atomics64(u32 *out, int iters, u32 words_mask)
{
    extern __shared__ u64 s64[];
    for (u32 i = threadIdx.x; i <= words_mask; i += blockDim.x) s64[i] = 0;
    __syncthreads();
    u32 seed = blockIdx.x * 9781u + threadIdx.x * 7u + 1u;
    u64 sink = 0;
    for (int it = 0; it < iters; ++it) {
        const u32 r = lcg(seed);
        const u32 base = r >> 8;
        const u64 v = ((u64)r << 24) | 1ull;
        if (MODE == 0) {
            atomicAdd(&s64[base & words_mask], v);
        } else {
            const u64 old = atomicAdd(&s64[base & words_mask], v);
            const u64 carry = (old + v < v) ? 1ull : 0ull;
            atomicAdd(&s64[(base + 1u) & words_mask], (u64)(r & 0xFFFu) + carry);
            sink += old;
        }
    }
    if (sink == 0x1234567ull) out[0] = (u32)sink;
}

// MODE 0: IMAD (mad.lo.u32), MODE 1: IMAD.WIDE (mad.wide.u32), MODE 2: 64 x 64 -> 128 product (mul.lo.u64 + mul.hi.u64), with
// CHAINS independent dependency chains per thread so that the pipe, not the latency, is measured.
template <int MODE, int CHAINS>
__global__ void k_imad(u64 *out, int iters, u32 seed0)
{
    u32 a[CHAINS], b[CHAINS];
    u64 w[CHAINS], h[CHAINS];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        a[c] = seed0 * (2u * c + 3u) + t;
        b[c] = (seed0 ^ 0x9E3779B9u) + 7u * c + t * 5u;
        w[c] = ((u64)a[c] << 32) | b[c];
        h[c] = 0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(a[c]));
            } else if (MODE == 1) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c]));
            } else {
                u64 lo, hi;
                asm volatile("mul.lo.u64 %0, %2, %3;\n\tmul.hi.u64 %1, %2, %3;" : "=l"(lo), "=l"(hi) : "l"(w[c]), "l"(w[c] | 1ull));
                w[c] = lo ^ hi;
                h[c] += hi;
            }
        }
    }
    u64 sink = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) sink += a[c] + w[c] + h[c];
    if (sink == 0x1234567ull) out[0] = sink; // (keeps the chains alive)
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    u32 *d_out;
    CK(cudaMalloc(&d_out, (sms * 8 + 8) * sizeof(u32)));

    // ---- shared atomics
    {
        const int iters = 4096, threads = 1024;
        for (int words_log = 10; words_log <= 15; words_log += 5) {
            const u32 mask = (1u << words_log) - 1;
            const size_t smem = (size_t)(mask + 1) * 4;
            auto run = [&](int mode) {
                float ms = 0;
                auto launch = [&]() {
                    switch (mode) {
                    case 0: k_smem_atomics<0><<<sms, threads, smem>>>(d_out, iters, mask); break;
                    case 1: k_smem_atomics<1><<<sms, threads, smem>>>(d_out, iters, mask); break;
                    case 2: k_smem_atomics<2><<<sms, threads, smem>>>(d_out, iters, mask); break;
                    default: k_smem_atomics<3><<<sms, threads, smem>>>(d_out, iters, mask); break;
                    }
                };
                CK(cudaFuncSetAttribute(k_smem_atomics<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CK(cudaFuncSetAttribute(k_smem_atomics<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CK(cudaFuncSetAttribute(k_smem_atomics<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CK(cudaFuncSetAttribute(k_smem_atomics<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                ms = time_ms(launch);
                CK(cudaGetLastError());
                const double per_thread = (mode & 1) ? 3.0 : 1.0;
                const double ops = (double)sms * threads * iters * per_thread;
                printf("smem_atomics words=2^%d mode=%d (%s, %s): %.3f ms, %.1f Gatom/s chip, %.2f atom/clk/SM @%.0f MHz\n",
                       words_log, mode, (mode >= 2) ? "lane-consecutive" : "random",
                       (mode & 1) ? "3-word carry chain" : "fire-and-forget", ms, ops / ms * 1e-6,
                       ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate * 1e-3);
            };
            for (int mode = 0; mode < 4; ++mode) run(mode);
        }
    }

    // ---- global RED.64 / hash probe / gathers at several table sizes
    for (int mb_log = 5; mb_log <= 12; mb_log += 1) { // 32 MB .. 4 GB
        const size_t bytes = (size_t)1 << (20 + mb_log);
        u64 *tab;
        if (cudaMalloc(&tab, bytes) != cudaSuccess) {
            printf("skip %zu MB\n", bytes >> 20);
            break;
        }
        const int iters = 256, threads = 256, blocks = sms * 16;
        const double n = (double)blocks * threads * iters;
        for (int lanes = 1; lanes <= 3; ++lanes) {
            CK(cudaMemset(tab, 0, bytes));
            struct L {
                static void go(u64 *t, u64 m, int ln, int it, int b, int th, int mode)
                {
                    if (mode == 0) k_global_red<0><<<b, th>>>(t, m, ln, it);
                    else k_global_red<1><<<b, th>>>(t, m, ln, it);
                }
            };
            const u64 slots_l = bytes / (8 * lanes);
            u64 mask = 1;
            while (mask * 2 <= slots_l) mask *= 2;
            for (int mode = 0; mode < 2; ++mode) {
                float t = time_ms([&]() { L::go(tab, mask - 1, lanes, iters, blocks, threads, mode); }, 3);
                printf("global_red64 table=%zuMB lanes=%d %s: %.3f ms, %.2f Gprod/s, %.2f Gatom/s\n", bytes >> 20, lanes,
                       mode ? "runs-of-8" : "random", t, n / t * 1e-6, n * lanes / t * 1e-6);
            }
        }
        {
            const u64 slots = bytes / 32;
            for (int fill = 0; fill < 2; ++fill) {
                const u64 key_space = fill ? slots / 2 : slots / 16;
                CK(cudaMemset(tab, 0xFF, bytes));
                k_hash_probe<<<blocks, threads>>>(tab, slots - 1, key_space, 2, iters); // populate (lanes garbage: fine)
                CK(cudaDeviceSynchronize());
                float t = time_ms([&]() { k_hash_probe<<<blocks, threads>>>(tab, slots - 1, key_space, 2, iters); }, 3);
                printf("hash_probe table=%zuMB load=%.3f lanes=2: %.3f ms, %.2f Gprod/s\n", bytes >> 20,
                       (double)key_space / slots, t, n / t * 1e-6);
            }
        }
        {
            float t = time_ms([&]() { k_gather16<<<blocks, threads>>>((const ulonglong2 *)tab, bytes / 16 - 1, iters, tab); }, 3);
            printf("gather16 table=%zuMB: %.3f ms, %.2f Gload/s\n", bytes >> 20, t, n / t * 1e-6);
        }
        CK(cudaFree(tab));
    }
    // ---- 64-bit shared atomics
    {
        const int iters = 4096, threads = 1024;
        const u32 mask = (1u << 13) - 1;
        const size_t smem = (size_t)(mask + 1) * 8;
        CK(cudaFuncSetAttribute(k_smem_atomics64<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_smem_atomics64<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int mode = 0; mode < 2; ++mode) {
            auto launch = [&]() {
                if (mode == 0) k_smem_atomics64<0><<<sms, threads, smem>>>(d_out, iters, mask);
                else k_smem_atomics64<1><<<sms, threads, smem>>>(d_out, iters, mask);
            };
            const float ms = time_ms(launch);
            CK(cudaGetLastError());
            const double ops = (double)sms * threads * iters * (mode ? 2.0 : 1.0);
            printf("smem_atomics64 mode=%d (%s): %.3f ms, %.1f Gatom/s chip, %.2f atom/clk/SM @%.0f MHz\n", mode,
                   mode ? "random, 2-word carry chain" : "random, fire-and-forget", ms, ops / ms * 1e-6,
                   ops / ms * 1e-6 / (double)sms / (prop.clockRate * 1e-6), prop.clockRate * 1e-3);
        }
    }
    // ---- INT32 multiply pipe
    {
        u64 *d_sink;
        CK(cudaMalloc(&d_sink, 64));
        const int iters = 4096, threads = 1024, chains = 8;
        const char *names[3] = {"IMAD 32x32+32", "IMAD.WIDE 32x32+64", "64x64->128 product (mul.lo.u64 + mul.hi.u64)"};
        for (int mode = 0; mode < 3; ++mode) {
            auto launch = [&]() {
                if (mode == 0) k_imad<0, chains><<<sms, threads>>>(d_sink, iters, 12345u);
                else if (mode == 1) k_imad<1, chains><<<sms, threads>>>(d_sink, iters, 12345u);
                else k_imad<2, chains><<<sms, threads>>>(d_sink, iters, 12345u);
            };
            const float ms = time_ms(launch);
            CK(cudaGetLastError());
            const double ops = (double)sms * threads * iters * chains;
            printf("int_pipe %s: %.3f ms, %.1f Gop/s chip, %.2f lane-op/clk/SM @%.0f MHz\n", names[mode], ms, ops / ms * 1e-6,
                   ops / ms * 1e-6 / (double)sms / (prop.clockRate * 1e-6), prop.clockRate * 1e-3);
        }
        CK(cudaFree(d_sink));
    }
    printf("done\n");
    return 0;
}
