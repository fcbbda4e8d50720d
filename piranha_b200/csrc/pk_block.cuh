// Block path (PK_ALGO_BLOCK): the zone scheme with zones aligned to the digits of the tight code, so that the
// double loop over terms becomes a list of small dense tiles -- the GPU form of blocked_multiplication
// (base_series_multiplier.hpp:449-504) inside the zone scheme of polynomial.hpp:1783-1966.
//
// The tight code of a term is a mixed-radix number whose digit k has radix (span_a + span_b) / step + 1, so a sum of
// two operand codes never carries from one digit into the next.  Cut the digits at a split index s:
//     code = hi * S + lo,   S = tradix_0 * ... * tradix_{s-1},   and   (a + b).hi = a.hi + b.hi,  (a + b).lo = a.lo + b.lo.
// Terms with equal `hi` are contiguous in the sorted operands (a "group").  Every product of group ga of A with group
// gb of B lands in output block h = hi(ga) + hi(gb), which owns the S consecutive codes [h*S, (h+1)*S).  So
//   * a zone = Hz consecutive blocks, accumulated in ONE thread block's shared memory exactly as in pk_zone.cuh
//     (dense slots, or bitmap + popcount-rank slots for sparse ranges; 32-bit word planes; ATOMS with carry chain);
//   * the work of a zone is a list of tiles (rows [a0, a0+an) of A) x (group gb of B): no binary search, no cursor,
//     no range test per product, and every lane of a warp does useful work (wide tiles: lane = term of gb, loop over
//     the rows; narrow tiles: the an x bn products are flattened over the lanes);
//   * the tile list is built on the device for the whole product at once: count per block -> exclusive scan ->
//     scatter (a counting sort by h), together with the exact work of every zone, which orders the zones
//     heaviest-first (LPT) for the dynamic zone claiming.
// Tile-level range pruning (this call's output partition [lo, hi), rounded to whole blocks) happens when the list is
// built: pairs of groups outside the partition never become tiles.
#pragma once

#include <cub/device/device_scan.cuh>

#include "pk_zone.cuh"

namespace pk
{

struct BlockArgs {
    ZoneArgs z;           // operands, accumulator geometry, output staging (shared with the zone path)
    const uint4 *tiles;   // {a0, b0, an | bn << 10, block index relative to h_base}, grouped by block
    const u32 *tstart;    // n_h + 1 entries: first tile of every block
    const u32 *order;     // zones, heaviest first (nullptr = natural order)
    const u32 *n_active;  // zones with at least one tile (device value)
    u32 h_base, n_h;      // blocks [h_base, h_base + n_h) belong to this call's partition
    u32 Hz;               // blocks per zone
    u32 S;                // codes per block
};

// code / S for 32-bit codes with magic = floor(2^64 / S) + 1
__device__ __forceinline__ u32 div_magic(u32 x, u64 magic) { return (u32)__umul64hi((u64)x, magic); }

// Group table of one operand: tab[hi] = {first term, one past the last term}; glist = the hi values that occur.
__global__ void k_block_groups(const uint4 *__restrict__ rec, u32 n, u64 magicS, uint2 *__restrict__ tab,
                               u32 *__restrict__ glist, u32 *__restrict__ gcount)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 hi = div_magic(__ldg(&rec[i].x), magicS);
    const u32 prev = i ? div_magic(__ldg(&rec[i - 1].x), magicS) : 0xFFFFFFFFu;
    const u32 next = (i + 1 < n) ? div_magic(__ldg(&rec[i + 1].x), magicS) : 0xFFFFFFFFu;
    if (hi != prev) {
        tab[hi].x = i;
        glist[atomicAdd(gcount, 1u)] = hi;
    }
    if (hi != next) tab[hi].y = i + 1;
}

constexpr u32 kTileCols = 64; // widest tile, in terms of gb

struct TileArgs {
    const u32 *glistA, *glistB;
    const uint2 *tabA, *tabB;
    u32 GA, GB;
    u32 h_base, n_h; // partition, in blocks
    u32 Hz;
    u32 target;      // products per tile
    u32 *cnt;        // tiles per block (count pass) / cursor per block (scatter pass)
    u64 *zwork;      // products per zone
    const u32 *tstart;
    uint4 *tiles;
    u32 max_tiles;
    MulCounters *ctr;
};

// One thread per pair of groups (ga, gb).  PASS 0: count the tiles of block h = hi(ga) + hi(gb) and add the pair's
// products to its zone.  PASS 1: write the tiles at the block's scanned offset.
template <int PASS>
__global__ void k_block_tiles(const TileArgs t)
{
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (u64)t.GA * t.GB) return;
    const u32 ia = (u32)(p / t.GB), ib = (u32)(p - (u64)ia * t.GB);
    const u32 ha = __ldg(&t.glistA[ia]), hb = __ldg(&t.glistB[ib]);
    const u32 h = ha + hb;
    if (h < t.h_base || h - t.h_base >= t.n_h) return; // range pruning at tile level
    const u32 hr = h - t.h_base;
    const uint2 ra = __ldg(&t.tabA[ha]), rb = __ldg(&t.tabB[hb]);
    const u32 na = ra.y - ra.x, nb = rb.y - rb.x;
    // tile shape: at most kTileCols terms of gb (balanced column chunks) x as many rows as reach `target` products,
    // so that a lane keeps its term of gb in registers over the rows of the tile
    const u32 ncol = (nb + kTileCols - 1) / kTileCols;
    const u32 cw = (nb + ncol - 1) / ncol;
    const u32 rpt = min(1023u, max(1u, t.target / cw)); // rows per tile
    const u32 nrow = (na + rpt - 1) / rpt;
    const u32 nt = nrow * ncol;
    if (PASS == 0) {
        atomicAdd(&t.cnt[hr], nt);
        atomicAdd(&t.zwork[hr / t.Hz], (u64)na * nb);
    } else {
        const u32 pos = __ldg(&t.tstart[hr]) + atomicAdd(&t.cnt[hr], nt);
        if (pos + nt > t.max_tiles) { // cannot happen: max_tiles is a proven bound
            atomicOr(&t.ctr->flags, kFlagTableFull);
            return;
        }
        u32 k = pos;
        for (u32 c = 0; c < ncol; ++c) {
            const u32 c0 = c * cw, cn = min(cw, nb - c0);
            for (u32 r = 0; r < nrow; ++r, ++k) {
                const u32 r0 = r * rpt;
                t.tiles[k] = make_uint4(ra.x + r0, rb.x + c0, min(rpt, na - r0) | (cn << 10), hr);
            }
        }
    }
}

// Zone order for the dynamic claiming: heaviest first (LPT), by a counting sort over 256 logarithmic work classes.
// One block.  Zones without work sort last and are never claimed (n_active).
__global__ void __launch_bounds__(1024) k_block_order(const u64 *__restrict__ zwork, u32 nz, u32 *__restrict__ order,
                                                      u32 *__restrict__ n_active, u32 *__restrict__ n_zones)
{
    __shared__ u32 s_hist[256], s_start[256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    auto cls = [](u64 w) -> u32 {
        if (w == 0) return 0u;
        const u32 e = 63u - (u32)__clzll((long long)w);
        const u32 m = e >= 2 ? (u32)(w >> (e - 2)) & 3u : 0u;
        return min(255u, 4u * e + m + 1u);
    };
    for (u32 z = threadIdx.x; z < nz; z += blockDim.x) atomicAdd(&s_hist[cls(zwork[z])], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (int c = 255; c >= 0; --c) {
            s_start[c] = run;
            run += s_hist[c];
        }
        *n_active = nz - s_hist[0];
        *n_zones = nz;
    }
    __syncthreads();
    for (u32 z = threadIdx.x; z < nz; z += blockDim.x) order[atomicAdd(&s_start[cls(zwork[z])], 1u)] = z;
}

__global__ void k_block_natural_order(u32 nz, u32 *__restrict__ n_active, u32 *__restrict__ n_zones)
{
    *n_active = nz;
    *n_zones = nz;
}

template <bool FULL>
__device__ __forceinline__ uint4 block_load(const uint4 *p)
{
    if (FULL) return __ldg(p);
    return make_uint4(__ldg(&p->x), 0u, 0u, 0u);
}

// Visit every product of one tile: f(a_record, b_record).  FULL = false loads only the codes (bitmap marking).
template <bool FULL, typename F>
__device__ __forceinline__ void block_tile_loop(const uint4 *__restrict__ A, const uint4 *__restrict__ B, const uint4 t,
                                                u32 lane, F &&f)
{
    const u32 a0 = t.x, b0 = t.y, an = t.z & 1023u, bn = t.z >> 10;
    if (bn >= 32u) {
        // wide: lane = term of gb (held in registers), loop over the rows (one broadcast load per row)
        for (u32 j0 = 0; j0 < bn; j0 += 32u) {
            const u32 j = j0 + lane;
            const bool ok = j < bn;
            const uint4 bv = block_load<FULL>(&B[b0 + (ok ? j : 0u)]);
            for (u32 i = 0; i < an; ++i) {
                const uint4 av = block_load<FULL>(&A[a0 + i]);
                if (ok) f(av, bv);
            }
        }
    } else {
        // narrow: the an x bn products are flattened over the lanes; (i, j) advance by 32 without a division
        const u32 total = an * bn;
        const u32 q = 32u / bn, r = 32u - q * bn;
        u32 i = lane / bn, j = lane - i * bn;
        for (u32 idx = lane; idx < total; idx += 32u) {
            const uint4 av = block_load<FULL>(&A[a0 + i]);
            const uint4 bv = block_load<FULL>(&B[b0 + j]);
            f(av, bv);
            i += q;
            j += r;
            if (j >= bn) {
                j -= bn;
                ++i;
            }
        }
    }
}

// MODE 0: dense zones (slot = code - zlo; the plan guarantees Hz * S <= cap, so one pass).
// MODE 1: bitmap zones (slot = rank of the code among the codes that occur; passes of at most cap entries).
template <int NW, int MODE, bool SIGNED, int TPB>
__global__ void __launch_bounds__(TPB, 1024 / TPB) k_block(const BlockArgs ba)
{
    constexpr u32 kThreads = TPB;
    const ZoneArgs &a = ba.z;
    extern __shared__ __align__(16) unsigned char zsmem[];
    u32 *acc = reinterpret_cast<u32 *>(zsmem + a.off_acc);
    u32 *bm = reinterpret_cast<u32 *>(zsmem + a.off_bm);
    u32 *pre = reinterpret_cast<u32 *>(zsmem + a.off_pre);
    u32 *list = reinterpret_cast<u32 *>(zsmem + a.off_list);
    __shared__ u32 s_ws[40];
    __shared__ u64 s_b64[2];
    __shared__ u32 s_b32[4];
    __shared__ u32 s_next; // dynamic tile counter

    const u32 tid = threadIdx.x, lane = tid & 31u;
    const u32 acc_s = smem_u32(acc), bm_s = smem_u32(bm);
    const u32 wstride = a.cap * 4u;
    u64 my_pairs = 0;
    u32 my_maxbits = 0, my_flags = 0;
    const u32 n_active = __ldg(ba.n_active);

    for (u32 i = tid; i < a.cap * NW; i += kThreads) acc[i] = 0;
    if (MODE == 1)
        for (u32 i = tid; i < a.bm_words; i += kThreads) bm[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_b32[0] = atomicAdd(a.chunk_counter, 1u);
        __syncthreads();
        const u32 claimed = s_b32[0];
        __syncthreads();
        if (claimed >= n_active) break;
        const u32 z = ba.order ? __ldg(&ba.order[claimed]) : claimed;
        const u32 hr0 = z * ba.Hz, hr1 = min(ba.n_h, hr0 + ba.Hz);
        const u32 zlo = (ba.h_base + hr0) * ba.S, zhi = (ba.h_base + hr1) * ba.S;
        const u32 t0 = __ldg(&ba.tstart[hr0]), t1 = __ldg(&ba.tstart[hr1]);
        if (t0 == t1) continue; // (only in natural order: a zone without tiles)
        const u32 zwords = (MODE == 1) ? min(a.bm_words, ((zhi - zlo + 127u) >> 7) << 2) : 0u;

        // ---------------- walk 1 (bitmap zones): mark the codes that occur
        u32 zone_count = 0;
        if (MODE == 1) {
            if (tid == 0) s_next = t0;
            __syncthreads();
            for (;;) {
                u32 k = 0;
                if (lane == 0) k = atomicAdd(&s_next, 1u);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k >= t1) break;
                const uint4 t = __ldg(&ba.tiles[k]);
                block_tile_loop<false>(a.A, a.B, t, lane, [&](const uint4 &av, const uint4 &bv) {
                    const u32 slot = av.x + bv.x - zlo;
                    reds_or(bm_s + (slot >> 5) * 4u, 1u << (slot & 31u));
                });
            }
            __syncthreads();
            const u32 n_groups = zwords >> 2;
            const u32 gpt = (n_groups + kThreads - 1) / kThreads;
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
                if (r1 < n_entries) { // the code of rank r1 bounds this pass
                    if (tid == 0) {
                        u32 lo_g = 0, hi_g = n_groups;
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
                const u32 gpt = (n_groups + kThreads - 1) / kThreads;
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
            if (tid == 0) s_next = t0;
            __syncthreads();
            for (;;) {
                u32 k = 0;
                if (lane == 0) k = atomicAdd(&s_next, 1u);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k >= t1) break;
                const uint4 t = __ldg(&ba.tiles[k]);
                if (lane == 0 && r0 == 0) my_pairs += (u64)(t.z & 1023u) * (t.z >> 10);
                block_tile_loop<true>(a.A, a.B, t, lane, [&](const uint4 &av, const uint4 &bv) {
                    const u32 code = av.x + bv.x;
                    if (!single && (code < c0 || code >= c1)) return;
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
                    const u64 ma = (u64)av.z | ((u64)av.w << 32), mb = (u64)bv.z | ((u64)bv.w << 32);
                    const u64 plo = ma * mb, phi = __umul64hi(ma, mb);
                    zone_accumulate<NW, SIGNED>(acc_s + e * 4u, wstride, (u32)plo, (u32)(plo >> 32), (u32)phi,
                                                (u32)(phi >> 32), a.pw, SIGNED && ((av.y ^ bv.y) != 0));
                });
            }
            __syncthreads();

            // ---- emit the entries [0, n_pass) of this pass in ascending code order
            const u32 n_pass = r1 - r0;
            if (!have_base) { // claim staging space: the exact count of a single-pass zone, else the entry count
                u32 nz = 0;
                if (single) {
                    for (u32 e = tid; e < n_pass; e += kThreads) {
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
            for (u32 base = 0; base < n_pass; base += kThreads) {
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
        if (MODE == 1) {
            for (u32 i = tid; i < zwords; i += kThreads) bm[i] = 0;
        }
        __syncthreads();
    }
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

// ------------------------------------------------------------------------------------------------ host side

template <int NW, int MODE, bool SIGNED, int TPB>
struct BlockKernel {
    static void set_smem(size_t bytes)
    {
        cudaFuncSetAttribute(k_block<NW, MODE, SIGNED, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    static void launch(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem)
    {
        PK_LAUNCH(ctx, (k_block<NW, MODE, SIGNED, TPB>), grid, TPB, smem, ba);
    }
};

template <int NW, int TPB>
void block_set_smem_nt(size_t b)
{
    BlockKernel<NW, 0, false, TPB>::set_smem(b);
    BlockKernel<NW, 0, true, TPB>::set_smem(b);
    BlockKernel<NW, 1, false, TPB>::set_smem(b);
    BlockKernel<NW, 1, true, TPB>::set_smem(b);
}

template <int NW>
void block_set_smem_nw(size_t optin)
{
    block_set_smem_nt<NW, 1024>(optin);
    block_set_smem_nt<NW, 512>(optin);
}

template <int NW, int TPB>
void block_launch_nt(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem, u32 mode, bool is_signed)
{
    if (mode == 0) {
        if (is_signed) BlockKernel<NW, 0, true, TPB>::launch(ctx, ba, grid, smem);
        else BlockKernel<NW, 0, false, TPB>::launch(ctx, ba, grid, smem);
    } else {
        if (is_signed) BlockKernel<NW, 1, true, TPB>::launch(ctx, ba, grid, smem);
        else BlockKernel<NW, 1, false, TPB>::launch(ctx, ba, grid, smem);
    }
}

template <int NW>
void block_launch_nw(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem, u32 mode, bool is_signed, u32 tpb)
{
    if (tpb == 512) block_launch_nt<NW, 512>(ctx, ba, grid, smem, mode, is_signed);
    else block_launch_nt<NW, 1024>(ctx, ba, grid, smem, mode, is_signed);
}

inline void block_configure(pk_ctx *ctx)
{
    BlockTuning &t = ctx->block_tuning;
    t.disable = env_u32("PK_BLOCK_DISABLE", t.disable);
    t.tile_target = std::max<u32>(32, env_u32("PK_BLOCK_TARGET", t.tile_target));
    t.zone_work = std::max<u32>(1, env_u32("PK_BLOCK_ZONE_WORK", t.zone_work));
    t.force_mode = env_u32("PK_BLOCK_MODE", t.force_mode);
    t.tpb = env_u32("PK_BLOCK_TPB", t.tpb);
    t.natural = env_u32("PK_BLOCK_NATURAL", t.natural);
    const size_t dyn = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
    block_set_smem_nw<2>(dyn);
    block_set_smem_nw<3>(dyn);
    block_set_smem_nw<4>(dyn);
    block_set_smem_nw<5>(dyn);
    block_set_smem_nw<6>(dyn);
}

struct BlockPlan {
    bool ok = false;
    u32 NW = 0, mode = 0, tpb = 1024, grid = 0;
    u32 split = 0; // digits [0, split) form `lo`
    u32 S = 0;     // codes per block
    u32 Hz_max = 1;
    u32 cap = 0, bm_words = 0;
    u32 off_acc = 0, off_bm = 0, off_pre = 0, off_list = 0;
    size_t smem = 0;
};

// The geometry of the block path for a pair of operands.  A function of the operands (tight radices, widths) and of
// the device only -- NOT of the output partition -- so every rank of a partitioned product derives the same S and
// the partition cuts can be rounded to whole blocks consistently.
inline BlockPlan block_plan(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB)
{
    const BlockTuning &t = ctx->block_tuning;
    BlockPlan bp;
    if (t.disable || pl.LA != 1 || pl.LB != 1 || pl.n_vars < 2) return bp;
    if (nA >= (1u << 22) || nB >= (1u << 22)) return bp; // tile records pack the column count in 22 bits
    if (pl.range > 0xFFFFFFFEull) return bp;
    u32 NW = (pl.prod_bits + pl.cnt_bits + 1 + 31) / 32;
    NW = std::max<u32>(NW, 2);
    if (NW > 6) return bp;
    const u32 tpb = (t.tpb == 512) ? 512 : 1024;
    const u32 ctas = 1024 / tpb;
    size_t budget = (ctx->smem_per_sm > ctas * 1024 ? (ctx->smem_per_sm - ctas * 1024) / ctas : 0);
    budget = std::min(budget, ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0);
    budget = budget > 1024 ? (budget - 512) & ~255ull : 0;
    if (budget < (16u << 10)) return bp;
    const double W = (double)nA * (double)nB;
    // dense zones when there is at least about one product per three codes, else bitmap zones
    u32 mode = (double)pl.range <= 3.0 * W ? 0u : 1u;
    if (t.force_mode == 1) mode = 0;
    if (t.force_mode == 2) mode = 1;
    const u64 cap_dense = (budget / (NW * 4)) & ~31ull;
    const u64 span_bitmap = 1ull << ctx->zone_tuning.span_log2_max;
    const u64 limit = mode == 0 ? cap_dense : span_bitmap;
    // the largest split whose block still fits; at least one digit must stay in `hi`
    u64 S = 1;
    u32 split = 0;
    for (u32 v = 0; v + 1 < pl.n_vars; ++v) {
        const u64 next = S * pl.cpR.tradix[v];
        if (next > limit) break;
        S = next;
        split = v + 1;
    }
    if (split == 0 || S < 32) return bp;                       // blocks too small to be worth a tile list
    if (pl.range / S > (8u << 20)) return bp;                  // group tables are indexed by hi
    if (pl.range + 2 * S >= 0xFFFFFFFFull) return bp;          // zone bounds are 32-bit
    bp.NW = NW;
    bp.mode = mode;
    bp.tpb = tpb;
    bp.split = split;
    bp.S = (u32)S;
    if (mode == 0) {
        bp.Hz_max = (u32)std::max<u64>(1, cap_dense / S);
        bp.cap = (u32)cap_dense; // trimmed to Hz * S at launch
        bp.off_acc = 0;
    } else {
        bp.Hz_max = (u32)std::max<u64>(1, span_bitmap / S);
    }
    bp.grid = (u32)ctx->sm_count * ctas;
    bp.ok = true;
    return bp;
}

// Shared-memory layout once Hz is known.  Returns false when the zone does not fit.
inline bool block_layout(pk_ctx *ctx, BlockPlan &bp, u32 Hz)
{
    const u32 ctas = 1024 / bp.tpb;
    size_t budget = (ctx->smem_per_sm > ctas * 1024 ? (ctx->smem_per_sm - ctas * 1024) / ctas : 0);
    budget = std::min(budget, ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0);
    budget = budget > 1024 ? (budget - 512) & ~255ull : 0;
    const u64 span = (u64)Hz * bp.S;
    const u32 NW = bp.NW;
    size_t off = 0;
    if (bp.mode == 0) {
        bp.cap = (u32)((span + 31) & ~31ull);
        bp.off_acc = 0;
        off = (size_t)bp.cap * NW * 4;
        bp.bm_words = 0;
    } else {
        const u64 span128 = (span + 127) & ~127ull;
        const size_t bm_bytes = span128 / 8, pre_bytes = span128 / 32;
        if (bm_bytes + pre_bytes + (size_t)(NW + 1) * 4 * 64 + 64 > budget) return false;
        bp.bm_words = (u32)(span128 / 32);
        bp.off_bm = 0;
        off = bm_bytes;
        bp.off_pre = (u32)off;
        off += (pre_bytes + 15) & ~15ull;
        u64 cap = ((budget - off - 16) / ((NW + 1) * 4)) & ~31ull;
        cap = std::min<u64>(cap, span128);
        if (ctx->zone_tuning.force_cap) cap = std::max<u64>(1, std::min<u64>(cap, ctx->zone_tuning.force_cap));
        bp.cap = (u32)cap;
        bp.off_list = (u32)off;
        off += ((size_t)bp.cap * 4 + 15) & ~15ull;
        bp.off_acc = (u32)off;
        off += (size_t)bp.cap * NW * 4;
    }
    bp.smem = off;
    return off <= budget;
}

// Returns the product restricted to [lo, hi) (both multiples of S), or nullptr when the operands do not suit the
// block path (tiny groups); the caller then uses the generic zone path.
inline pk_dpoly *block_multiply(pk_ctx *ctx, const Plan &pl, BlockPlan bp, const pk_dpoly *A, const pk_dpoly *B,
                                const u64 *keyA, const u64 *keyB, u64 lo, u64 hi, u32 nv)
{
    const u64 nA = A->n, nB = B->n;
    const BlockTuning &bt = ctx->block_tuning;
    const u32 S = bp.S;
    if (lo % S != 0 || (hi % S != 0 && hi != pl.range) || hi <= lo) return nullptr;
    const u32 h_base = (u32)(lo / S);
    const u32 n_h = (u32)((hi - lo + S - 1) / S);
    const unsigned __int128 w_bound = (unsigned __int128)nA * nB;
    const u64 out_cap = (u64)std::min<unsigned __int128>(w_bound, hi - lo);
    if ((unsigned __int128)out_cap * (9 + 8 * pl.out_limbs) > (unsigned __int128)ctx->total_mem / 2) return nullptr;

    DevBuf pa(ctx, (nA + 1) * 16), pb(ctx, (nB + 1) * 16);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nA, 256), 256, 0, keyA, A->planes, pa.as<uint4>(), nA);
    PK_LAUNCH(ctx, k_pack_operand, grid_for(nB, 256), 256, 0, keyB, B->planes, pb.as<uint4>(), nB);

    // ---- groups of equal hi
    const u64 magicS = (u64)((((unsigned __int128)1) << 64) / S) + 1;
    const u64 HR = pl.range / S + 2;
    DevBuf tabA(ctx, HR * 8), tabB(ctx, HR * 8), glA(ctx, nA * 4), glB(ctx, nB * 4), gcnt(ctx, 16);
    CUDA_CHECK(cudaMemsetAsync(tabA.p, 0, HR * 8, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(tabB.p, 0, HR * 8, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(gcnt.p, 0, 16, ctx->stream));
    PK_LAUNCH(ctx, k_block_groups, grid_for(nA, 256), 256, 0, pa.as<uint4>(), (u32)nA, magicS, tabA.as<uint2>(), glA.as<u32>(),
              gcnt.as<u32>());
    PK_LAUNCH(ctx, k_block_groups, grid_for(nB, 256), 256, 0, pb.as<uint4>(), (u32)nB, magicS, tabB.as<uint2>(), glB.as<u32>(),
              gcnt.as<u32>() + 1);
    u32 *hg = reinterpret_cast<u32 *>(static_cast<char *>(ctx->h_scratch) + 2048);
    CUDA_CHECK(cudaMemcpyAsync(hg, gcnt.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const u32 GA = hg[0], GB = hg[1];
    const u64 pairs = (u64)GA * GB;
    // tiny groups: the tile list would be as long as the product itself
    if ((unsigned __int128)pairs * bt.min_tile > w_bound || pairs >= (1ull << 31)) return nullptr;

    // ---- zone geometry: about zone_work products per zone, at least 4 zones per thread block when possible
    const double avg_h = (double)w_bound / (double)n_h;
    u64 Hz = (u64)std::max(1.0, (double)bt.zone_work / std::max(1.0, avg_h));
    Hz = std::min<u64>(Hz, std::max<u64>(1, n_h / (4ull * bp.grid)));
    Hz = std::max<u64>(1, std::min<u64>(Hz, bp.Hz_max));
    if (bt.force_hz) Hz = std::max<u64>(1, std::min<u64>(bt.force_hz, bp.Hz_max));
    while (!block_layout(ctx, bp, (u32)Hz)) {
        if (Hz == 1) return nullptr;
        Hz = (Hz + 1) / 2;
    }
    const u32 nz = (u32)((n_h + Hz - 1) / Hz);

    // ---- tile list: count -> scan -> scatter
    const u64 max_tiles64 = 4 * (u64)(w_bound / bt.tile_target) + 2 * pairs + 2 * ((u64)GA * nB) / kTileCols + 64;
    if (max_tiles64 >= (1ull << 31)) return nullptr;
    const u32 max_tiles = (u32)max_tiles64;
    DevBuf cnt(ctx, ((size_t)n_h + 1) * 4), tstart(ctx, ((size_t)n_h + 1) * 4), tiles(ctx, (size_t)max_tiles * 16);
    // bookkeeping: zone work (u64), offsets (u64), prefix (u64) per zone; bump; counters; zone count + order (u32)
    DevBuf book(ctx, (size_t)nz * 32 + 128);
    u64 *zwork = book.as<u64>();
    u64 *zone_off = zwork + nz;
    u64 *zone_prefix = zone_off + nz;
    u64 *bump = zone_prefix + nz;
    u32 *chunk_counter = reinterpret_cast<u32 *>(bump + 1);
    u32 *n_zones_dev = chunk_counter + 1;
    u32 *n_active_dev = chunk_counter + 2;
    u32 *zone_count = chunk_counter + 4;
    u32 *order = zone_count + nz;
    CUDA_CHECK(cudaMemsetAsync(book.p, 0, (size_t)nz * 32 + 128, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, ((size_t)n_h + 1) * 4, ctx->stream));
    TileArgs ta{};
    ta.glistA = glA.as<u32>();
    ta.glistB = glB.as<u32>();
    ta.tabA = tabA.as<uint2>();
    ta.tabB = tabB.as<uint2>();
    ta.GA = GA;
    ta.GB = GB;
    ta.h_base = h_base;
    ta.n_h = n_h;
    ta.Hz = (u32)Hz;
    ta.target = bt.tile_target;
    ta.cnt = cnt.as<u32>();
    ta.zwork = zwork;
    ta.tstart = tstart.as<u32>();
    ta.tiles = tiles.as<uint4>();
    ta.max_tiles = max_tiles;
    ta.ctr = ctx->d_ctr;
    const u32 pgrid = grid_for(pairs, 256);
    PK_LAUNCH(ctx, (k_block_tiles<0>), pgrid, 256, 0, ta);
    {
        size_t tmp_bytes = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.as<u32>(), tstart.as<u32>(), (int)(n_h + 1), ctx->stream));
        DevBuf tmp(ctx, tmp_bytes);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.as<u32>(), tstart.as<u32>(), (int)(n_h + 1), ctx->stream));
    }
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, ((size_t)n_h + 1) * 4, ctx->stream));
    PK_LAUNCH(ctx, (k_block_tiles<1>), pgrid, 256, 0, ta);
    const bool lpt = !bt.natural && nz <= (1u << 17);
    if (lpt) PK_LAUNCH(ctx, k_block_order, 1, 1024, 0, zwork, nz, order, n_active_dev, n_zones_dev);
    else PK_LAUNCH(ctx, k_block_natural_order, 1, 1, 0, nz, n_active_dev, n_zones_dev);

    DPolyGuard g{ctx, new_dpoly(ctx, out_cap, nv, pl.out_limbs)};
    BlockArgs ba{};
    ZoneArgs &za = ba.z;
    za.A = pa.as<uint4>();
    za.nA = (u32)nA;
    za.B = pb.as<uint4>();
    za.nB = (u32)nB;
    za.lo = (u32)lo;
    za.hi = (u32)hi;
    za.cap = bp.cap;
    za.pw = std::min<u32>(4, (pl.prod_bits + 31) / 32);
    za.bm_words = bp.bm_words;
    za.off_acc = bp.off_acc;
    za.off_bm = bp.off_bm;
    za.off_pre = bp.off_pre;
    za.off_list = bp.off_list;
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
    ba.tiles = tiles.as<uint4>();
    ba.tstart = tstart.as<u32>();
    ba.order = lpt ? order : nullptr;
    ba.n_active = n_active_dev;
    ba.h_base = h_base;
    ba.n_h = n_h;
    ba.Hz = (u32)Hz;
    ba.S = S;
    const bool is_signed = A->any_neg || B->any_neg;
    const u32 grid = std::min<u32>(bp.grid, nz);
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
    switch (bp.NW) {
    case 2: block_launch_nw<2>(ctx, ba, grid, bp.smem, bp.mode, is_signed, bp.tpb); break;
    case 3: block_launch_nw<3>(ctx, ba, grid, bp.smem, bp.mode, is_signed, bp.tpb); break;
    case 4: block_launch_nw<4>(ctx, ba, grid, bp.smem, bp.mode, is_signed, bp.tpb); break;
    case 5: block_launch_nw<5>(ctx, ba, grid, bp.smem, bp.mode, is_signed, bp.tpb); break;
    default: block_launch_nw<6>(ctx, ba, grid, bp.smem, bp.mode, is_signed, bp.tpb); break;
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
    PK_LAUNCH(ctx, k_zone_scan, 1, 1024, 0, zone_count, zone_prefix, n_zones_dev, &ctx->d_ctr->n_out);
    MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
    CUDA_CHECK(cudaMemcpyAsync(hc, ctx->d_ctr, sizeof(MulCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const u64 n_out = hc->n_out;
    if (n_out > out_cap || (hc->flags & kFlagTableFull)) fail(PK_E_CUDA, "block path: a count exceeds its bound (internal error)");
    ctx->stats.table_slots = (u64)bp.cap;
    ctx->stats.acc_lanes = bp.NW;
    ctx->stats.chunk_bits = 0;
    ctx->stats.retries = bp.mode | (bp.tpb << 8);
    DPolyGuard e{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
    if (n_out) {
        const u32 ggrid = std::min<u32>(nz, (u32)ctx->sm_count * 16);
        PK_LAUNCH(ctx, k_zone_gather, ggrid, 256, 0, zone_count, zone_off, zone_prefix, n_zones_dev, g.p->key, g.p->planes,
                  g.p->stride, g.p->sign, e.p->key, e.p->planes, e.p->stride, e.p->sign, pl.out_limbs);
    }
    return e.release();
}

} // namespace pk
