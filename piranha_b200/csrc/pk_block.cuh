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

#include <chrono>

#include <cub/device/device_scan.cuh>

#include "pk_zone.cuh"

namespace pk
{

struct TruncDev;

struct BlockArgs {
    ZoneArgs z;           // operands, accumulator geometry, output staging (shared with the zone path)
    const uint4 *tiles;   // {a0, b0, an | bn << 10, block index relative to h_base}, grouped by block
    const u32 *tstart;    // n_h + 1 entries: first tile of every block
    const u32 *order;     // zones, heaviest first (nullptr = natural order)
    const u32 *n_active;  // zones with at least one tile (device value)
    u32 h_base, n_h;      // blocks [h_base, h_base + n_h) belong to this call's partition
    const u32 *zstart;    // first block of every zone (n_zones + 1 entries, ascending)
    u32 S;                // codes per block
    const TruncDev *trunc; // degree truncation (nullptr = none)
    u32 off_keys;          // MODE 1: byte offset of the hash keys (cap words) in dynamic shared memory
    u32 hash;              // MODE 1: 1 = try the one-walk hash accumulation first
    u32 pair_walk;         // one-limb operands: 1 = the accumulating walk takes two products at a time
};

// Degree truncation in the block path (polynomial.hpp:1402-1531 restated per tile).  Records carry the term's degree as
// an offset from its operand's smallest degree; a pair (i, j) is multiplied iff off_a + off_b <= dlim, where
// dlim = max_degree - min_deg_a - min_deg_b.  Groups record their smallest and largest offset, so a pair of groups is
// dropped (no tile) when even its smallest degrees exceed the limit and walked without any test when its largest fit.
struct TruncDev {
    i64 min_a, max_a, min_b, max_b; // degree extremes of the operands
    u32 dlim;                       // limit on the sum of offsets
    u32 none;                       // 1 = no pair can pass (the limit is below the smallest degree sum)
};

__global__ void k_trunc_init(TruncDev *t)
{
    t->min_a = t->min_b = INT64_MAX;
    t->max_a = t->max_b = INT64_MIN;
    t->dlim = 0;
    t->none = 0;
}

__global__ void k_deg_minmax(const i64 *__restrict__ deg, u32 n, i64 *mn, i64 *mx)
{
    i64 lo = INT64_MAX, hi = INT64_MIN;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const i64 d = deg[i];
        lo = min(lo, d);
        hi = max(hi, d);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicMin(mn, lo);
        atomicMax(mx, hi);
    }
}

// dlim, and the reference's checked subtraction max_degree - d1 over the first operand (degree_sub -> safe_int_sub,
// polynomial.hpp:1243-1252, :1504): it overflows for some term iff it overflows at one of the two extremes.
__global__ void k_trunc_limits(TruncDev *t, i64 max_degree, MulCounters *ctr)
{
    const i64 ext[2] = {t->min_a, t->max_a};
    for (int k = 0; k < 2; ++k) {
        const i64 c = (i64)((u64)max_degree - (u64)ext[k]);
        if (((max_degree ^ ext[k]) & (max_degree ^ c)) < 0) atomicOr(&ctr->flags, kFlagDegreeOverflow);
    }
    const __int128 lim = (__int128)max_degree - t->min_a - t->min_b;
    if (lim < 0) {
        t->none = 1;
        t->dlim = 0;
    } else {
        t->dlim = lim > (__int128)0xFFFFFFFEu ? 0xFFFFFFFEu : (u32)lim;
    }
}

// packed record with the degree offset: y = sign | offset << 1
__global__ void k_pack_operand_deg(const u64 *__restrict__ key, const u64 *__restrict__ m0, const i64 *__restrict__ deg,
                                   const i64 *__restrict__ deg_min, uint4 *__restrict__ out, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 k = key[i], m = m0[i];
        const u32 off = (u32)(deg[i] - *deg_min); // < 2^31: checked on the host from the exponent bounds
        out[i] = make_uint4((u32)(k & 0xFFFFFFFFull), (u32)(k >> 63) | (off << 1), (u32)m, (u32)(m >> 32));
    }
}

// code / S for 32-bit codes with magic = floor(2^64 / S) + 1
__device__ __forceinline__ u32 div_magic(u32 x, u64 magic) { return (u32)__umul64hi((u64)x, magic); }

// Group table of one operand: tab[hi] = {first term, one past the last term}; glist = the hi values that occur.
__global__ void k_block_groups(const uint4 *__restrict__ rec, u32 n, u64 magicS, uint2 *__restrict__ tab,
                               u32 *__restrict__ glist, u32 *__restrict__ gcount, uint2 *__restrict__ gdeg)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 hi = div_magic(__ldg(&rec[i].x), magicS);
    if (gdeg) { // truncation: {smallest, ~largest} degree offset of the group (both by atomicMin from an all-ones fill)
        const u32 off = __ldg(&rec[i].y) >> 1;
        atomicMin(&gdeg[hi].x, off);
        atomicMin(&gdeg[hi].y, ~off);
    }
    const u32 prev = i ? div_magic(__ldg(&rec[i - 1].x), magicS) : 0xFFFFFFFFu;
    const u32 next = (i + 1 < n) ? div_magic(__ldg(&rec[i + 1].x), magicS) : 0xFFFFFFFFu;
    if (hi != prev) {
        tab[hi].x = i;
        glist[atomicAdd(gcount, 1u)] = hi;
    }
    if (hi != next) tab[hi].y = i + 1;
}

// Bank-conflict-free lanes.  In a 32- or 64-term tile lane j holds term j of a chunk of gb and every lane adds into slot
// lo_a + lo_b(j), so the bank of an ATOMS is (const + code_b(j)) mod 32: distinct banks for the 32 lanes iff their codes are
// distinct mod 32.  Terms of a group in code order put 2-3 short rows of the lowest digit on a warp, which overlap by a
// few banks (measured: 2.4 wavefronts per ATOMS).  The order of the terms INSIDE a group is free (a tile is any set of
// terms of gb), so every group of B can be re-ordered by (rank within its residue class, residue class): positions
// 32k .. 32k+31 then hold one term of each class as long as every class has more than k terms.  One warp per group.
// Measured effect on the pairing kernel: -3 % (fateman1), -2 % (fateman2): the shared-memory atomic unit retires about
// 11.4 lane-atomics per clock whether or not the lanes conflict (tools/microbench.cu), so the extra wavefronts that ncu
// counts are not the bound.  Optional (tuning key "block_permute").
__global__ void __launch_bounds__(256) k_block_permute(const uint4 *__restrict__ rec, const uint2 *__restrict__ hi_limb,
                                                       const uint2 *__restrict__ tab, const u32 *__restrict__ glist,
                                                       const u32 *__restrict__ gcount, u32 *__restrict__ rank_tmp,
                                                       uint4 *__restrict__ out, uint2 *__restrict__ out_hi)
{
    __shared__ u32 s_cnt[8][32];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const u32 G = __ldg(gcount);
    for (u32 g = blockIdx.x * 8u + warp; g < G; g += gridDim.x * 8u) {
        const uint2 r = __ldg(&tab[__ldg(&glist[g])]);
        const u32 s = r.x, n = r.y - r.x;
        s_cnt[warp][lane] = 0;
        __syncwarp();
        for (u32 i = lane; i < n; i += 32u) rank_tmp[s + i] = atomicAdd(&s_cnt[warp][__ldg(&rec[s + i].x) & 31u], 1u);
        __syncwarp();
        for (u32 i = lane; i < n; i += 32u) {
            const u32 c = __ldg(&rec[s + i].x) & 31u, rk = rank_tmp[s + i];
            u32 pos = 0; // terms before (rk, c) in (rank, class) order
            for (u32 q = 0; q < 32u; ++q) {
                const u32 cq = s_cnt[warp][q];
                pos += min(cq, rk) + ((q < c && cq > rk) ? 1u : 0u);
            }
            out[s + pos] = __ldg(&rec[s + i]);
            if (hi_limb) out_hi[s + pos] = __ldg(&hi_limb[s + i]);
        }
        __syncwarp();
    }
}

constexpr u32 kTileCols = 32; // a full tile holds exactly one term of gb per lane

struct TileArgs {
    const u32 *glistA, *glistB;
    const uint2 *tabA, *tabB;
    u32 GA, GB;
    u32 h_base, n_h; // partition, in blocks
    u32 target;      // products per tile
    u32 wide;        // 1 = use 64-term chunks of gb (two terms per lane)
    const uint2 *gdegA, *gdegB; // truncation: degree offset extremes per group (nullptr = untruncated)
    const TruncDev *trunc;
    u32 *cnt;        // tiles per block (count pass) / cursor per block (scatter pass)
    const u32 *tstart;
    uint4 *tiles;
    u32 max_tiles;
    MulCounters *ctr;
};

// One thread per pair of groups (ga, gb).  PASS 0: count the tiles of block h = hi(ga) + hi(gb) and add the pair's
// products to its zone.  PASS 1: write the tiles at the block's scanned offset.
// Tile shapes: gb is cut into chunks of exactly 32 terms (FULL tiles: lane = term of gb, held in registers over the
// rows of the tile) plus one ragged chunk of nb mod 32 terms (FLAT tiles: rows x terms flattened over the lanes); rows
// are cut so that a tile has about `target` products.
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
    u32 filt = 0; // bit 31 of the tile word: products of the tile need the degree test
    if (t.trunc) {
        if (t.trunc->none) return;
        const uint2 da = __ldg(&t.gdegA[ha]), db = __ldg(&t.gdegB[hb]);
        const u32 dlim = t.trunc->dlim;
        if ((u64)da.x + db.x > dlim) return; // degree pruning at tile level: not even the smallest degrees fit
        if ((u64)(~da.y) + (~db.y) > dlim) filt = 0x80000000u;
    }
    // gb = n64 chunks of 64 terms (two per lane) + at most one chunk of 32 (one per lane) + a ragged rest (flat)
    const u32 n64 = t.wide ? nb / 64u : 0u;
    const u32 n32 = (nb - n64 * 64u) / kTileCols, rem = nb - n64 * 64u - n32 * kTileCols;
    const u32 rpt_w = min(1023u, max(1u, t.target / 64u));
    const u32 rpt_f = min(1023u, max(1u, t.target / kTileCols));
    const u32 rpt_r = rem ? min(1023u, max(1u, t.target / rem)) : 1u;
    const u32 nrow_w = (na + rpt_w - 1) / rpt_w, nrow_f = (na + rpt_f - 1) / rpt_f, nrow_r = rem ? (na + rpt_r - 1) / rpt_r : 0u;
    const u32 nt = n64 * nrow_w + n32 * nrow_f + nrow_r;
    if (PASS == 0) {
        atomicAdd(&t.cnt[hr], nt);
    } else {
        const u32 pos = __ldg(&t.tstart[hr]) + atomicAdd(&t.cnt[hr], nt);
        if (pos + nt > t.max_tiles) { // cannot happen: max_tiles is a proven bound
            atomicOr(&t.ctr->flags, kFlagTableFull);
            return;
        }
        u32 k = pos, c0 = rb.x;
        for (u32 c = 0; c < n64; ++c, c0 += 64u)
            for (u32 r = 0; r < nrow_w; ++r, ++k)
                t.tiles[k] = make_uint4(ra.x + r * rpt_w, c0, min(rpt_w, na - r * rpt_w) | (64u << 10) | filt, hr);
        for (u32 c = 0; c < n32; ++c, c0 += kTileCols)
            for (u32 r = 0; r < nrow_f; ++r, ++k)
                t.tiles[k] = make_uint4(ra.x + r * rpt_f, c0, min(rpt_f, na - r * rpt_f) | (kTileCols << 10) | filt, hr);
        for (u32 r = 0; r < nrow_r; ++r, ++k)
            t.tiles[k] = make_uint4(ra.x + r * rpt_r, c0, min(rpt_r, na - r * rpt_r) | (rem << 10) | filt, hr);
    }
}

// Zones of Hz consecutive blocks each (dense zones: the slots of Hz blocks are what shared memory holds).
__global__ void k_block_uniform_zones(u32 n_h, u32 Hz, u32 nz, u32 *__restrict__ zstart, u32 *__restrict__ n_zones)
{
    const u32 z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z <= nz) zstart[z] = min(n_h, z * Hz);
    if (z == 0) *n_zones = nz;
}

// Zones of about equal work (bitmap zones): the blocks are cut at the quantiles of the tile count (nq of them), and a
// quantile longer than span_max blocks (the bitmap of a zone covers span_max blocks) is cut into equal pieces.  Where
// blocks are heavy a zone is a few blocks, in the sparse tail it is span_max blocks -- every zone costs a thread block
// about the same time, whatever part of the code range (or which output partition) it comes from.  One block.
__global__ void __launch_bounds__(1024) k_block_equal_zones(const u32 *__restrict__ tstart, u32 n_h, u32 span_max, u32 nq,
                                                            u32 max_zones, u32 *__restrict__ zstart, u32 *__restrict__ n_zones)
{
    __shared__ u32 s_ws[40];
    __shared__ u32 s_run;
    const u32 total = tstart[n_h];
    const u32 per = max(1u, (total + nq - 1) / nq);
    auto bound = [&](u32 q) -> u32 { // first block whose tiles start at or after tile q * per
        if (q == 0) return 0u;
        if (q >= nq) return n_h;
        const u64 want = (u64)q * per;
        u32 lo = 0, hi = n_h;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (tstart[mid] < want) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 base = 0; base < nq; base += blockDim.x) {
        const u32 q = base + threadIdx.x;
        u32 b0 = 0, len = 0, pieces = 0;
        if (q < nq) {
            b0 = bound(q);
            len = bound(q + 1) - b0;
            pieces = (len + span_max - 1) / span_max;
        }
        u32 tot;
        const u32 excl = block_excl_scan(pieces, s_ws, tot);
        const u32 run = s_run;
        if (pieces) {
            const u32 step = (len + pieces - 1) / pieces;
            for (u32 j = 0; j < pieces; ++j)
                if (run + excl + j < max_zones) zstart[run + excl + j] = b0 + j * step;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const u32 nz = min(s_run, max_zones); // (max_zones is a proven bound: nq + n_h / span_max + 1)
        zstart[nz] = n_h;
        *n_zones = nz;
    }
}

// Zone order for the dynamic claiming: heaviest first (LPT), by a counting sort over 256 logarithmic work classes.
// One block.  Zones without work sort last and are never claimed (n_active).
// uniform_hz != 0: the zones are uniform_nz zones of uniform_hz blocks each and this kernel writes zstart / n_zones itself
// (saves the launch of k_block_uniform_zones).
__global__ void __launch_bounds__(1024) k_block_order(const u32 *__restrict__ tstart, u32 *__restrict__ zstart,
                                                      u32 *__restrict__ n_zones, u32 *__restrict__ order,
                                                      u32 *__restrict__ n_active, u32 n_h, u32 uniform_hz, u32 uniform_nz)
{
    if (uniform_hz) {
        for (u32 z = threadIdx.x; z <= uniform_nz; z += blockDim.x) zstart[z] = min(n_h, z * uniform_hz);
        if (threadIdx.x == 0) *n_zones = uniform_nz;
        __syncthreads();
    }
    const u32 nz = *n_zones;
    // work of a zone = its number of tiles (every tile holds about the same number of products)
    auto zwork_of = [&](u32 z) -> u64 { return (u64)(tstart[zstart[z + 1]] - tstart[zstart[z]]); };
    __shared__ u32 s_hist[256], s_start[256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    auto cls = [](u64 w) -> u32 {
        if (w == 0) return 0u;
        const u32 e = 63u - (u32)__clzll((long long)w);
        const u32 m = e >= 2 ? (u32)(w >> (e - 2)) & 3u : 0u;
        return min(255u, 4u * e + m + 1u);
    };
    for (u32 z = threadIdx.x; z < nz; z += blockDim.x) atomicAdd(&s_hist[cls(zwork_of(z))], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (int c = 255; c >= 0; --c) {
            s_start[c] = run;
            run += s_hist[c];
        }
        *n_active = nz - s_hist[0];
    }
    __syncthreads();
    for (u32 z = threadIdx.x; z < nz; z += blockDim.x) order[atomicAdd(&s_start[cls(zwork_of(z))], 1u)] = z;
}

__global__ void k_block_natural_order(const u32 *__restrict__ n_zones, u32 *__restrict__ n_active)
{
    *n_active = *n_zones;
}

// LOAD 0: the code only (bitmap marking), 1: code + sign/degree word (marking with the degree test), 2: whole record
template <int LOAD>
__device__ __forceinline__ uint4 block_load(const uint4 *p)
{
    if (LOAD == 2) return __ldg(p);
    if (LOAD == 1) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        return make_uint4(v.x, v.y, 0u, 0u);
    }
    return make_uint4(__ldg(&p->x), 0u, 0u, 0u);
}

// second limb of a magnitude (operands of up to two limbs keep it in a parallel 8-byte array): loaded only by the
// accumulating walk of the two-limb kernels
template <int LOAD, int LW>
__device__ __forceinline__ uint2 block_load_hi(const uint2 *p, u32 idx)
{
    if (LOAD == 2 && LW == 2) return __ldg(&p[idx]);
    return make_uint2(0u, 0u);
}

// Visit every product of one tile: f(a_record, b_record, a_high_limb, b_high_limb).  FULL selects how much of a record
// is loaded (block_load); LW = limbs per operand magnitude (1 or 2).
template <int FULL, int LW, typename F>
__device__ __forceinline__ void block_tile_loop(const uint4 *__restrict__ A, const uint4 *__restrict__ B,
                                                const uint2 *__restrict__ A1, const uint2 *__restrict__ B1, const uint4 t,
                                                u32 lane, F &&f)
{
    const u32 a0 = t.x, b0 = t.y, an = t.z & 1023u, bn = (t.z >> 10) & 127u;
    if (bn == 64u) {
        // two terms of gb per lane: one broadcast load of a row feeds 64 products
        const uint4 bv0 = block_load<FULL>(&B[b0 + lane]), bv1 = block_load<FULL>(&B[b0 + 32u + lane]);
        const uint2 bh0 = block_load_hi<FULL, LW>(B1, b0 + lane), bh1 = block_load_hi<FULL, LW>(B1, b0 + 32u + lane);
        for (u32 i = 0; i < an; ++i) {
            const uint4 av = block_load<FULL>(&A[a0 + i]);
            const uint2 ah = block_load_hi<FULL, LW>(A1, a0 + i);
            f(av, bv0, ah, bh0);
            f(av, bv1, ah, bh1);
        }
    } else if (bn == kTileCols) {
        // full tile: lane = term of gb (held in registers), loop over the rows (one broadcast load per row)
        const uint4 bv = block_load<FULL>(&B[b0 + lane]);
        const uint2 bh = block_load_hi<FULL, LW>(B1, b0 + lane);
        for (u32 i = 0; i < an; ++i) {
            const uint4 av = block_load<FULL>(&A[a0 + i]);
            const uint2 ah = block_load_hi<FULL, LW>(A1, a0 + i);
            f(av, bv, ah, bh);
        }
    } else {
        // flat tile: the an x bn products are flattened over the lanes; (i, j) advance by 32 without a division
        const u32 total = an * bn;
        // lane / bn and 32 / bn for bn < 32 through one float reciprocal (the +0.5 keeps the quotient far from an integer boundary)
        const float rbn = __frcp_rn((float)bn);
        const u32 q = (u32)(32.5f * rbn), r = 32u - q * bn;
        u32 i = (u32)(((float)lane + 0.5f) * rbn), j = lane - i * bn;
        for (u32 idx = lane; idx < total; idx += 32u) {
            const uint4 av = block_load<FULL>(&A[a0 + i]);
            const uint4 bv = block_load<FULL>(&B[b0 + j]);
            f(av, bv, block_load_hi<FULL, LW>(A1, a0 + i), block_load_hi<FULL, LW>(B1, b0 + j));
            i += q;
            j += r;
            if (j >= bn) {
                j -= bn;
                ++i;
            }
        }
    }
}

// The accumulating walk of one-limb operands visits the products of a tile TWO at a time:
// f2(a0, b0, a1, b1, on1) -- product 0 is always valid, product 1 iff on1.  The two multiply-accumulate chains are
// independent, so their shared-memory atomics (whose returned values feed the carries) are in flight together; one
// product at a time leaves a warp waiting on the ATOMS round trip of every word (the `memory` clobber of the atomics
// also keeps the next product's operand loads behind them).  Used by the dense zones only (see the walk).
template <typename F2>
__device__ __forceinline__ void block_tile_loop2(const uint4 *__restrict__ A, const uint4 *__restrict__ B, const uint4 t, u32 lane,
                                                 F2 &&f2)
{
    const u32 a0 = t.x, b0 = t.y, an = t.z & 1023u, bn = (t.z >> 10) & 127u;
    if (bn == 64u) {
        const uint4 bv0 = __ldg(&B[b0 + lane]), bv1 = __ldg(&B[b0 + 32u + lane]);
        for (u32 i = 0; i < an; ++i) {
            const uint4 av = __ldg(&A[a0 + i]);
            f2(av, bv0, av, bv1, true);
        }
    } else if (bn == kTileCols) {
        const uint4 bv = __ldg(&B[b0 + lane]);
        u32 i = 0;
        for (; i + 1 < an; i += 2) {
            const uint4 av0 = __ldg(&A[a0 + i]), av1 = __ldg(&A[a0 + i + 1]);
            f2(av0, bv, av1, bv, true);
        }
        if (i < an) {
            const uint4 av = __ldg(&A[a0 + i]);
            f2(av, bv, av, bv, false);
        }
    } else {
        // flat tile: positions lane + 32 k of the an x bn products; two consecutive positions of a lane per iteration
        const u32 total = an * bn;
        const float rbn = __frcp_rn((float)bn);
        const u32 q = (u32)(32.5f * rbn), r = 32u - q * bn;
        u32 i = (u32)(((float)lane + 0.5f) * rbn), j = lane - i * bn;
        for (u32 idx = lane; idx < total; idx += 64u) {
            const uint4 av0 = __ldg(&A[a0 + i]), bv0 = __ldg(&B[b0 + j]);
            i += q;
            j += r;
            if (j >= bn) {
                j -= bn;
                ++i;
            }
            const bool on1 = idx + 32u < total;
            const uint4 av1 = on1 ? __ldg(&A[a0 + i]) : av0, bv1 = on1 ? __ldg(&B[b0 + j]) : bv0;
            f2(av0, bv0, av1, bv1, on1);
            i += q;
            j += r;
            if (j >= bn) {
                j -= bn;
                ++i;
            }
        }
    }
}

// One word of an accumulation chain after its atomic returned `old`: x = the word that was added (subtracted when neg),
// ovf = the overflow that came with it, pn = the next word of the product (0 above it).  Leaves the next word to add in x
// and its overflow in ovf.  keep_ovf: the add chain drops the overflow above the product (it cannot occur there).
template <bool SIGNED>
__device__ __forceinline__ void block_chain_next(u32 old, u32 &x, u32 &ovf, bool neg, u32 pn, bool below_pw)
{
    if (SIGNED && neg) {
        const u32 borrow = ((old < x) ? 1u : 0u) | ovf; // x == 0 with ovf set: the word is unchanged, the borrow moves on
        x = pn + borrow;
        ovf = (x < borrow) ? 1u : 0u;
    } else {
        u32 xn, ovfn;
        // xn = pn + ovf + carry(old + x);  ovfn = carry of that sum (ovf and the carry are never both set)
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\taddc.cc.u32 %0, %4, %5;\n\taddc.u32 %1, 0, 0;\n\t}"
            : "=r"(xn), "=r"(ovfn)
            : "r"(old), "r"(x), "r"(pn), "r"(ovf));
        x = xn;
        ovf = below_pw ? ovfn : 0u;
    }
}

// block_accumulate for two independent products: the atomics of word w of both chains are issued back to back, then
// both carries are formed, so the two ATOMS round trips overlap.  on0 / on1 switch a chain off (filtered product).
template <int NW, int PW, bool SIGNED>
__device__ __forceinline__ void block_accumulate2(u32 addr0, const u32 (&p0)[4], bool neg0, bool on0, u32 addr1,
                                                  const u32 (&p1)[4], bool neg1, bool on1, u32 wstride)
{
    u32 x0 = p0[0], x1 = p1[0], o0 = 0, o1 = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w >= PW) { // above the product only a pending carry / borrow travels on
            on0 = on0 && ((x0 | o0) != 0u);
            on1 = on1 && ((x1 | o1) != 0u);
            if (!on0 && !on1) return;
        }
        const u32 v0 = (SIGNED && neg0) ? 0u - x0 : x0, v1 = (SIGNED && neg1) ? 0u - x1 : x1;
        if (w == NW - 1) {
            if (on0) reds_add(addr0, v0);
            if (on1) reds_add(addr1, v1);
            return;
        }
        u32 old0 = 0, old1 = 0;
        if (on0) old0 = atoms_add(addr0, v0);
        if (on1) old1 = atoms_add(addr1, v1);
        const u32 pn0 = (w + 1 < PW) ? p0[(w + 1 < 4) ? w + 1 : 3] : 0u, pn1 = (w + 1 < PW) ? p1[(w + 1 < 4) ? w + 1 : 3] : 0u;
        block_chain_next<SIGNED>(old0, x0, o0, neg0, pn0, w + 1 < PW);
        block_chain_next<SIGNED>(old1, x1, o1, neg1, pn1, w + 1 < PW);
        addr0 += wstride;
        addr1 += wstride;
    }
}

// Add (subtract when neg) the magnitude p[0..PW) -- PW significant 32-bit words, PW <= 8 -- into the NW-word two's
// complement accumulator whose word w lives at saddr0 + w * wstride.  Words below PW always issue their atomic; the
// value an ATOMS returns gives the carry into the next word (one add.cc / addc chain per word, no branches); above PW
// only a pending carry travels on.  The top word's carry-out is the overflow the host's width bound proves impossible.
template <int NW, int PW, bool SIGNED>
__device__ __forceinline__ void block_accumulate(u32 saddr0, u32 wstride, const u32 (&p)[8], bool neg)
{
    u32 addr = saddr0;
    if (SIGNED && neg) {
        u32 x = p[0], ovf = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (w >= PW && x == 0 && ovf == 0) return;
            if (w == NW - 1) {
                reds_add(addr, 0u - x);
                return;
            }
            const u32 old = atoms_add(addr, 0u - x);
            const u32 borrow = ((old < x) ? 1u : 0u) | ovf; // x == 0 with ovf set: the word is unchanged, the borrow moves on
            const u32 pn = (w + 1 < PW) ? p[(w + 1 < 8) ? w + 1 : 7] : 0u;
            x = pn + borrow;
            ovf = (x < borrow) ? 1u : 0u;
            addr += wstride;
        }
    } else {
        u32 x = p[0], ovf = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (w >= PW && x == 0) return; // (ovf needs x == 0xffffffff + 1, impossible above PW where the word is just the carry)
            if (w == NW - 1) {
                reds_add(addr, x);
                return;
            }
            const u32 old = atoms_add(addr, x);
            const u32 pn = (w + 1 < PW) ? p[(w + 1 < 8) ? w + 1 : 7] : 0u;
            u32 xn, ovfn;
            // xn = pn + ovf + carry(old + x);  ovfn = carry of that sum (ovf and the carry are never both set)
            asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\taddc.cc.u32 %0, %4, %5;\n\taddc.u32 %1, 0, 0;\n\t}"
                : "=r"(xn), "=r"(ovfn)
                : "r"(old), "r"(x), "r"(pn), "r"(ovf));
            x = xn;
            ovf = (w + 1 < PW) ? ovfn : 0u;
            addr += wstride;
        }
    }
}

// Open-addressing insert into the zone's shared-memory key table (linear probing; key = slot + 1, 0 = empty).  Returns the
// table index of the slot's accumulator, or 0xFFFFFFFF when the probe limit is hit (the zone then falls back to the
// two-walk bitmap scheme, which handles any number of entries in passes).
constexpr u32 kHashProbeLimit = 32;
__device__ __forceinline__ u32 block_hash_slot(u32 keys_s, u32 cap, u32 slot)
{
    const u32 key = slot + 1u;
    u32 h = __umulhi(slot * 0x9E3779B1u, cap);
    for (u32 probes = 0; probes < kHashProbeLimit; ++probes) {
        u32 cur;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(cur) : "r"(keys_s + h * 4u) : "memory");
        if (cur == key) return h;
        if (cur == 0u) {
            u32 old;
            asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(keys_s + h * 4u), "r"(0u), "r"(key) : "memory");
            if (old == 0u || old == key) return h;
        }
        if (++h == cap) h = 0;
    }
    return 0xFFFFFFFFu;
}

// Dynamic tile claiming for one walk over the tiles [*, tend) of a zone: a warp claims its next tile (shared-memory
// counter) and loads that tile's record BEFORE it processes the current one, so the claim and the record's L2 latency
// hide behind useful work.  The counter must have been set to the first tile before the block-wide barrier.
template <typename F>
__device__ __forceinline__ void block_for_each_tile(const uint4 *__restrict__ tiles, u32 *s_next, u32 tend, u32 lane, F &&body)
{
    u32 k = 0;
    if (lane == 0) k = atomicAdd(s_next, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= tend) return;
    uint4 t = __ldg(&tiles[k]);
    for (;;) {
        u32 kn = 0;
        if (lane == 0) kn = atomicAdd(s_next, 1u);
        kn = __shfl_sync(0xffffffffu, kn, 0);
        uint4 tn = make_uint4(0u, 0u, 0u, 0u);
        if (kn < tend) tn = __ldg(&tiles[kn]);
        body(t);
        if (kn >= tend) return;
        t = tn;
    }
}

// MODE 0: dense zones (slot = code - zlo; the plan guarantees Hz * S <= cap, so one pass).
// MODE 1: bitmap zones (slot = rank of the code among the codes that occur; passes of at most cap entries).  The
//         bitmap is an array of {bits, rank of the word's first bit} pairs, so a rank is one 8-byte shared load, one
//         mask and one popcount.
template <int NW, int PW, int MODE, bool SIGNED, int LW>
__global__ void __launch_bounds__(1024, 1) k_block(const BlockArgs ba)
{
    constexpr u32 kThreads = 1024;
    const ZoneArgs &a = ba.z;
    extern __shared__ __align__(16) unsigned char zsmem[];
    u32 *acc = reinterpret_cast<u32 *>(zsmem + a.off_acc);
    uint2 *bm = reinterpret_cast<uint2 *>(zsmem + a.off_bm); // MODE 1: {bits, prefix}
    u32 *list = reinterpret_cast<u32 *>(zsmem + a.off_list);
    u32 *hkeys = reinterpret_cast<u32 *>(zsmem + ba.off_keys); // MODE 1: keys of the hash attempt
    __shared__ u32 s_ws[40];
    __shared__ u64 s_b64[2];
    __shared__ u32 s_b32[4];
    __shared__ u32 s_next; // dynamic tile counter
    __shared__ u32 s_ovf;  // the hash attempt ran out of probes

    const u32 tid = threadIdx.x, lane = tid & 31u;
    const u32 acc_s = smem_u32(acc), bm_s = smem_u32(bm);
    const u32 wstride = a.cap * 4u;
    u64 my_pairs = 0;
    u32 my_maxbits = 0, my_flags = 0;
    const u32 n_active = __ldg(ba.n_active);
    const u32 dlim = ba.trunc ? __ldg(&ba.trunc->dlim) : 0xFFFFFFFFu;

    for (u32 i = tid; i < a.cap * NW; i += kThreads) acc[i] = 0;
    if (MODE == 1) {
        for (u32 i = tid; i < a.bm_words; i += kThreads) bm[i] = make_uint2(0u, 0u);
        if (ba.hash)
            for (u32 i = tid; i < a.cap; i += kThreads) hkeys[i] = 0;
    }
    __syncthreads();

    for (;;) {
        if (tid == 0) s_b32[0] = atomicAdd(a.chunk_counter, 1u);
        __syncthreads();
        const u32 claimed = s_b32[0];
        __syncthreads();
        if (claimed >= n_active) break;
        const u32 z = ba.order ? __ldg(&ba.order[claimed]) : claimed;
        const u32 hr0 = __ldg(&ba.zstart[z]), hr1 = __ldg(&ba.zstart[z + 1]);
        const u32 zlo = (ba.h_base + hr0) * ba.S, zhi = (ba.h_base + hr1) * ba.S;
        const u32 t0 = __ldg(&ba.tstart[hr0]), t1 = __ldg(&ba.tstart[hr1]);
        if (t0 == t1) continue; // (only in natural order: a zone without tiles)
        const u32 zwords = (MODE == 1) ? min(a.bm_words, (zhi - zlo + 31u) >> 5) : 0u; // bitmap words in use

        // ---------------- sparse zones, experiment (tuning key "block_hash", off): ONE walk that accumulates into a
        // shared-memory hash table keyed by the code (the bitmap is then built from the ENTRIES, not from the products, and
        // only orders the emission).  A zone whose entries do not fit (probe limit) is cleared and redone by the two-walk
        // scheme below.  Measured on pearce1: 0.60 ms against 0.39 ms for the two-walk scheme -- the probe loop leaves the
        // lanes of a warp diverged for the multiply-accumulate that follows (9 of 32 lanes active in the ncu source view),
        // and the heaviest 9 % of the zones overflow the table and pay for both schemes.
        if (MODE == 1 && ba.hash) {
            const u32 keys_s = smem_u32(hkeys);
            if (tid == 0) {
                s_next = t0;
                s_ovf = 0;
            }
            __syncthreads();
            u64 zone_pairs = 0;
            auto hash_accumulate = [&](const uint4 &av, const uint4 &bv, const uint2 &ah, const uint2 &bh) {
                const u32 h = block_hash_slot(keys_s, a.cap, av.x + bv.x - zlo);
                if (h == 0xFFFFFFFFu) {
                    s_ovf = 1;
                    return;
                }
                const u64 ma = (u64)av.z | ((u64)av.w << 32), mb = (u64)bv.z | ((u64)bv.w << 32);
                u32 pw[8];
                if (LW == 1) {
                    const u64 plo = ma * mb, phi = __umul64hi(ma, mb);
                    pw[0] = (u32)plo, pw[1] = (u32)(plo >> 32), pw[2] = (u32)phi, pw[3] = (u32)(phi >> 32);
                    pw[4] = pw[5] = pw[6] = pw[7] = 0u;
                } else {
                    const u64 xa[2] = {ma, (u64)ah.x | ((u64)ah.y << 32)}, xb[2] = {mb, (u64)bh.x | ((u64)bh.y << 32)};
                    u64 pr[4];
                    mag_mul<2, 2>(xa, xb, pr);
#pragma unroll
                    for (int q = 0; q < 4; ++q) pw[2 * q] = (u32)pr[q], pw[2 * q + 1] = (u32)(pr[q] >> 32);
                }
                block_accumulate<NW, PW, SIGNED>(acc_s + h * 4u, wstride, pw, SIGNED && (((av.y ^ bv.y) & 1u) != 0));
            };
            block_for_each_tile(ba.tiles, &s_next, t1, lane, [&](const uint4 &t) {
                if (*((volatile u32 *)&s_ovf)) return; // the attempt is lost: drain the tile list quickly
                if (t.z >> 31) {
                    block_tile_loop<2, LW>(a.A, a.B, a.A1, a.B1, t, lane, [&](const uint4 &av, const uint4 &bv, const uint2 &ah, const uint2 &bh) {
                        if ((av.y >> 1) + (bv.y >> 1) > dlim) return;
                        ++zone_pairs;
                        hash_accumulate(av, bv, ah, bh);
                    });
                } else {
                    if (lane == 0) zone_pairs += (u64)(t.z & 1023u) * ((t.z >> 10) & 127u);
                    block_tile_loop<2, LW>(a.A, a.B, a.A1, a.B1, t, lane, hash_accumulate);
                }
            });
            __syncthreads();
            if (s_ovf == 0) {
                my_pairs += zone_pairs;
                // entries -> bitmap -> ranks -> list of table indices in ascending code order
                for (u32 h = tid; h < a.cap; h += kThreads) {
                    const u32 k = hkeys[h];
                    if (k) reds_or(bm_s + ((k - 1u) >> 5) * 8u, 1u << ((k - 1u) & 31u));
                }
                __syncthreads();
                const u32 wpt = (zwords + kThreads - 1) / kThreads;
                const u32 w0 = min(zwords, tid * wpt), w1 = min(zwords, w0 + wpt);
                u32 local = 0;
                for (u32 w = w0; w < w1; ++w) local += __popc(bm[w].x);
                u32 n_ent;
                u32 run = block_excl_scan(local, s_ws, n_ent);
                for (u32 w = w0; w < w1; ++w) {
                    bm[w].y = run;
                    run += __popc(bm[w].x);
                }
                __syncthreads();
                for (u32 h = tid; h < a.cap; h += kThreads) {
                    const u32 k = hkeys[h];
                    if (k) {
                        const uint2 bw = bm[(k - 1u) >> 5];
                        list[bw.y + __popc(bw.x & ((1u << ((k - 1u) & 31u)) - 1u))] = h;
                    }
                }
                __syncthreads();
                // emission in rank order (one run of consecutive ranks per thread), zeros dropped, table cleared on the way
                const u32 per = ((n_ent + kThreads - 1) / kThreads) | 1u;
                const u32 e0 = min(n_ent, tid * per), e1 = min(n_ent, e0 + per);
                u32 nzc = 0;
                for (u32 e = e0; e < e1; ++e) {
                    const u32 h = list[e];
                    u32 any = 0;
#pragma unroll
                    for (int q = 0; q < NW; ++q) any |= acc[q * a.cap + h];
                    nzc += any != 0;
                }
                u32 total;
                const u32 excl = block_excl_scan(nzc, s_ws, total);
                if (tid == 0) s_b64[1] = atomicAdd(a.bump, (u64)total);
                __syncthreads();
                const u64 zone_base = s_b64[1];
                u64 pos = zone_base + excl;
                for (u32 e = e0; e < e1; ++e) {
                    const u32 h = list[e];
                    u32 wd[NW], any = 0;
#pragma unroll
                    for (int q = 0; q < NW; ++q) {
                        wd[q] = acc[q * a.cap + h];
                        acc[q * a.cap + h] = 0;
                        any |= wd[q];
                    }
                    const u32 slot = hkeys[h] - 1u;
                    hkeys[h] = 0;
                    if (any) {
                        zone_emit<NW>(a, pos, zlo + slot, wd, my_maxbits, my_flags);
                        ++pos;
                    }
                }
                if (tid == 0) {
                    a.zone_count[z] = total;
                    a.zone_off[z] = zone_base;
                }
                for (u32 i = tid; i < zwords; i += kThreads) bm[i].x = 0;
                __syncthreads();
                continue;
            }
            // lost: clear what the attempt left behind, then the two-walk scheme
            if (tid == 0) atomicAdd(&a.ctr->aux, 1ull); // (statistics: zones that fell back)
            for (u32 i = tid; i < a.cap; i += kThreads) hkeys[i] = 0;
            for (u32 i = tid; i < a.cap * NW; i += kThreads) acc[i] = 0;
            __syncthreads();
        }

        // ---------------- walk 1 (bitmap zones): mark the codes that occur
        u32 zone_count = 0;
        if (MODE == 1) {
            if (tid == 0) s_next = t0;
            __syncthreads();
            block_for_each_tile(ba.tiles, &s_next, t1, lane, [&](const uint4 &t) {
                if (t.z >> 31) { // truncation: some products of this tile exceed the degree limit
                    block_tile_loop<1, LW>(a.A, a.B, a.A1, a.B1, t, lane, [&](const uint4 &av, const uint4 &bv, const uint2 &, const uint2 &) {
                        if ((av.y >> 1) + (bv.y >> 1) > dlim) return;
                        ++my_pairs;
                        const u32 slot = av.x + bv.x - zlo;
                        reds_or(bm_s + (slot >> 5) * 8u, 1u << (slot & 31u));
                    });
                } else {
                    if (lane == 0) my_pairs += (u64)(t.z & 1023u) * ((t.z >> 10) & 127u);
                    block_tile_loop<0, LW>(a.A, a.B, a.A1, a.B1, t, lane, [&](const uint4 &av, const uint4 &bv, const uint2 &, const uint2 &) {
                        const u32 slot = av.x + bv.x - zlo;
                        reds_or(bm_s + (slot >> 5) * 8u, 1u << (slot & 31u));
                    });
                }
            });
            __syncthreads();
            // exclusive popcount prefix per bitmap word
            const u32 wpt = (zwords + kThreads - 1) / kThreads;
            const u32 w0 = min(zwords, tid * wpt), w1 = min(zwords, w0 + wpt);
            u32 local = 0;
            for (u32 w = w0; w < w1; ++w) local += __popc(bm[w].x);
            u32 total;
            u32 run = block_excl_scan(local, s_ws, total);
            for (u32 w = w0; w < w1; ++w) {
                bm[w].y = run;
                run += __popc(bm[w].x);
            }
            zone_count = total;
            __syncthreads();
        }

        // ---------------- accumulate + emit, in passes of at most `cap` entries
        const u32 n_entries = (MODE == 1) ? zone_count : (zhi - zlo);
        const bool single = (MODE == 0) || n_entries <= a.cap;
        u64 zone_base = 0, zone_written = 0;
        bool have_base = false;
        u32 r0 = 0;   // first rank (bitmap) / slot (dense) of this pass
        u32 c0 = zlo; // first code of this pass
        do {
            const u32 r1 = min(n_entries, r0 + a.cap);
            u32 c1 = zhi;
            u32 tp0 = t0, tp1 = t1; // tiles of this pass
            if (MODE == 1) {
                if (r1 < n_entries) { // the code of rank r1 bounds this pass
                    if (tid == 0) {
                        u32 lo_w = 0, hi_w = zwords; // last word with prefix <= r1
                        while (hi_w - lo_w > 1) {
                            const u32 mid = (lo_w + hi_w) >> 1;
                            if (bm[mid].y <= r1) lo_w = mid;
                            else hi_w = mid;
                        }
                        u32 rem = r1 - bm[lo_w].y, w = lo_w;
                        while (rem >= (u32)__popc(bm[w].x)) {
                            rem -= __popc(bm[w].x);
                            ++w;
                        }
                        s_b32[1] = zlo + w * 32u + __fns(bm[w].x, 0, (int)rem + 1);
                    }
                    __syncthreads();
                    c1 = s_b32[1];
                }
                if (!single) { // the codes of a pass are consecutive, so only the tiles of its blocks are walked
                    tp0 = __ldg(&ba.tstart[hr0 + (c0 - zlo) / ba.S]);
                    tp1 = __ldg(&ba.tstart[hr0 + (c1 - 1u - zlo) / ba.S + 1u]);
                }
                // rank -> slot list of this pass, so that emission runs with full warps
                const u32 wpt = (zwords + kThreads - 1) / kThreads;
                const u32 w0 = min(zwords, tid * wpt), w1 = min(zwords, w0 + wpt);
                for (u32 w = w0; w < w1; ++w) {
                    const uint2 bw = bm[w];
                    u32 rank = bw.y, bits = bw.x;
                    if (rank >= r1 || rank + (u32)__popc(bits) <= r0) continue;
                    while (bits) {
                        const u32 b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (rank >= r0 && rank < r1) list[rank - r0] = w * 32u + b;
                        ++rank;
                    }
                }
            }
            if (tid == 0) s_next = tp0;
            __syncthreads();
            auto accumulate = [&](const uint4 &av, const uint4 &bv, const uint2 &ah, const uint2 &bh) {
                const u32 code = av.x + bv.x;
                if (MODE == 1 && !single && (code < c0 || code >= c1)) return;
                const u32 slot = code - zlo;
                u32 e;
                if (MODE == 1) {
                    const uint2 bw = bm[slot >> 5];
                    e = bw.y + __popc(bw.x & ((1u << (slot & 31u)) - 1u)) - r0;
                } else {
                    e = slot;
                }
                const u64 ma = (u64)av.z | ((u64)av.w << 32), mb = (u64)bv.z | ((u64)bv.w << 32);
                u32 pw[8];
                if (LW == 1) {
                    const u64 plo = ma * mb, phi = __umul64hi(ma, mb);
                    pw[0] = (u32)plo, pw[1] = (u32)(plo >> 32), pw[2] = (u32)phi, pw[3] = (u32)(phi >> 32);
                    pw[4] = pw[5] = pw[6] = pw[7] = 0u;
                } else {
                    const u64 xa[2] = {ma, (u64)ah.x | ((u64)ah.y << 32)}, xb[2] = {mb, (u64)bh.x | ((u64)bh.y << 32)};
                    u64 pr[4];
                    mag_mul<2, 2>(xa, xb, pr);
#pragma unroll
                    for (int q = 0; q < 4; ++q) pw[2 * q] = (u32)pr[q], pw[2 * q + 1] = (u32)(pr[q] >> 32);
                }
                block_accumulate<NW, PW, SIGNED>(acc_s + e * 4u, wstride, pw, SIGNED && (((av.y ^ bv.y) & 1u) != 0));
            };
            // the same for two products at a time (one-limb operands): slots and products of both first, then the two
            // accumulation chains interleaved word by word
            auto accumulate2 = [&](const uint4 &av0, const uint4 &bv0, bool on0, const uint4 &av1, const uint4 &bv1, bool on1) {
                const u32 code0 = av0.x + bv0.x, code1 = av1.x + bv1.x;
                if (MODE == 1 && !single) {
                    on0 = on0 && code0 >= c0 && code0 < c1;
                    on1 = on1 && code1 >= c0 && code1 < c1;
                }
                const u32 slot0 = on0 ? code0 - zlo : 0u, slot1 = on1 ? code1 - zlo : 0u;
                u32 e0 = slot0, e1 = slot1;
                if (MODE == 1) {
                    const uint2 bw0 = bm[slot0 >> 5], bw1 = bm[slot1 >> 5];
                    e0 = on0 ? bw0.y + __popc(bw0.x & ((1u << (slot0 & 31u)) - 1u)) - r0 : 0u;
                    e1 = on1 ? bw1.y + __popc(bw1.x & ((1u << (slot1 & 31u)) - 1u)) - r0 : 0u;
                }
                const u64 ma0 = (u64)av0.z | ((u64)av0.w << 32), mb0 = (u64)bv0.z | ((u64)bv0.w << 32);
                const u64 ma1 = (u64)av1.z | ((u64)av1.w << 32), mb1 = (u64)bv1.z | ((u64)bv1.w << 32);
                const u64 lo0 = ma0 * mb0, hi0 = __umul64hi(ma0, mb0), lo1 = ma1 * mb1, hi1 = __umul64hi(ma1, mb1);
                const u32 p0[4] = {(u32)lo0, (u32)(lo0 >> 32), (u32)hi0, (u32)(hi0 >> 32)};
                const u32 p1[4] = {(u32)lo1, (u32)(lo1 >> 32), (u32)hi1, (u32)(hi1 >> 32)};
                block_accumulate2<NW, (PW < 4 ? PW : 4), SIGNED>(acc_s + e0 * 4u, p0, SIGNED && (((av0.y ^ bv0.y) & 1u) != 0), on0,
                                                                 acc_s + e1 * 4u, p1, SIGNED && (((av1.y ^ bv1.y) & 1u) != 0), on1, wstride);
            };
            block_for_each_tile(ba.tiles, &s_next, tp1, lane, [&](const uint4 &t) {
                // (measured: fateman2, five-word chains, 3.63 -> 3.43 ms; fateman1 unchanged; bitmap zones 5 % SLOWER --
                // their rank look-ups and range tests double too -- and degree-tested tiles 30 % slower, so only the
                // untested tiles of dense zones take the pair walk)
                if (LW == 1 && MODE == 0 && ba.pair_walk && !(t.z >> 31)) {
                    if (lane == 0) my_pairs += (u64)(t.z & 1023u) * ((t.z >> 10) & 127u);
                    block_tile_loop2(a.A, a.B, t, lane, [&](const uint4 &av0, const uint4 &bv0, const uint4 &av1, const uint4 &bv1, bool on1) {
                        accumulate2(av0, bv0, true, av1, bv1, on1);
                    });
                    return;
                }
                if (t.z >> 31) { // truncation: test the degree of every product of this tile
                    block_tile_loop<2, LW>(a.A, a.B, a.A1, a.B1, t, lane, [&](const uint4 &av, const uint4 &bv, const uint2 &ah, const uint2 &bh) {
                        if ((av.y >> 1) + (bv.y >> 1) > dlim) return;
                        if (MODE == 0) ++my_pairs;
                        accumulate(av, bv, ah, bh);
                    });
                } else {
                    if (MODE == 0 && lane == 0) my_pairs += (u64)(t.z & 1023u) * ((t.z >> 10) & 127u);
                    block_tile_loop<2, LW>(a.A, a.B, a.A1, a.B1, t, lane, accumulate);
                }
            });
            __syncthreads();

            // ---- emit the entries [0, n_pass) of this pass in ascending code order: every thread owns a run of
            // consecutive entries (odd length: conflict-free strided reads), counts its non-zero ones, ONE block scan
            // gives its first output position; staging space is claimed with one atomicAdd per zone (the exact count
            // of a single-pass zone, else the entry count)
            const u32 n_pass = r1 - r0;
            const u32 per = ((n_pass + kThreads - 1) / kThreads) | 1u;
            const u32 e0 = min(n_pass, tid * per), e1 = min(n_pass, e0 + per);
            u32 nzc = 0;
            for (u32 e = e0; e < e1; ++e) {
                u32 any = 0;
#pragma unroll
                for (int q = 0; q < NW; ++q) any |= acc[q * a.cap + e];
                nzc += any != 0;
            }
            u32 total;
            const u32 excl = block_excl_scan(nzc, s_ws, total);
            if (!have_base) {
                if (tid == 0) s_b64[1] = atomicAdd(a.bump, single ? (u64)total : (u64)n_entries);
                __syncthreads();
                zone_base = s_b64[1];
                have_base = true;
            }
            u64 pos = zone_base + zone_written + excl;
            for (u32 e = e0; e < e1; ++e) {
                u32 wd[NW], any = 0;
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    wd[q] = acc[q * a.cap + e];
                    acc[q * a.cap + e] = 0;
                    any |= wd[q];
                }
                if (any) {
                    const u32 slot = (MODE == 1) ? list[e] : e;
                    zone_emit<NW>(a, pos, zlo + slot, wd, my_maxbits, my_flags);
                    ++pos;
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
        if (MODE == 1) {
            for (u32 i = tid; i < zwords; i += kThreads) bm[i].x = 0;
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

template <int NW, int PW, int MODE, bool SIGNED, int LW>
struct BlockKernel {
    static void set_smem(size_t bytes)
    {
        cudaFuncSetAttribute(k_block<NW, PW, MODE, SIGNED, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    static void launch(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem)
    {
        PK_LAUNCH(ctx, (k_block<NW, PW, MODE, SIGNED, LW>), grid, 1024, smem, ba);
    }
};

// One-limb operands: the (accumulator words, product words) pairs the width bound can ask for:
// NW = ceil((prod_bits + cnt_bits + 1) / 32) with cnt_bits <= 31, PW = min(4, ceil(prod_bits / 32)).
#define PK_BLOCK_WIDTHS(X) X(2, 1) X(2, 2) X(3, 2) X(3, 3) X(4, 3) X(4, 4) X(5, 4) X(6, 4)
// Two-limb operands (products of 65 .. 256 bits): PW = ceil(prod_bits / 32) and always one more accumulator word (the
// count bits fit in it), signed accumulation for every sign mix -- fewer instances for the rarer case.
#define PK_BLOCK_WIDTHS2(X) X(4, 3) X(5, 4) X(6, 5) X(7, 6) X(8, 7) X(9, 8)

template <int NW, int PW>
void block_set_smem_w(size_t b)
{
    BlockKernel<NW, PW, 0, false, 1>::set_smem(b);
    BlockKernel<NW, PW, 0, true, 1>::set_smem(b);
    BlockKernel<NW, PW, 1, false, 1>::set_smem(b);
    BlockKernel<NW, PW, 1, true, 1>::set_smem(b);
}

template <int NW, int PW>
void block_set_smem_w2(size_t b)
{
    BlockKernel<NW, PW, 0, true, 2>::set_smem(b);
    BlockKernel<NW, PW, 1, true, 2>::set_smem(b);
}

template <int NW, int PW>
void block_launch_w(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem, u32 mode, bool is_signed)
{
    if (mode == 0) {
        if (is_signed) BlockKernel<NW, PW, 0, true, 1>::launch(ctx, ba, grid, smem);
        else BlockKernel<NW, PW, 0, false, 1>::launch(ctx, ba, grid, smem);
    } else {
        if (is_signed) BlockKernel<NW, PW, 1, true, 1>::launch(ctx, ba, grid, smem);
        else BlockKernel<NW, PW, 1, false, 1>::launch(ctx, ba, grid, smem);
    }
}

template <int NW, int PW>
void block_launch_w2(pk_ctx *ctx, const BlockArgs &ba, u32 grid, size_t smem, u32 mode)
{
    if (mode == 0) BlockKernel<NW, PW, 0, true, 2>::launch(ctx, ba, grid, smem);
    else BlockKernel<NW, PW, 1, true, 2>::launch(ctx, ba, grid, smem);
}

inline bool block_width_supported(u32 NW, u32 PW, u32 LW)
{
#define X(nw, pw) if (LW == 1 && NW == nw && PW == pw) return true;
    PK_BLOCK_WIDTHS(X)
#undef X
#define X(nw, pw) if (LW == 2 && NW == nw && PW == pw) return true;
    PK_BLOCK_WIDTHS2(X)
#undef X
    return false;
}

inline void block_configure(pk_ctx *ctx)
{
    BlockTuning &t = ctx->block_tuning;
    t.disable = env_u32("PK_BLOCK_DISABLE", t.disable);
    t.tile_target = std::max<u32>(32, env_u32("PK_BLOCK_TARGET", t.tile_target));
    t.zone_work = std::max<u32>(1, env_u32("PK_BLOCK_ZONE_WORK", t.zone_work));
    t.force_mode = env_u32("PK_BLOCK_MODE", t.force_mode);
    t.natural = env_u32("PK_BLOCK_NATURAL", t.natural);
    t.pair_walk = env_u32("PK_BLOCK_PAIR_WALK", t.pair_walk);
    t.wide = env_u32("PK_BLOCK_WIDE", t.wide);
    t.permute = env_u32("PK_BLOCK_PERMUTE", t.permute);
    t.hash = env_u32("PK_BLOCK_HASH", t.hash);
    t.smem_limit = env_u32("PK_BLOCK_SMEM", t.smem_limit);
    t.radix0_residue = env_u32("PK_BLOCK_RADIX0", t.radix0_residue);
    const size_t dyn = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
#define X(nw, pw) block_set_smem_w<nw, pw>(dyn);
    PK_BLOCK_WIDTHS(X)
#undef X
#define X(nw, pw) block_set_smem_w2<nw, pw>(dyn);
    PK_BLOCK_WIDTHS2(X)
#undef X
}

struct BlockPlan {
    bool ok = false;
    u32 NW = 0, PW = 0, LW = 1, mode = 0, grid = 0;
    u32 split = 0; // digits [0, split) form `lo`
    u32 S = 0;     // codes per block
    u32 Hz_max = 1;
    u32 cap = 0, bm_words = 0;
    u32 off_acc = 0, off_bm = 0, off_list = 0, off_keys = 0;
    size_t smem = 0;
};

inline size_t block_smem_budget(pk_ctx *ctx)
{
    size_t budget = ctx->smem_per_sm > 1024 ? ctx->smem_per_sm - 1024 : 0;
    budget = std::min(budget, ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0);
    if (ctx->block_tuning.smem_limit) budget = std::min<size_t>(budget, ctx->block_tuning.smem_limit);
    return budget > 1024 ? (budget - 512) & ~255ull : 0; // static shared memory + alignment slack
}

// The geometry of the block path for a pair of operands.  A function of the operands (tight radices, widths) and of
// the device only -- NOT of the output partition -- so every rank of a partitioned product derives the same S and
// the partition cuts can be rounded to whole blocks consistently.
inline BlockPlan block_plan(pk_ctx *ctx, const Plan &pl, u64 nA, u64 nB)
{
    const BlockTuning &t = ctx->block_tuning;
    BlockPlan bp;
    if (t.disable || pl.LA > 2 || pl.LB > 2 || pl.n_vars < 2) return bp;
    if (nA >= (1u << 22) || nB >= (1u << 22)) return bp; // tile records pack the column count in 22 bits
    if (pl.range > 0xFFFFFFFEull) return bp;
    const u32 LW = std::max(pl.LA, pl.LB);
    u32 NW = std::max<u32>(2, (pl.prod_bits + pl.cnt_bits + 1 + 31) / 32);
    u32 PW = std::min<u32>(4, (pl.prod_bits + 31) / 32);
    if (LW == 2) { // (two-limb operands whose product still fits 64 bits are rare: they take PW = 3 like the rest)
        PW = std::max<u32>(3, (pl.prod_bits + 31) / 32);
        NW = PW + 1;
    }
    if (!block_width_supported(NW, PW, LW)) return bp;
    const size_t budget = block_smem_budget(ctx);
    if (budget < (16u << 10)) return bp;
    const double W = (double)nA * (double)nB;
    // dense zones when there is at least about one product per three codes, else bitmap zones
    u32 mode = (double)pl.range <= 3.0 * W ? 0u : 1u;
    if (t.force_mode == 1) mode = 0;
    if (t.force_mode == 2) mode = 1;
    const u64 cap_dense = (budget / (NW * 4)) & ~31ull;
    // bitmap zones: 8 bytes of {bits, prefix} per 32 codes, and room for at least 2048 entries
    u64 span_bitmap = 1ull << ctx->zone_tuning.span_log2_max;
    while (span_bitmap > 1024 && span_bitmap / 4 + (size_t)(NW + 1) * 4 * 2048 > budget) span_bitmap >>= 1;
    const u64 limit = mode == 0 ? cap_dense : span_bitmap;
    // the largest split whose block still fits; at least one digit must stay in `hi`
    // (bitmap zones: a block should also hold its entries in one pass -- the products and the codes of a block bound them --
    // because only multi-block zones can restrict a pass to the tiles of the blocks it covers)
    const u64 cap_bitmap = (budget / 2 / ((NW + 1) * 4)) & ~31ull;
    u64 S = 1;
    u32 split = 0;
    for (u32 v = 0; v + 1 < pl.n_vars; ++v) {
        const u64 next = S * pl.cpR.tradix[v];
        if (next > limit) break;
        if (mode == 1 && split >= 1 && std::min(W / ((double)pl.range / (double)next), (double)next) > (double)cap_bitmap) break;
        S = next;
        split = v + 1;
    }
    if (split == 0 || S < 32) return bp;                       // blocks too small to be worth a tile list
    if (pl.range / S > (8u << 20)) return bp;                  // group tables are indexed by hi
    if (pl.range + 2 * S >= 0xFFFFFFFFull) return bp;          // zone bounds are 32-bit
    bp.NW = NW;
    bp.PW = PW;
    bp.LW = LW;
    bp.mode = mode;
    bp.split = split;
    bp.S = (u32)S;
    bp.Hz_max = (u32)std::max<u64>(1, limit / S);
    bp.grid = (u32)ctx->sm_count;
    bp.ok = true;
    return bp;
}

// Shared-memory layout once Hz is known.  Returns false when the zone does not fit.
inline bool block_layout(pk_ctx *ctx, BlockPlan &bp, u32 Hz)
{
    const size_t budget = block_smem_budget(ctx);
    const u64 span = (u64)Hz * bp.S;
    const u32 NW = bp.NW;
    size_t off = 0;
    if (bp.mode == 0) {
        bp.cap = (u32)((span + 31) & ~31ull);
        bp.off_acc = 0;
        off = (size_t)bp.cap * NW * 4;
        bp.bm_words = 0;
    } else {
        const u64 words = (span + 31) / 32;
        const size_t bm_bytes = (words * 8 + 15) & ~15ull;
        if (bm_bytes + (size_t)(NW + 1) * 4 * 64 + 64 > budget) return false;
        bp.bm_words = (u32)words;
        bp.off_bm = 0;
        off = bm_bytes;
        // per entry: NW accumulator words + one list word (+ one key word of the hash attempt)
        const u32 per_entry = (NW + 1 + (ctx->block_tuning.hash ? 1u : 0u)) * 4;
        u64 cap = ((budget - off - 48) / per_entry) & ~31ull;
        if (!ctx->block_tuning.hash) cap = std::min<u64>(cap, (span + 31) & ~31ull);
        if (ctx->zone_tuning.force_cap) cap = std::max<u64>(1, std::min<u64>(cap, ctx->zone_tuning.force_cap));
        bp.cap = (u32)cap;
        bp.off_list = (u32)off;
        off += ((size_t)bp.cap * 4 + 15) & ~15ull;
        bp.off_keys = (u32)off;
        if (ctx->block_tuning.hash) off += ((size_t)bp.cap * 4 + 15) & ~15ull;
        bp.off_acc = (u32)off;
        off += (size_t)bp.cap * NW * 4;
    }
    bp.smem = off;
    return off <= budget;
}

// Returns the product restricted to [lo, hi) (both multiples of S), or nullptr when the operands do not suit the
// block path (tiny groups); the caller then uses the generic zone path.  parts = number of output partitions the
// whole product is cut into (this call computes one of them).
inline pk_dpoly *block_multiply(pk_ctx *ctx, const Plan &pl, BlockPlan bp, const pk_dpoly *A, const pk_dpoly *B,
                                const u64 *keyA, const u64 *keyB, u64 lo, u64 hi, u32 nv, const i64 *degA, const i64 *degB,
                                i64 max_degree, const uint4 *prepackA, const uint4 *prepackB, u32 parts)
{
    const u64 nA = A->n, nB = B->n;
    const bool trunc = degA != nullptr;
    const bool timing = getenv("PK_TIMING") != nullptr; // debug: host wall time of every phase (adds stream syncs)
    auto t_prev = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pk block] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    if (trunc) { // degree offsets are 31-bit: bound the degree span of each operand by its exponent spans
        for (const pk_dpoly *P : {A, B}) {
            unsigned __int128 span = 0;
            for (u32 v = 0; v < nv; ++v)
                if (pl.cpA.mask[v]) span += (unsigned __int128)((__int128)P->emax[v] - P->emin[v]);
            if (span >= ((unsigned __int128)1 << 31)) return nullptr;
        }
    }
    const BlockTuning &bt = ctx->block_tuning;
    const u32 S = bp.S;
    if (lo % S != 0 || (hi % S != 0 && hi != pl.range) || hi <= lo) return nullptr;
    const u32 h_base = (u32)(lo / S);
    const u32 n_h = (u32)((hi - lo + S - 1) / S);
    const unsigned __int128 w_bound = (unsigned __int128)nA * nB;
    const u64 out_cap = (u64)std::min<unsigned __int128>(w_bound, hi - lo);
    if ((unsigned __int128)out_cap * (9 + 8 * pl.out_limbs) > (unsigned __int128)ctx->total_mem / 2) return nullptr;

    const u64 HR = pl.range / S + 2;
    // ---- zone geometry: about zone_work products per zone, at least 4 zones per thread block when possible
    // (an output partition holds about 1 / parts of the products: the cuts are quantiles of sampled pair sums)
    const double avg_h = (double)w_bound / (double)std::max<u32>(1, parts) / (double)n_h;
    u64 Hz = (u64)std::max(1.0, (double)bt.zone_work / std::max(1.0, avg_h));
    Hz = std::min<u64>(Hz, std::max<u64>(1, n_h / (4ull * bp.grid)));
    if (bp.mode == 1) { // bitmap zones: keep the products of a zone near what one pass of entries holds (measured: pearce1)
        const double cap_est = (double)(block_smem_budget(ctx) / ((bp.NW + 1) * 4));
        Hz = std::min<u64>(Hz, (u64)std::max(1.0, cap_est / std::max(1.0, avg_h)));
    }
    Hz = std::max<u64>(1, std::min<u64>(Hz, bp.Hz_max));
    if (bt.force_hz) Hz = std::max<u64>(1, std::min<u64>(bt.force_hz, bp.Hz_max));
    while (!block_layout(ctx, bp, (u32)Hz)) {
        if (Hz == 1) return nullptr;
        Hz = (Hz + 1) / 2;
    }
    // bitmap zones are cut at equal work (k_block_equal_zones): nq quantiles of at most zone_work products each, at least
    // 4 per thread block, no zone longer than Hz blocks; dense zones are Hz blocks each.  nz = bound on the zone count.
    // (measured on pearce1: zones of ~one pass of entries (14 k products) are 20 % SLOWER than zones of 65 k -- a walk
    // hands out one 512-product tile per warp and round, so a zone needs several rounds to keep 32 warps busy.)
    // By default only the output partitions of a multi-GPU product use them: a partition starts or ends in the dense
    // middle of the code range, where zones of Hz blocks are several times heavier than the average zone and the
    // heaviest-first schedule cannot balance them (pearce1, halves: 0.228 / 0.314 ms with zones of Hz blocks,
    // 0.210 / 0.232 ms cut at equal work; whole products are within 1 % (pearce1) or 5 % slower (pearce2) and keep Hz).
    const bool equal_zones = bp.mode == 1 && (bt.equal_zones == 2 || (bt.equal_zones == 1 && parts > 1));
    u32 nq = 0;
    if (equal_zones) {
        const double per_zone = bt.zone_products ? (double)bt.zone_products : (double)bt.zone_work;
        const double want = (double)w_bound / (double)std::max<u32>(1, parts) / std::max(1.0, per_zone);
        nq = (u32)std::min<double>(1u << 16, std::max<double>(4.0 * bp.grid, want));
    }
    const u32 nz = equal_zones ? nq + (u32)(n_h / Hz) + 2 : (u32)((n_h + Hz - 1) / Hz);
    // every array that must start at zero lives in ONE allocation cleared by ONE memset: group tables, group counters,
    // zone bookkeeping, the tile counters of the count pass and the tile cursors of the scatter pass
    const size_t z_tab = (size_t)HR * 8, z_book = (size_t)nz * 36 + 160, z_cnt = (((size_t)n_h + 1) * 4 + 15) & ~15ull;
    const size_t zo_tabA = 0, zo_tabB = z_tab, zo_gcnt = 2 * z_tab, zo_book = zo_gcnt + 16, zo_cnt = zo_book + z_book,
                 zo_cur = zo_cnt + z_cnt, z_total = zo_cur + z_cnt;
    DevBuf zeroed(ctx, z_total);
    CUDA_CHECK(cudaMemsetAsync(zeroed.p, 0, z_total, ctx->stream));
    char *const zbase = zeroed.as<char>();
    uint2 *const tabA_p = reinterpret_cast<uint2 *>(zbase + zo_tabA), *const tabB_p = reinterpret_cast<uint2 *>(zbase + zo_tabB);
    u32 *const gcnt_p = reinterpret_cast<u32 *>(zbase + zo_gcnt);
    u32 *const cnt_p = reinterpret_cast<u32 *>(zbase + zo_cnt), *const cur_p = reinterpret_cast<u32 *>(zbase + zo_cur);
    const bool packed = prepackA && prepackB && !trunc;
    DevBuf pa_own(ctx, packed ? 0 : (nA + 1) * 16), pb_own(ctx, packed ? 0 : (nB + 1) * 16), tdev(ctx, trunc ? sizeof(TruncDev) : 0);
    uint4 *const pa_p = packed ? const_cast<uint4 *>(prepackA) : pa_own.as<uint4>();
    uint4 *const pb_p = packed ? const_cast<uint4 *>(prepackB) : pb_own.as<uint4>();
    DevBuf gdegA(ctx, trunc ? HR * 8 : 0), gdegB(ctx, trunc ? HR * 8 : 0);
    TruncDev *td = tdev.as<TruncDev>();
    if (trunc) {
        PK_LAUNCH(ctx, k_trunc_init, 1, 1, 0, td);
        const u32 rg = std::min<u32>(grid_for(std::max(nA, nB), 256), (u32)ctx->sm_count * 4);
        PK_LAUNCH(ctx, k_deg_minmax, rg, 256, 0, degA, (u32)nA, &td->min_a, &td->max_a);
        PK_LAUNCH(ctx, k_deg_minmax, rg, 256, 0, degB, (u32)nB, &td->min_b, &td->max_b);
        PK_LAUNCH(ctx, k_trunc_limits, 1, 1, 0, td, max_degree, ctx->d_ctr);
        PK_LAUNCH(ctx, k_pack_operand_deg, grid_for(nA, 256), 256, 0, keyA, A->planes, degA, &td->min_a, pa_p, nA);
        PK_LAUNCH(ctx, k_pack_operand_deg, grid_for(nB, 256), 256, 0, keyB, B->planes, degB, &td->min_b, pb_p, nB);
        CUDA_CHECK(cudaMemsetAsync(gdegA.p, 0xFF, HR * 8, ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(gdegB.p, 0xFF, HR * 8, ctx->stream));
    } else if (!packed) {
        PK_LAUNCH(ctx, k_pack_operand, grid_for(nA, 256), 256, 0, keyA, A->planes, pa_p, nA);
        PK_LAUNCH(ctx, k_pack_operand, grid_for(nB, 256), 256, 0, keyB, B->planes, pb_p, nB);
    }

    phase("pack");
    // ---- groups of equal hi
    const u64 magicS = (u64)((((unsigned __int128)1) << 64) / S) + 1;
    DevBuf glA(ctx, nA * 4), glB(ctx, nB * 4);
    PK_LAUNCH(ctx, k_block_groups, grid_for(nA, 256), 256, 0, pa_p, (u32)nA, magicS, tabA_p, glA.as<u32>(),
              gcnt_p, trunc ? gdegA.as<uint2>() : nullptr);
    PK_LAUNCH(ctx, k_block_groups, grid_for(nB, 256), 256, 0, pb_p, (u32)nB, magicS, tabB_p, glB.as<u32>(),
              gcnt_p + 1, trunc ? gdegB.as<uint2>() : nullptr);
    // lanes of a tile on distinct banks: re-order the terms inside every group of B (k_block_permute)
    const bool permute = bt.permute != 0;
    DevBuf pb_perm(ctx, permute ? (nB + 1) * 16 : 0), hi_perm(ctx, permute && bp.LW == 2 ? (nB + 1) * 8 : 0), rank_tmp(ctx, permute ? nB * 4 : 0);
    DevBuf zero1(ctx, bp.LW == 2 ? std::max(nA, nB) * 8 + 16 : 0);
    const uint2 *hiA = nullptr, *hiB = nullptr;
    if (bp.LW == 2) {
        // second limbs: a limb plane of u64 is bit-identical to the {lo, hi} word pairs the kernel loads; an operand that
        // needs only one limb gets a zero plane
        if (pl.LA < 2 || pl.LB < 2) CUDA_CHECK(cudaMemsetAsync(zero1.p, 0, std::max(nA, nB) * 8 + 16, ctx->stream));
        hiA = (pl.LA == 2 && A->limbs >= 2) ? reinterpret_cast<const uint2 *>(A->planes + A->stride) : zero1.as<uint2>();
        hiB = (pl.LB == 2 && B->limbs >= 2) ? reinterpret_cast<const uint2 *>(B->planes + B->stride) : zero1.as<uint2>();
    }
    const uint4 *pb_use = pb_p;
    if (permute) {
        const u32 pgrid = std::min<u32>(grid_for(nB, 64), (u32)ctx->sm_count * 8);
        PK_LAUNCH(ctx, k_block_permute, pgrid, 256, 0, pb_p, hiB, tabB_p, glB.as<u32>(), gcnt_p + 1,
                  rank_tmp.as<u32>(), pb_perm.as<uint4>(), hi_perm.as<uint2>());
        pb_use = pb_perm.as<uint4>();
        if (bp.LW == 2) hiB = hi_perm.as<uint2>();
    }
    u32 *hg = reinterpret_cast<u32 *>(static_cast<char *>(ctx->h_scratch) + 2048);
    readback(ctx, hg, gcnt_p, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const u32 GA = hg[0], GB = hg[1];
    phase("groups");
    const u64 pairs = (u64)GA * GB;
    // tiny groups: the tile list would be as long as the product itself
    if ((unsigned __int128)pairs * bt.min_tile > w_bound || pairs >= (1ull << 31)) return nullptr;

    if (getenv("PK_BLOCK_DEBUG"))
        fprintf(stderr, "block path: S=%u split=%u n_h=%u Hz=%llu (max %u) zones<=%u mode=%u NW=%u PW=%u cap=%u bm_words=%u smem=%zu GA=%u GB=%u\n",
                S, bp.split, n_h, (unsigned long long)Hz, bp.Hz_max, nz, bp.mode, bp.NW, bp.PW, bp.cap, bp.bm_words, bp.smem, GA, GB);

    // ---- tile list: count -> scan -> scatter
    // every chunk of gb (at most nb / 32 + 2 per pair) gives at most 2 * na * width / target + 1 tiles (+ na / 1023 when the
    // row count per tile hits its 10-bit cap)
    const u64 max_tiles64 = 2 * (u64)(w_bound / std::min<u32>(bt.tile_target, 512)) + ((u64)GA * nB) / kTileCols + 2 * pairs
                            + (nB / kTileCols + 2 * (u64)GB) * (nA / 1023 + 1) + 64;
    if (max_tiles64 >= (1ull << 31)) return nullptr;
    const u32 max_tiles = (u32)max_tiles64;
    DevBuf tstart(ctx, ((size_t)n_h + 1) * 4), tiles(ctx, (size_t)max_tiles * 16);
    // bookkeeping: (spare u64), offsets (u64), prefix (u64) per zone; bump; counters; zone count + order (u32)
    u64 *zwork = reinterpret_cast<u64 *>(zbase + zo_book);
    u64 *zone_off = zwork + nz;
    u64 *zone_prefix = zone_off + nz;
    u64 *bump = zone_prefix + nz;
    u32 *chunk_counter = reinterpret_cast<u32 *>(bump + 1);
    u32 *n_zones_dev = chunk_counter + 1;
    u32 *n_active_dev = chunk_counter + 2;
    u32 *zone_count = chunk_counter + 4;
    u32 *order = zone_count + nz;
    u32 *zstart = order + nz; // nz + 1 entries
    TileArgs ta{};
    ta.glistA = glA.as<u32>();
    ta.glistB = glB.as<u32>();
    ta.tabA = tabA_p;
    ta.tabB = tabB_p;
    ta.GA = GA;
    ta.GB = GB;
    ta.h_base = h_base;
    ta.n_h = n_h;
    // (measured: dense zones like 1024-2048 products per tile, bitmap zones 512 -- their tiles are walked twice and the
    // zones are smaller, so finer tiles balance the warps better)
    ta.target = bp.mode == 1 ? std::min<u32>(bt.tile_target, 512) : bt.tile_target;
    ta.wide = bt.wide;
    ta.gdegA = trunc ? gdegA.as<uint2>() : nullptr;
    ta.gdegB = trunc ? gdegB.as<uint2>() : nullptr;
    ta.trunc = trunc ? td : nullptr;
    ta.cnt = cnt_p;
    ta.tstart = tstart.as<u32>();
    ta.tiles = tiles.as<uint4>();
    ta.max_tiles = max_tiles;
    ta.ctr = ctx->d_ctr;
    const u32 pgrid = grid_for(pairs, 256);
    PK_LAUNCH(ctx, (k_block_tiles<0>), pgrid, 256, 0, ta);
    {
        size_t tmp_bytes = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt_p, tstart.as<u32>(), (int)(n_h + 1), ctx->stream));
        DevBuf tmp(ctx, tmp_bytes);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt_p, tstart.as<u32>(), (int)(n_h + 1), ctx->stream));
    }
    ta.cnt = cur_p; // the scatter pass counts again, into its own zeroed cursors
    PK_LAUNCH(ctx, (k_block_tiles<1>), pgrid, 256, 0, ta);
    const bool lpt = !bt.natural && nz <= (1u << 17);
    if (equal_zones) PK_LAUNCH(ctx, k_block_equal_zones, 1, 1024, 0, tstart.as<u32>(), n_h, (u32)Hz, nq, nz, zstart, n_zones_dev);
    else if (!lpt) PK_LAUNCH(ctx, k_block_uniform_zones, grid_for((u64)nz + 1, 256), 256, 0, n_h, (u32)Hz, nz, zstart, n_zones_dev);
    if (lpt)
        PK_LAUNCH(ctx, k_block_order, 1, 1024, 0, tstart.as<u32>(), zstart, n_zones_dev, order, n_active_dev, n_h,
                  equal_zones ? 0u : (u32)Hz, equal_zones ? 0u : nz);
    else PK_LAUNCH(ctx, k_block_natural_order, 1, 1, 0, n_zones_dev, n_active_dev);

    phase("tile list");
    const StagingView g = staging_acquire(ctx, out_cap, nv, pl.out_limbs);
    phase("staging");
    BlockArgs ba{};
    ZoneArgs &za = ba.z;
    za.A1 = hiA;
    za.B1 = hiB;
    za.A = pa_p;
    za.nA = (u32)nA;
    za.B = pb_use;
    za.nB = (u32)nB;
    za.lo = (u32)lo;
    za.hi = (u32)hi;
    za.cap = bp.cap;
    za.pw = std::min<u32>(4, (pl.prod_bits + 31) / 32);
    za.bm_words = bp.bm_words;
    za.off_acc = bp.off_acc;
    za.off_bm = bp.off_bm;
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
    ba.zstart = zstart;
    ba.S = S;
    ba.trunc = trunc ? td : nullptr;
    ba.off_keys = bp.off_keys;
    ba.hash = (bp.mode == 1 && bt.hash) ? 1u : 0u;
    ba.pair_walk = bt.pair_walk;
    const bool is_signed = A->any_neg || B->any_neg;
    const u32 grid = std::min<u32>(bp.grid, nz);
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul0, ctx->stream));
#define X(nw, pw) \
    if (bp.LW == 1 && bp.NW == nw && bp.PW == pw) block_launch_w<nw, pw>(ctx, ba, grid, bp.smem, bp.mode, is_signed);
    PK_BLOCK_WIDTHS(X)
#undef X
#define X(nw, pw) \
    if (bp.LW == 2 && bp.NW == nw && bp.PW == pw) block_launch_w2<nw, pw>(ctx, ba, grid, bp.smem, bp.mode);
    PK_BLOCK_WIDTHS2(X)
#undef X
    CUDA_CHECK(cudaEventRecord(ctx->ev_mul1, ctx->stream));
    phase("k_block");
    PK_LAUNCH(ctx, k_zone_scan, 1, 1024, 0, zone_count, zone_prefix, n_zones_dev, &ctx->d_ctr->n_out);
    MulCounters *hc = static_cast<MulCounters *>(ctx->h_scratch);
    readback(ctx, hc, ctx->d_ctr, sizeof(MulCounters));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const u64 n_out = hc->n_out;
    if (getenv("PK_BLOCK_DEBUG")) fprintf(stderr, "block path: %llu of %u zones fell back from the hash attempt\n", (unsigned long long)hc->aux, nz);
    if (n_out > out_cap || (hc->flags & kFlagTableFull)) fail(PK_E_CUDA, "block path: a count exceeds its bound (internal error)");
    ctx->stats.table_slots = (u64)bp.cap;
    ctx->stats.acc_lanes = bp.NW;
    ctx->stats.chunk_bits = bp.PW; // (block path: 32-bit words of a product = unconditional ATOMS per term product)
    ctx->stats.retries = bp.mode | (1024u << 8);
    phase("scan + count");
    DPolyGuard e{ctx, new_dpoly(ctx, n_out, nv, pl.out_limbs)};
    phase("result alloc");
    if (n_out) {
        PK_LAUNCH(ctx, k_zone_gather, grid_for(n_out, kGatherPerBlock), 256, 0, zone_count, zone_off, zone_prefix, n_zones_dev,
                  g.p->key, g.p->planes, g.p->stride, g.p->sign, e.p->key, e.p->planes, e.p->stride, e.p->sign, pl.out_limbs, n_out);
    }
    phase("gather");
    return e.release();
}

} // namespace pk
