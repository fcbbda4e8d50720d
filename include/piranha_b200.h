/*
 * piranha_b200.h -- C ABI of the B200-native sparse polynomial multiplier.
 *
 * This is the drop-in boundary for ONE path of bluescarni/piranha:
 *     series_multiplier<polynomial<Cf, kronecker_monomial<>>>  and the base_series_multiplier machinery.
 * The reference has no FFI for this path: the plugin mechanism is the class-template protocol of
 * series_multiplier (include/piranha/series_multiplier.hpp:37-56 -- construct from two series, then
 * operator()() returns the product).  A specialisation of that template marshals the operands' terms to the
 * structure-of-arrays form below, calls pk_mul(), and re-inserts the returned terms with
 * rehash/_unique_insert/_update_size (INTEGRATION.md shows the binding).  Every entry point cites the
 * reference interface it replaces; paths are relative to /root/reference/include/piranha/.
 *
 * Plain pointers and sizes only; no C++ or torch types.  Re-entrant: all mutable state lives in a pk_ctx
 * (one CUDA device + one stream, or -- pk_ctx_create_multi -- several devices driven from the one calling thread);
 * different contexts may be used concurrently from different threads (base_series_multiplier.hpp:404-408: multipliers
 * are built and consumed on one thread).  No call changes the calling thread's current CUDA device, and device memory
 * comes from a pool private to the context (the device's default memory pool is not touched).
 *
 * Term format (both directions):
 *   key[i]   int64 Kronecker code of the exponent vector, piranha's kronecker_array<int64_t> codification for
 *            n_vars components (kronecker_array.hpp:274-307).  The per-component limits M_k the codification
 *            uses are taken from the library's own restatement of kronecker_array.hpp:128-229, or from
 *            pk_ctx_set_limits() when the caller's table differs (it depends on the C++ standard library).
 *   mag[i*limbs .. ]  magnitude of the integer coefficient, `limbs` 64-bit limbs, least significant first
 *            (the layout of mppp::integer's static storage / GMP's mpz limbs).
 *   sign[i]  +1 or -1 (0 is accepted for a zero coefficient and contributes nothing).
 * Results are returned sorted by key (canonical order), with unique keys and no zero coefficients
 * (series invariants, series.hpp:1312-1320; sanitise_series, base_series_multiplier.hpp:855-949).
 */
#ifndef PIRANHA_B200_H
#define PIRANHA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes and the exception the reference would have thrown (polynomial.hpp:1283, :1315;
 * base_series_multiplier.hpp:356, :843-846). */
#define PK_OK 0
#define PK_E_ARGS 1      /* std::invalid_argument: mismatched n_vars, bad sizes, bad options */
#define PK_E_KEY_RANGE 2 /* std::overflow_error "Kronecker monomial components are out of bounds"
                            (check_bounds, polynomial.hpp:1118-1194) */
#define PK_E_CF_WIDTH 3  /* coefficient wider than the library computes: operands of more than 4 limbs (256 bits) or a result
                            bound of more than 9 limbs.  Never a silent wrap and never a CPU computation (the reference promotes
                            to heap mpz instead, integer.hpp:67-85).  Width protocol: operands of 1-2 limbs run on the pairing
                            kernels (results up to 5 limbs); operands of 3-4 limbs are multiplied as 128-bit halves through the
                            same kernels and summed with shifts on the device (benchmarks/fateman1_dynamic.cpp chains) */
#define PK_E_NOMEM 4     /* std::bad_alloc */
#define PK_E_CUDA 5      /* std::runtime_error: CUDA failure, or no sm_100 device present */
#define PK_E_DEGREE 6    /* std::overflow_error from the checked degree subtraction of the truncation logic
                            (degree_sub -> safe_int_sub, polynomial.hpp:1243-1252, :1504) */

#define PK_E_UNSUPPORTED 7 /* std::runtime_error: a product the reference would compute but this library does not handle -- exponent
                            spans whose re-encoded code range reaches 2^62 (e.g. 1 + x + x^(2^62)), operands of 2^32 terms or
                            more.  Distinct from PK_E_KEY_RANGE: the Kronecker bounds hold, nothing overflowed */

#define PK_MAX_VARS 40u  /* kronecker_array<int64_t> supports < 40 components */

/* Accumulation strategy (pk_opts.algo). */
#define PK_ALGO_AUTO 0u
#define PK_ALGO_DENSE 1u /* dense accumulator in HBM indexed by the re-encoded (tight mixed-radix) code */
#define PK_ALGO_HASH 2u  /* open-addressing hash table in HBM (replaces hash_set buckets, hash_set.hpp:1251-1327) */
#define PK_ALGO_ZONE 3u  /* output range cut into zones, each accumulated in one SM's shared memory
                            (the GPU form of the zone scheme of polynomial.hpp:1783-1966) */
#define PK_ALGO_BLOCK 4u /* zones aligned to the digits of the re-encoded code: the double loop becomes a list of dense
                            tiles (blocked_multiplication, base_series_multiplier.hpp:449-504, inside the zone scheme) */

typedef struct pk_ctx pk_ctx;     /* one device + one stream + cached workspaces */
typedef struct pk_dpoly pk_dpoly; /* a polynomial resident in HBM (sorted by key, with exponent/width metadata) */

/* Host view of a polynomial: what fill_term_pointers (base_series_multiplier.hpp:89-154) walks, flattened. */
typedef struct pk_poly {
    uint64_t n;          /* number of terms (series::size()) */
    uint32_t n_vars;     /* size of the symbol set (m_ss.size()) */
    uint32_t limbs;      /* 64-bit limbs per coefficient magnitude, 1..4 for operands (results: up to 9) */
    const int64_t *key;  /* n Kronecker codes, any order (hash_set order is unspecified) */
    const uint64_t *mag; /* n * limbs */
    const int8_t *sign;  /* n */
} pk_poly;

/* Library-owned result (pinned host memory); release with pk_result_free(). */
typedef struct pk_result {
    uint64_t n;
    uint32_t n_vars;
    uint32_t limbs;
    int64_t *key;
    uint64_t *mag;
    int8_t *sign;
    void *owner;
} pk_result;

/* Options of one multiplication.  Zero-initialise, set struct_size = sizeof(pk_opts). */
typedef struct pk_opts {
    uint32_t struct_size;
    /* Truncation (polynomial::get_auto_truncate_degree(), polynomial.hpp:736-740, :1555-1573):
     * 0 = none (untruncated_kronecker_mult), 1 = total degree, 2 = partial degree over degree_mask.
     * A pair (i, j) is multiplied iff deg(t1_i) + deg(t2_j) <= max_degree (polynomial.hpp:1495-1509). */
    int32_t trunc_mode;
    int64_t max_degree;
    const uint8_t *degree_mask; /* n_vars bytes, non-zero = variable counts (mode 2); ignored otherwise */
    /* Output partition: this call computes only the result terms whose re-encoded code lies in part
     * `part_index` of `part_count` disjoint, ordered ranges (cut so that pair counts are balanced; the cuts are
     * a deterministic function of the operands, so every rank derives the same ones).  0/0 or 0/1 = everything. */
    uint32_t part_index;
    uint32_t part_count;
    uint32_t algo;     /* PK_ALGO_* */
    uint32_t reserved; /* must be 0 */
} pk_opts;

/* Counters of the most recent pk_mul()/pk_mul_dev() on a context, for benchmarks and tests. */
typedef struct pk_stats {
    uint64_t n1, n2;              /* operand sizes after the larger-first swap */
    uint64_t term_products;       /* W: pairs (i, j) actually multiplied in this call's partition */
    uint64_t result_terms;        /* R of this call's partition */
    uint64_t code_range;          /* number of codes in the re-encoded output range */
    uint64_t table_slots;         /* accumulator slots allocated (dense range, hash capacity, or zone capacity) */
    uint32_t algo;                /* PK_ALGO_* actually used */
    uint32_t acc_lanes;           /* 64-bit lanes (dense/hash) or 32-bit words (zone) per accumulator */
    uint32_t chunk_bits;          /* dense/hash: payload bits per lane (deferred-carry radix); block path: 32-bit words
                                     of a product (= shared-memory atomics per term product); 0 for the zone path */
    uint32_t result_limbs;        /* limbs per result coefficient */
    uint32_t kernel_launches;     /* kernels launched by this call */
    uint32_t retries;             /* hash-capacity / zone-split retries */
    uint64_t total_launches;      /* kernels launched on this context since creation */
    float mul_kernel_ms;          /* CUDA-event time of the dominant (pairing + accumulate) kernel */
    float device_ms;              /* CUDA-event time of all device work of the call */
} pk_stats;

/* ---- context ---------------------------------------------------------------------------------------------- */

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t to run on (e.g. the caller's current
 * stream so its CUDA events see the work) or NULL for a private non-blocking stream.
 * Fails with PK_E_CUDA when no usable GPU is present: there is no CPU fallback. */
int pk_ctx_create(int device, void *stream, pk_ctx **out);
/* Create a context that spreads every product over `n_devices` <= 16 GPUs of this box (device_ids = NULL: devices
 * 0..n-1; a device may be listed more than once, its parts then share it -- how the tests run the N-part path on one GPU).
 * This is the multiplier's own parallelism -- the zones that the reference's thread pool claims inside ONE call of
 * series_multiplier::operator()() (polynomial.hpp:1783-1966, settings::set_n_threads) -- lifted to devices: the output
 * Kronecker-code range is cut into n_devices ordered parts balanced by work, device r computes part r from its own
 * replica of the operands, and pk_mul() / pk_download() copy the parts over n_devices PCIe links at once, each straight
 * to its offset of ONE pinned host result (ordered ranges: no merge, no device-to-device traffic).  Products of
 * pk_mul_dev() stay partitioned in HBM and are replicated peer-to-peer only when they become operands.  One worker thread
 * per device; the handle is used from one thread at a time like any other context.  All entry points accept the handle
 * except the single-device exports (pk_dpoly_export*, pk_dpoly_publish_size, pk_dpoly_export_peers) and pk_opts.part_count
 * > 1, which return PK_E_ARGS; polynomials belong to the context that created them. */
int pk_ctx_create_multi(uint32_t n_devices, const int *device_ids, pk_ctx **out);
/* Number of usable (sm_100) CUDA devices on this box; 0 when there is none (no context can be created then). */
int pk_device_count(void);
/* Block until everything the context has queued on its stream(s) has finished (cudaStreamSynchronize).  A caller that
 * mixes its own streams with a context on a private stream orders the two with this. */
int pk_ctx_synchronize(pk_ctx *ctx);
/* Number of devices behind the handle (1 for pk_ctx_create). */
uint32_t pk_ctx_device_count(const pk_ctx *ctx);
void pk_ctx_destroy(pk_ctx *ctx);
/* Release what the context keeps between products to make them cheap: the staging arrays of the shared-memory paths
 * (up to min(term products, code range) terms, ~1 GB after a pearce1-sized product), idle pinned result buffers and the
 * cached blocks of the device's stream-ordered memory pool.  The next product re-creates what it needs. */
int pk_ctx_trim(pk_ctx *ctx);
/* Message of the last error on this context (never NULL). */
const char *pk_last_error(const pk_ctx *ctx);
/* Override the Kronecker limits M_0..M_{n_vars-1} for `n_vars` components with the caller's
 * kronecker_array<int64_t>::get_limits()[n_vars] (kronecker_array.hpp:251-254). */
int pk_ctx_set_limits(pk_ctx *ctx, uint32_t n_vars, const int64_t *limits);
/* Read the limits the context uses for `n_vars` components; returns PK_E_ARGS if n_vars is unsupported. */
int pk_ctx_get_limits(const pk_ctx *ctx, uint32_t n_vars, int64_t *limits_out, int64_t *h_min_out, int64_t *h_max_out);
int pk_get_stats(pk_ctx *ctx, pk_stats *out);
/* Kernel tuning knobs (the analogue of piranha::tuning, tuning.hpp:54-60).  Never change results, only how they are
 * computed.  Keys:
 *   zone path   "zone_target" (zones per product, 0 = 12 per SM), "zone_chunk" (consecutive zones per work unit),
 *               "zone_span_log2" (max log2 codes per bitmap zone), "zone_cap" (force a small accumulator capacity; tests),
 *               "zone_mode" (0 auto, 1 dense, 2 bitmap), "zone_disable", "zone_cursor_bytes" (memory for row cursors,
 *               0 = binary search per row per zone), "zone_tpb" (accepted, no effect: 1024 threads per block), "zone_uniform" (1 =
 *               equal-span zones)
 *   block path  "block_disable", "block_target" (products per tile), "block_zone_work" (products per zone aimed at),
 *               "block_min_tile" (decline when the average pair of groups is smaller), "block_mode" (0 auto, 1 dense,
 *               2 bitmap), "block_hz" (blocks per zone), "block_natural" (1 = zones claimed in code order instead of
 *               heaviest first), "block_equal_zones" (bitmap mode, zones cut at equal work instead of block_hz blocks each: 0 never,
 *               1 = for output partitions only, the default, 2 always), "block_zone_products" (products per equal-work zone, 0 = auto), "block_wide" (64-term tiles), "block_pair_walk" (1 = the accumulating walk takes two products at a time, default), "block_smem" (bytes of shared memory per zone, 0 = all)
 *               and the measured-and-rejected experiments, all off: "block_permute" (bank-conflict-free order inside
 *               groups), "block_hash" (one-walk hash accumulation for sparse zones), "block_radix0" (pad the radix of
 *               digit 0 to a residue mod 32)
 *   pk_mul      "pipe_parts" (output parts of a pipelined product whose device-to-host copy overlaps the computation of
 *               the next part; 0 = off, the default), "pipe_min_products", "pipe_test_small" (tests) */
int pk_ctx_set_tuning(pk_ctx *ctx, const char *key, uint64_t value);

/* ---- host-to-host product: the call behind series_multiplier::operator()() --------------------------------- */

/* out = a * b.  Replaces series_multiplier<polynomial<Cf, kronecker_monomial<>>>(a, b)()
 * (polynomial.hpp:1291-1298 ctor incl. check_bounds, :1332-1336 operator(), :1619-1628 execute,
 * :1632-1672 untruncated_kronecker_mult, :1402-1449 _truncated_multiplication).
 * Strong guarantee: on failure *out is zeroed (polynomial.hpp:1961-1965). */
int pk_mul(pk_ctx *ctx, const pk_poly *a, const pk_poly *b, const pk_opts *opts, pk_result *out);
void pk_result_free(pk_ctx *ctx, pk_result *r);

/* ---- device-resident operands: chained products (f *= tmp, pow) without re-uploading ------------------------ */

/* Copy a host polynomial to HBM, sort it by key and record its per-variable exponent bounds and coefficient
 * width (the work of fill_term_pointers + the operand half of check_bounds). */
int pk_upload(pk_ctx *ctx, const pk_poly *p, pk_dpoly **out);
/* out = a * b with everything staying in HBM (same semantics and errors as pk_mul). */
int pk_mul_dev(pk_ctx *ctx, const pk_dpoly *a, const pk_dpoly *b, const pk_opts *opts, pk_dpoly **out);
/* The bounds check of the multiplier's constructor alone (check_bounds, polynomial.hpp:1118-1194, :1291-1298): PK_OK, or
 * PK_E_KEY_RANGE when an exponent of a * b could leave the Kronecker limits, PK_E_ARGS for mismatching symbol sets.
 * Host arithmetic on the exponent bounds recorded at upload; nothing is computed on the device. */
int pk_check_bounds(pk_ctx *ctx, const pk_dpoly *a, const pk_dpoly *b);
/* Copy a device polynomial back into library-owned pinned host memory. */
int pk_download(pk_ctx *ctx, const pk_dpoly *p, pk_result *out);
/* The same terms, grouped for a hash table: ordered by ((uint64_t)key >> shift) & (2^bits - 1), terms of equal group in
 * ascending key order (a stable device-side sort, 1 <= bits <= 16).  With the identity hash of the reference's packed
 * monomial (kronecker_monomial::hash, kronecker_monomial.hpp:441-444) and a power-of-two bucket count 2^b
 * (hash_set.hpp:1298-1317: bucket = hash mod bucket count), shift = b - bits groups the terms by the TOP bits of their
 * bucket, so a fill of the container (rehash + _unique_insert, polynomial.hpp:1666-1667) walks the bucket array window by
 * window instead of at random.  PK_E_UNSUPPORTED on a multi-device context (use pk_download). */
int pk_download_grouped(pk_ctx *ctx, const pk_dpoly *p, uint32_t shift, uint32_t bits, pk_result *out);
/* pk_download_grouped in n_chunks pieces (1 .. 256, in term order), so that the caller can consume the first terms while
 * the rest is still crossing PCIe: after every piece fn(user, out, terms_ready) runs on the CALLING thread, and
 * out->key / mag / sign [0, terms_ready) are final (n, n_vars, limbs and the array addresses are final from the first
 * call on; a result of zero terms makes no call).  fn must not throw and must not call into this context.
 * group_offsets: NULL, or room for 2^bits + 1 entries, filled BEFORE the first call of fn: group g is the terms
 * [group_offsets[g], group_offsets[g + 1]) -- with shift = b - bits as above, the terms of one group fall into a range of
 * buckets that no other group touches, so a thread that fills whole groups needs no atomics on the bucket array. */
typedef void (*pk_progress_fn)(void *user, const pk_result *out, uint64_t terms_ready);
int pk_download_grouped_progress(pk_ctx *ctx, const pk_dpoly *p, uint32_t shift, uint32_t bits, uint32_t n_chunks,
                                 pk_progress_fn fn, void *user, uint64_t *group_offsets, pk_result *out);
void pk_dpoly_free(pk_ctx *ctx, pk_dpoly *p);
uint64_t pk_dpoly_size(const pk_dpoly *p);
uint32_t pk_dpoly_limbs(const pk_dpoly *p);
uint32_t pk_dpoly_nvars(const pk_dpoly *p);
/* Copy the terms into caller-provided DEVICE buffers (key: n int64, mag: n*limbs uint64 row-major, sign: n int8),
 * e.g. torch tensors about to be all-gathered over NCCL.  Any pointer may be NULL to skip that array. */
int pk_dpoly_export(pk_ctx *ctx, const pk_dpoly *p, void *d_key, void *d_mag, void *d_sign);
/* Same, with magnitude rows of mag_limbs >= pk_dpoly_limbs(p) words (high limbs zero): lets every rank of a
 * partitioned product write its terms straight into its slice of one common result buffer. */
int pk_dpoly_export_padded(pk_ctx *ctx, const pk_dpoly *p, void *d_key, void *d_mag, void *d_sign, uint32_t mag_limbs);
/* Term exchange of a partitioned product over peer-mapped device memory (NVLink / NVSwitch; e.g. the buffers of a torch
 * symmetric-memory allocation), the "gather result sizes and terms" step of the multi-GPU path without a collective
 * library on the data path.  All pointer tables are host arrays of n_peers <= 16 DEVICE addresses, valid on this
 * context's GPU, index = rank (own rank included).
 *   pk_dpoly_publish_size: writes {term count, limbs} of p to slot_ptrs[j] (2 x uint64 each) on every peer j.
 *   pk_dpoly_export_peers: after a barrier across the ranks, d_sizes (this rank's own table: n_peers x {count, limbs}) is
 *     complete; the kernel derives this rank's term offset (sum of the counts of lower ranks) and the common row width
 *     (max limbs) from it and stores its terms into key_ptrs[j] / mag_ptrs[j] / sign_ptrs[j] of every peer j.  Nothing is
 *     written when the whole result exceeds cap_terms terms or cap_limbs <= 8 limbs: every rank sees the same table, so
 *     every rank skips; the caller grows the buffers and repeats.  A second barrier completes the exchange. */
int pk_dpoly_publish_size(pk_ctx *ctx, const pk_dpoly *p, uint32_t n_peers, const uint64_t *slot_ptrs);
int pk_dpoly_export_peers(pk_ctx *ctx, const pk_dpoly *p, uint32_t n_peers, uint32_t rank, const void *d_sizes,
                          const uint64_t *key_ptrs, const uint64_t *mag_ptrs, const uint64_t *sign_ptrs, uint64_t cap_terms,
                          uint32_t cap_limbs);
/* Order-independent 64-bit digest of (key, coefficient) over all terms, computed on the device.  Digests of
 * disjoint partitions add up (mod 2^64) to the digest of the whole product. */
int pk_dpoly_digest(pk_ctx *ctx, const pk_dpoly *p, uint64_t *digest_out);

#ifdef __cplusplus
}
#endif
#endif /* PIRANHA_B200_H */
