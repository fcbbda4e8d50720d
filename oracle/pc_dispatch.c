/*
 * TEST INFRASTRUCTURE ONLY -- entry point of the CPU restatement (oracle/piranha_cpu.c): picks the accumulator width.
 *
 * piranha_cpu.c is compiled once per width (PC_L = 2, 3, 4 limbs).  mppp::integer<N> keeps a coefficient in its N static
 * limbs for as long as it fits (/root/reference/include/piranha/integer.hpp:67-85); the restatement does the same up
 * front from the bound  |coefficient| <= max|a| * max|b| * min(n1, n2)  and runs the narrowest instance that holds it
 * plus the sign bit.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t n;
    int64_t *key;
    uint64_t *mag;
    int8_t *sign;
    uint32_t limbs;
    double seconds;
    uint64_t buckets;
    uint32_t threads_used;
} pc_result;

#define PC_DECL(name)                                                                                                   \
    int name(uint64_t n1, const int64_t *key1, const uint64_t *mag1, uint32_t limbs1, const int8_t *sign1, uint64_t n2, \
             const int64_t *key2, const uint64_t *mag2, uint32_t limbs2, const int8_t *sign2, uint32_t n_vars,         \
             const int64_t *limits, int64_t h_min, uint32_t n_threads, uint64_t min_work, int trunc_mode,             \
             int64_t max_degree, const uint8_t *deg_mask, pc_result *out)
PC_DECL(pc_mul_L2);
PC_DECL(pc_mul_L3);
PC_DECL(pc_mul_L4);

static unsigned max_bits(uint64_t n, const uint64_t *mag, uint32_t limbs)
{
    unsigned best = 0;
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t l = limbs; l-- > 0;) {
            const uint64_t w = mag[i * limbs + l];
            if (w) {
                const unsigned b = 64u * l + (64u - (unsigned)__builtin_clzll(w));
                if (b > best) best = b;
                break;
            }
        }
    return best;
}

PC_DECL(pc_mul)
{
    if (!out || limbs1 < 1 || limbs1 > 2 || limbs2 < 1 || limbs2 > 2) return 1;
    const uint64_t cnt = n1 < n2 ? n1 : n2;
    unsigned cnt_bits = 0;
    for (uint64_t c = cnt; c; c >>= 1) ++cnt_bits;
    const unsigned need = max_bits(n1, mag1, limbs1) + max_bits(n2, mag2, limbs2) + cnt_bits + 1; /* + sign */
    if (need <= 128)
        return pc_mul_L2(n1, key1, mag1, limbs1, sign1, n2, key2, mag2, limbs2, sign2, n_vars, limits, h_min, n_threads, min_work,
                         trunc_mode, max_degree, deg_mask, out);
    if (need <= 192)
        return pc_mul_L3(n1, key1, mag1, limbs1, sign1, n2, key2, mag2, limbs2, sign2, n_vars, limits, h_min, n_threads, min_work,
                         trunc_mode, max_degree, deg_mask, out);
    if (need <= 256 + 1)
        return pc_mul_L4(n1, key1, mag1, limbs1, sign1, n2, key2, mag2, limbs2, sign2, n_vars, limits, h_min, n_threads, min_work,
                         trunc_mode, max_degree, deg_mask, out);
    return 1;
}

void pc_result_free(pc_result *r)
{
    if (!r) return;
    free(r->key);
    free(r->mag);
    free(r->sign);
    memset(r, 0, sizeof(*r));
}
