// Shared-memory zone path (PK_ALGO_ZONE): placeholder until the kernel lands; the dispatcher falls back to
// the HBM/L2 accumulators when zone_multiply() declines.
#pragma once

#include "pk_host.cuh"

namespace pk
{

inline void zone_configure(pk_ctx *) {}

inline u32 choose_algo(pk_ctx *, const Plan &, u64, u64, u64, bool) { return PK_ALGO_AUTO_FALLBACK; }

inline pk_dpoly *zone_multiply(pk_ctx *, const Plan &, const pk_dpoly *, const pk_dpoly *, const u64 *, const u64 *,
                               const u64 *, u64, const u32 *, u64, u64, bool, u32)
{
    return nullptr;
}

} // namespace pk
