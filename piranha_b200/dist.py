"""Multi-GPU plumbing: one process per GPU, each computing a disjoint output Kronecker-code range.

Both operands are replicated (they are at most a few MB); rank r calls pk_mul_dev with part_index=r and
part_count=world_size, so every term product is performed by exactly one rank and no reduction is needed.
The only exchange is the final gather of result sizes and terms (torch.distributed: NCCL over NVLink on GPUs,
gloo in the CPU tests).  Ranges are ordered, so concatenating the per-rank sorted outputs in rank order gives the
globally sorted result.  This is the GPU form of the zone scheme of polynomial.hpp:1783-1966 lifted to ranks.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def gather_sizes(n_local: int, limbs_local: int, device, group=None) -> Tuple[List[int], int]:
    """All-gather (term count, limbs) of every rank; returns (counts per rank, max limbs)."""
    world = dist.get_world_size(group)
    mine = torch.tensor([n_local, limbs_local], dtype=torch.int64, device=device)
    allv = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().view(world, 2)
    return [int(x) for x in allv[:, 0]], int(allv[:, 1].max())


def gather_terms(key: torch.Tensor, mag: torch.Tensor, sign: torch.Tensor, counts: List[int], limbs: int, group=None):
    """Gather variable-size per-rank term arrays into the global (rank-ordered) arrays on every rank.

    key int64[n_r], mag int64[n_r, limbs] (bit pattern of the uint64 limbs), sign int8[n_r].  Buffers are padded to
    the largest count so that one all_gather_into_tensor per array suffices (three collectives in total)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = key.device
    nmax = max(max(counts), 1)
    n = counts[rank]

    def padded(t, shape, dtype):
        buf = torch.zeros(shape, dtype=dtype, device=device)
        if n:
            buf[:n] = t[:n]
        return buf

    if mag.dim() == 1:
        mag = mag.view(-1, 1)
    if mag.shape[1] < limbs:
        mag = torch.nn.functional.pad(mag, (0, limbs - mag.shape[1]))
    # flat outputs (rank-major concatenation): the one layout both NCCL and gloo accept
    k_all = torch.empty(world * nmax, dtype=torch.int64, device=device)
    m_all = torch.empty(world * nmax * limbs, dtype=torch.int64, device=device)
    s_all = torch.empty(world * nmax, dtype=torch.int8, device=device)
    dist.all_gather_into_tensor(k_all, padded(key, (nmax,), torch.int64), group=group)
    dist.all_gather_into_tensor(m_all, padded(mag, (nmax, limbs), torch.int64).view(-1), group=group)
    dist.all_gather_into_tensor(s_all, padded(sign, (nmax,), torch.int8), group=group)
    k_all, m_all, s_all = k_all.view(world, nmax), m_all.view(world, nmax, limbs), s_all.view(world, nmax)
    keys = torch.cat([k_all[r, :counts[r]] for r in range(world)])
    mags = torch.cat([m_all[r, :counts[r]] for r in range(world)])
    signs = torch.cat([s_all[r, :counts[r]] for r in range(world)])
    return keys, mags, signs


def export_local(dpoly, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Copy a device polynomial into torch tensors on the same GPU (device-to-device, no host round trip)."""
    n, limbs = len(dpoly), max(1, dpoly.limbs)
    key = torch.empty(n, dtype=torch.int64, device=device)
    mag = torch.empty((n, limbs), dtype=torch.int64, device=device)
    sign = torch.empty(n, dtype=torch.int8, device=device)
    if n:
        dpoly.export_to(key.data_ptr(), mag.data_ptr(), sign.data_ptr())
    return key, mag, sign


def multiply_partitioned(ctx, da, db, device, group=None, gather: bool = True, **opt):
    """One multi-GPU product: this rank's partition through the C ABI, then the size + term gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    part = ctx.mul_dev(da, db, part_index=rank, part_count=world, **opt)
    if not gather:
        return part, None
    key, mag, sign = export_local(part, device)
    counts, limbs = gather_sizes(len(part), max(1, part.limbs), device, group)
    return part, gather_terms(key, mag, sign, counts, limbs, group)
