"""Multi-GPU plumbing: one process per GPU, each computing a disjoint output Kronecker-code range.

Both operands are replicated (they are at most a few MB); rank r calls pk_mul_dev with part_index=r and
part_count=world_size, so every term product is performed by exactly one rank and no reduction is needed.
The only exchange is the final gather of result sizes and terms (torch.distributed: NCCL over NVLink on GPUs,
gloo in the CPU tests).  Ranges are ordered, so laying the per-rank sorted outputs out in rank order gives the
globally sorted result.  This is the GPU form of the zone scheme of polynomial.hpp:1783-1966 lifted to ranks.

The term gather moves every byte once: the sizes are all-gathered first, every rank then writes its own terms
straight into its slice of the common result buffers (device-to-device, pk_dpoly_export_padded) and each slice is
broadcast in place from its owner -- no padding to the largest rank, no staging copies, no concatenation.
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


def _offsets(counts: List[int]) -> List[int]:
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    return off


def exchange_slices(keys: torch.Tensor, mags: torch.Tensor, signs: torch.Tensor, counts: List[int], group=None):
    """In-place exchange: rank r owns rows [off_r, off_r + counts[r]) of the three result buffers (already filled
    locally) and broadcasts them to everyone.  3 * world collectives, all asynchronous, each moving its bytes once."""
    world = dist.get_world_size(group)
    off = _offsets(counts)
    work = []
    for r in range(world):
        if counts[r] == 0:
            continue
        src = dist.get_global_rank(group, r) if group is not None else r
        for buf in (keys, mags, signs):
            work.append(dist.broadcast(buf[off[r]:off[r + 1]], src=src, group=group, async_op=True))
    for w in work:
        w.wait()
    return keys, mags, signs


def gather_terms(key: torch.Tensor, mag: torch.Tensor, sign: torch.Tensor, counts: List[int], limbs: int, group=None):
    """Gather variable-size per-rank term arrays (torch tensors) into the global rank-ordered arrays on every rank.

    key int64[n_r], mag int64[n_r, limbs_r] (bit pattern of the uint64 limbs), sign int8[n_r]."""
    rank = dist.get_rank(group)
    device = key.device
    off = _offsets(counts)
    n = counts[rank]
    keys = torch.empty(off[-1], dtype=torch.int64, device=device)
    mags = torch.zeros((off[-1], limbs), dtype=torch.int64, device=device) if mag.dim() == 1 or mag.shape[-1] < limbs \
        else torch.empty((off[-1], limbs), dtype=torch.int64, device=device)
    signs = torch.empty(off[-1], dtype=torch.int8, device=device)
    if n:
        keys[off[rank]:off[rank + 1]] = key[:n]
        m = mag.view(n, -1) if mag.dim() == 1 else mag[:n]
        mags[off[rank]:off[rank + 1], :m.shape[1]] = m
        signs[off[rank]:off[rank + 1]] = sign[:n]
    return exchange_slices(keys, mags, signs, counts, group)


def export_local(dpoly, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Copy a device polynomial into torch tensors on the same GPU (device-to-device, no host round trip)."""
    n, limbs = len(dpoly), max(1, dpoly.limbs)
    key = torch.empty(n, dtype=torch.int64, device=device)
    mag = torch.empty((n, limbs), dtype=torch.int64, device=device)
    sign = torch.empty(n, dtype=torch.int8, device=device)
    if n:
        dpoly.export_to(key.data_ptr(), mag.data_ptr(), sign.data_ptr())
    return key, mag, sign


class GatherWorkspace:
    """Result buffers reused across products (grow-only).  Tensors that took part in an NCCL collective are released
    lazily by the caching allocator (they are recorded on NCCL's stream), so allocating 100+ MB of fresh result buffers
    per product costs milliseconds; a workspace owned by the caller avoids it."""

    def __init__(self):
        self.keys = self.mags = self.signs = None

    def get(self, n: int, limbs: int, device):
        if self.keys is None or self.keys.numel() < n or self.mags.numel() < n * limbs:
            cap = max(n + n // 8, 1)
            self.keys = torch.empty(cap, dtype=torch.int64, device=device)
            self.mags = torch.empty(cap * limbs, dtype=torch.int64, device=device)
            self.signs = torch.empty(cap, dtype=torch.int8, device=device)
        return self.keys[:n], self.mags[:n * limbs].view(n, limbs), self.signs[:n]


def multiply_partitioned(ctx, da, db, device, group=None, gather: bool = True, workspace: "GatherWorkspace" = None, **opt):
    """One multi-GPU product: this rank's partition through the C ABI, then the size + term gather (into `workspace`
    when given: the returned tensors are then views that the next product overwrites)."""
    import os
    import time

    trace = os.environ.get("PK_DIST_TRACE") and dist.get_rank(group) == 0
    t = [time.perf_counter()]

    def mark():
        if trace:
            torch.cuda.synchronize(device)
            t.append(time.perf_counter())

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    part = ctx.mul_dev(da, db, part_index=rank, part_count=world, **opt)
    mark()
    if not gather:
        return part, None
    counts, limbs = gather_sizes(len(part), max(1, part.limbs), device, group)
    mark()
    off = _offsets(counts)
    if workspace is not None:
        keys, mags, signs = workspace.get(off[-1], limbs, device)
    else:
        keys = torch.empty(off[-1], dtype=torch.int64, device=device)
        mags = torch.empty((off[-1], limbs), dtype=torch.int64, device=device)
        signs = torch.empty(off[-1], dtype=torch.int8, device=device)
    mark()
    if counts[rank]:  # this rank's terms go straight into its slice of the result
        lo = off[rank]
        part.export_to(keys[lo:].data_ptr(), mags[lo:].data_ptr(), signs[lo:].data_ptr(), mag_limbs=limbs)
    mark()
    out = exchange_slices(keys, mags, signs, counts, group)
    mark()
    if trace:
        names = ("mul_dev", "sizes", "alloc", "export", "exchange")
        print("[pk dist] " + "  ".join(f"{n} {1e3 * (b - a):.3f} ms" for n, a, b in zip(names, t, t[1:])), flush=True)
    return part, out
