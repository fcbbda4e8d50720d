"""Multi-GPU plumbing: one process per GPU, each computing a disjoint output Kronecker-code range.

Both operands are replicated (they are at most a few MB); rank r calls pk_mul_dev with part_index=r and
part_count=world_size, so every term product is performed by exactly one rank and no reduction is needed.
The only exchange is the final gather of result sizes and terms: on GPUs PeerExchange (stores over NVLink into
peer-mapped result arrays by a kernel of this library, torch symmetric memory for the mapping and the barriers), or
torch.distributed collectives (NCCL; gloo in the CPU tests).  Ranges are ordered, so laying the per-rank sorted outputs out in rank order gives the
globally sorted result.  This is the GPU form of the zone scheme of polynomial.hpp:1783-1966 lifted to ranks.

The term gather is two collectives: the sizes are all-gathered first; every rank then writes its own terms straight
into its segment of one exchange buffer (device-to-device, pk_dpoly_export_padded), ONE all-gather moves all segments,
and the rank-ordered key / magnitude / sign arrays are cut out of the gathered segments.
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


class GatherWorkspace:
    """Buffers of the term exchange, reused across products (grow-only).  Tensors that took part in an NCCL collective
    are released lazily by the caching allocator (they are recorded on NCCL's stream), so allocating 100+ MB of fresh
    buffers per product costs milliseconds (measured: 7.4 ms instead of 1.4 ms per pearce1 product at 2 GPUs)."""

    def __init__(self):
        self.send = self.recv = self.keys = self.mags = self.signs = None

    @staticmethod
    def _grow(t, n, dtype, device):
        if t is None or t.numel() < n:
            t = torch.empty(max(n + n // 8, 16), dtype=dtype, device=device)
        return t

    def exchange(self, seg_bytes: int, world: int, device):
        self.send = self._grow(self.send, seg_bytes, torch.uint8, device)
        self.recv = self._grow(self.recv, seg_bytes * world, torch.uint8, device)
        return self.send[:seg_bytes], self.recv[:seg_bytes * world]

    def result(self, n: int, limbs: int, device):
        self.keys = self._grow(self.keys, n, torch.int64, device)
        self.mags = self._grow(self.mags, n * limbs, torch.int64, device)
        self.signs = self._grow(self.signs, n, torch.int8, device)
        return self.keys[:n], self.mags[:n * limbs].view(n, limbs), self.signs[:n]


def segment_layout(nmax: int, limbs: int) -> Tuple[int, int, int]:
    """One rank's segment of the exchange buffer: [keys nmax x 8 B | magnitudes nmax x limbs x 8 B | signs nmax B],
    padded to 16 bytes.  Returns (byte offset of the magnitudes, byte offset of the signs, segment bytes)."""
    mag_off = nmax * 8
    sign_off = mag_off + nmax * limbs * 8
    return mag_off, sign_off, (sign_off + nmax + 15) // 16 * 16


def gather_segments(send: torch.Tensor, recv: torch.Tensor, counts: List[int], limbs: int, workspace: GatherWorkspace,
                    group=None):
    """ONE collective for all terms: every rank contributes its segment (already filled), then the rank-ordered key,
    magnitude and sign arrays are cut out of the gathered segments (three concatenations into the workspace)."""
    world = dist.get_world_size(group)
    nmax = max(max(counts), 1)
    mag_off, sign_off, seg = segment_layout(nmax, limbs)
    dist.all_gather_into_tensor(recv, send, group=group)
    segs = recv.view(world, seg)
    total = sum(counts)
    keys, mags, signs = workspace.result(total, limbs, send.device)
    live = [r for r in range(world) if counts[r]]
    if live:
        torch.cat([segs[r, :counts[r] * 8].view(torch.int64) for r in live], out=keys)
        torch.cat([segs[r, mag_off:mag_off + counts[r] * limbs * 8].view(torch.int64) for r in live], out=mags.view(-1))
        torch.cat([segs[r, sign_off:sign_off + counts[r]].view(torch.int8) for r in live], out=signs)
    return keys, mags, signs


def gather_terms(key: torch.Tensor, mag: torch.Tensor, sign: torch.Tensor, counts: List[int], limbs: int, group=None,
                 workspace: GatherWorkspace = None):
    """Gather variable-size per-rank term arrays (torch tensors) into the global rank-ordered arrays on every rank.

    key int64[n_r], mag int64[n_r, limbs_r] (bit pattern of the uint64 limbs), sign int8[n_r]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    workspace = workspace or GatherWorkspace()
    n = counts[rank]
    nmax = max(max(counts), 1)
    mag_off, sign_off, seg = segment_layout(nmax, limbs)
    send, recv = workspace.exchange(seg, world, key.device)
    if n:
        send[:n * 8].view(torch.int64).copy_(key[:n])
        m = mag.view(n, -1) if mag.dim() == 1 else mag[:n]
        dst = send[mag_off:mag_off + n * limbs * 8].view(torch.int64).view(n, limbs)
        dst.zero_()
        dst[:, :m.shape[1]] = m
        send[sign_off:sign_off + n].view(torch.int8).copy_(sign[:n])
    return gather_segments(send, recv, counts, limbs, workspace, group)


def _order_ctx_before_torch(ctx):
    """Work queued on the context's stream must be complete before torch-side work (symmetric-memory barrier, NCCL
    collective, tensor views) that depends on it.  Nothing to do when the context was created on torch's current
    stream; a context on a private stream is synchronised."""
    if ctx.stream is None or ctx.stream != torch.cuda.current_stream().cuda_stream:
        ctx.synchronize()


def _order_torch_before_ctx(ctx):
    """The reverse: torch-side work (the barrier) must be complete before the context's next kernel."""
    if ctx.stream is None or ctx.stream != torch.cuda.current_stream().cuda_stream:
        torch.cuda.current_stream().synchronize()


def export_local(dpoly, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Copy a device polynomial into torch tensors on the same GPU (device-to-device, no host round trip)."""
    n, limbs = len(dpoly), max(1, dpoly.limbs)
    key = torch.empty(n, dtype=torch.int64, device=device)
    mag = torch.empty((n, limbs), dtype=torch.int64, device=device)
    sign = torch.empty(n, dtype=torch.int8, device=device)
    if n:
        dpoly.export_to(key.data_ptr(), mag.data_ptr(), sign.data_ptr())
    return key, mag, sign


class PeerExchange:
    """Term exchange of a partitioned product over peer-mapped memory (NVLink / NVSwitch), without a collective on the
    data path: the result arrays of every rank live in one torch symmetric-memory allocation, every rank publishes the
    size of its part in all size tables, and after a barrier ONE kernel of this library (pk_dpoly_export_peers) stores
    the rank's terms straight into the result arrays of all ranks at the offset it derives from the size table.  A second
    barrier completes the product; only then does the host read the size table (the single host round trip).

    Layout of the symmetric buffer (bytes): [2 size tables of world x {count, limbs} uint64, alternating between
    products | keys cap x 8 | magnitudes cap x cap_limbs x 8 | signs cap].  Buffers grow collectively: every rank sees the
    same size table, so every rank takes the same decision."""

    HEADER = 1024

    def __init__(self, device, group=None, capacity: int = 1 << 16, limbs: int = 2):
        self.device, self.group = device, group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if 2 * self.world * 16 > self.HEADER:
            raise ValueError("too many ranks for the size tables")
        self.parity = 0
        self.buf = self.hdl = None
        self._allocate(capacity, limbs)

    def _allocate(self, capacity: int, limbs: int):
        import torch.distributed._symmetric_memory as symm

        self.cap, self.cap_limbs = int(capacity), int(limbs)
        self.key_off = self.HEADER
        self.mag_off = self.key_off + self.cap * 8
        self.sign_off = self.mag_off + self.cap * self.cap_limbs * 8
        nbytes = (self.sign_off + self.cap + 255) // 256 * 256
        self.buf = self.hdl = None  # release the old allocation on every rank before the new rendezvous
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf[:self.HEADER].zero_()
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self.hdl.barrier()

    def _slots(self, parity: int):
        off = parity * self.world * 16 + self.rank * 16
        return [b + off for b in self.base]

    def exchange(self, part):
        """All ranks call this with their part of one product; returns (keys, mags, signs) views of the complete result
        in rank (= code) order, valid until the next exchange."""
        w = self.world
        while True:
            par = self.parity
            self.parity ^= 1
            part.publish_size(self._slots(par))
            _order_ctx_before_torch(part.ctx)  # the size stores precede the barrier
            self.hdl.barrier()
            _order_torch_before_ctx(part.ctx)  # the push kernel reads a complete size table
            sizes_ptr = self.base[self.rank] + par * w * 16
            part.export_peers(self.rank, sizes_ptr, [b + self.key_off for b in self.base], [b + self.mag_off for b in self.base],
                              [b + self.sign_off for b in self.base], self.cap, self.cap_limbs)
            _order_ctx_before_torch(part.ctx)  # the term stores precede the barrier that completes the exchange
            self.hdl.barrier()
            sizes = self.buf[par * w * 16:(par + 1) * w * 16].view(torch.int64).cpu().view(w, 2)  # the host round trip
            counts = [int(x) for x in sizes[:, 0]]
            total = sum(counts)
            limbs = max([int(sizes[r, 1]) for r in range(w) if counts[r]] or [1])
            if total <= self.cap and limbs <= self.cap_limbs:
                break
            self._allocate(max(total + total // 8, self.cap), max(limbs, self.cap_limbs))  # nothing was written: grow, repeat
        keys = self.buf[self.key_off:self.key_off + total * 8].view(torch.int64)
        mags = self.buf[self.mag_off:self.mag_off + total * limbs * 8].view(torch.int64).view(total, limbs)
        signs = self.buf[self.sign_off:self.sign_off + total].view(torch.int8)
        return counts, (keys, mags, signs)


def multiply_partitioned(ctx, da, db, device, group=None, gather: bool = True, workspace: "GatherWorkspace" = None,
                         peers: "PeerExchange" = None, **opt):
    """One multi-GPU product: this rank's partition through the C ABI, then the size + term gather -- over peer memory
    when a PeerExchange is given, else two collectives (into `workspace` when given).  The returned tensors are views
    that the next product overwrites."""
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
    if peers is not None:  # peer-memory exchange (NVLink stores by this library's kernel, no collective on the data path)
        _, out = peers.exchange(part)
        mark()
        if trace:
            print("[pk dist] " + "  ".join(f"{n} {1e3 * (b - a):.3f} ms" for n, a, b in zip(("mul_dev", "peer exchange"), t, t[1:])), flush=True)
        return part, out
    counts, limbs = gather_sizes(len(part), max(1, part.limbs), device, group)
    mark()
    workspace = workspace or GatherWorkspace()
    nmax = max(max(counts), 1)
    mag_off, sign_off, seg = segment_layout(nmax, limbs)
    send, recv = workspace.exchange(seg, world, device)
    mark()
    if counts[rank]:  # this rank's terms go straight into its segment of the exchange buffer
        base = send.data_ptr()
        part.export_to(base, base + mag_off, base + sign_off, mag_limbs=limbs)
        _order_ctx_before_torch(ctx)  # the segment is complete before NCCL reads it
    mark()
    out = gather_segments(send, recv, counts, limbs, workspace, group)
    mark()
    if trace:
        names = ("mul_dev", "sizes", "alloc", "export", "exchange")
        print("[pk dist] " + "  ".join(f"{n} {1e3 * (b - a):.3f} ms" for n, a, b in zip(names, t, t[1:])), flush=True)
    return part, out
