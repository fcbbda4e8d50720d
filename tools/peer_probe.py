"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work between the ranks of this box?
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    buf = symm.empty(1 << 20, dtype=torch.int64, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    torch.cuda.synchronize()
    print(f"[{rank}] rendezvous {1e3 * (time.perf_counter() - t0):.1f} ms; multicast {hdl.has_multicast_support}; "
          f"ptrs {[hex(p) for p in hdl.buffer_ptrs]}; signal pad {hdl.signal_pad_size} B", flush=True)
    buf.zero_()
    hdl.barrier()
    mine = torch.full((16,), rank + 1, dtype=torch.int64, device=dev)
    for peer in range(world):
        hdl.get_buffer(peer, (16,), torch.int64, 16 * rank).copy_(mine)
    hdl.barrier()
    got = buf[:16 * world].view(world, 16)[:, 0].tolist()
    assert got == [r + 1 for r in range(world)], got
    # timing of a barrier and of a big peer copy
    big = torch.ones(16 << 20, dtype=torch.uint8, device=dev)
    peer = (rank + 1) % world
    dst = hdl.get_buffer(peer, (8 << 20,), torch.uint8, 0)
    for _ in range(3):
        hdl.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(10):
        hdl.barrier()
    e[1].record()
    for _ in range(10):
        dst.copy_(big[:8 << 20])
    e[2].record()
    torch.cuda.synchronize()
    print(f"[{rank}] ok; barrier {e[0].elapsed_time(e[1]) * 100:.1f} us; 8 MiB peer copy {e[1].elapsed_time(e[2]) * 100:.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
