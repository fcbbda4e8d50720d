#!/usr/bin/env python
"""Phase breakdown of one partitioned product (run under torchrun): python -m torch.distributed.run --nproc-per-node N tools/scale_phases.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import piranha_b200 as pb
from piranha_b200 import dist as pdist
from bench import build_operands

def main():
    w = sys.argv[1] if len(sys.argv) > 1 else "pearce1"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    device = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=device)
    a, b, nv, L = build_operands(w)
    ctx = pb.Context(lr, torch.cuda.current_stream().cuda_stream)
    da, db = ctx.upload(a, nv), ctx.upload(b, nv)
    def sync():
        torch.cuda.synchronize(device)
        return time.perf_counter()
    for it in range(6):
        dist.barrier(); t0 = sync()
        part = ctx.mul_dev(da, db, part_index=rank, part_count=world); t1 = sync()
        st = ctx.stats()
        counts, limbs = pdist.gather_sizes(len(part), max(1, part.limbs), device); t2 = sync()
        off = pdist._offsets(counts)
        keys = torch.empty(off[-1], dtype=torch.int64, device=device)
        mags = torch.empty((off[-1], limbs), dtype=torch.int64, device=device)
        signs = torch.empty(off[-1], dtype=torch.int8, device=device); t3 = sync()
        lo = off[rank]
        part.export_to(keys[lo:].data_ptr(), mags[lo:].data_ptr(), signs[lo:].data_ptr(), mag_limbs=limbs); t4 = sync()
        pdist.exchange_slices(keys, mags, signs, counts); t5 = sync()
        if rank == 0 and it >= 2:
            print(f"{w} x{world}: mul_dev {1e3*(t1-t0):.3f} ms (device {st['device_ms']:.3f}, kernel {st['mul_kernel_ms']:.3f}) | sizes {1e3*(t2-t1):.3f} | alloc {1e3*(t3-t2):.3f} | "
                  f"export {1e3*(t4-t3):.3f} | exchange {1e3*(t5-t4):.3f} | total {1e3*(t5-t0):.3f}  counts={counts}", flush=True)
        del keys, mags, signs, part
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
