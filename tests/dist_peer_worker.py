"""Worker of test_peer_exchange_two_gpus (run under torch.distributed.run, one process per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import piranha_b200 as pb  # noqa: E402
from piranha_b200 import dist as pdist  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from polyhelper import assert_terms_equal, terms  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = pb.Context(local, stream.cuda_stream)
        peers = pdist.PeerExchange(dev, capacity=1000, limbs=1)  # small on purpose: the first large product grows it
        ws = pdist.GatherWorkspace()
        cases = [(po.fateman_operands(8), 4, {}), (po.pearce_operands(6), 5, {}), (po.fateman_operands(8), 4, dict(trunc_mode=1, max_degree=9)),
                 (po.pearce_operands(3), 5, {}), (po.fateman_operands(8), 4, {})]
        for (f, g), nv, opt in cases:
            a, b = terms(f, nv, 1), terms(g, nv, 2)
            da, db = ctx.upload(a, nv), ctx.upload(b, nv)
            whole = ctx.mul_dev(da, db, **opt).download()
            for _ in range(2):  # twice: both size tables
                _, (k, m, s) = pdist.multiply_partitioned(ctx, da, db, dev, peers=peers, **opt)
                got = (k.cpu().numpy(), m.cpu().numpy().view(np.uint64), s.cpu().numpy())
                assert_terms_equal(got, whole)
            _, (k2, m2, s2) = pdist.multiply_partitioned(ctx, da, db, dev, workspace=ws, **opt)
            assert torch.equal(k, k2) and torch.equal(m, m2) and torch.equal(s, s2)
        torch.cuda.synchronize(dev)
    print("PEER-EXCHANGE-OK", rank, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
