"""The N>1 exchange step on CPU: two ranks (gloo) each hold the terms of a disjoint output Kronecker-code range and
gather sizes + terms exactly as bench.py / piranha_b200.dist do over NCCL.  The per-rank partitions come from the CPU
oracle (the GPU computes them on the box; tests/test_gpu_parity.py::test_output_partitions covers that half)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from polyhelper import cpu_ref, terms


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, parts, whole, ragged):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from piranha_b200 import dist as pdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        key, mag, sign = parts[rank]
        k = torch.from_numpy(key.copy())
        m = torch.from_numpy(mag.view(np.int64).copy())
        s = torch.from_numpy(sign.copy())
        counts, limbs = pdist.gather_sizes(len(key), mag.shape[1], torch.device("cpu"))
        assert counts == [len(p[0]) for p in parts]
        assert limbs == max(p[1].shape[1] for p in parts)
        gk, gm, gs = pdist.gather_terms(k, m, s, counts, limbs)
        wk, wm, ws = whole
        assert np.array_equal(gk.numpy(), wk)
        got = gm.numpy().view(np.uint64)
        L = max(got.shape[1], wm.shape[1])
        assert np.array_equal(np.pad(got, ((0, 0), (0, L - got.shape[1]))), np.pad(wm, ((0, 0), (0, L - wm.shape[1]))))
        assert np.array_equal(gs.numpy(), ws)
        # ranges are ordered, so the gathered keys are globally sorted
        assert np.all(gk.numpy()[1:] > gk.numpy()[:-1]) or len(wk) < 2
    finally:
        dist.destroy_process_group()


def _split(whole, cuts, trim):
    key, mag, sign = whole
    parts, lo = [], 0
    for hi in list(cuts) + [len(key)]:
        m = mag[lo:hi]
        if trim:  # a rank whose coefficients are narrow reports fewer limbs
            L = m.shape[1]
            while L > 1 and not m[:, L - 1].any():
                L -= 1
            m = np.ascontiguousarray(m[:, :L])
        parts.append((key[lo:hi].copy(), m.copy(), sign[lo:hi].copy()))
        lo = hi
    return parts


@pytest.mark.parametrize("case", ["balanced", "ragged", "empty_rank"])
def test_two_rank_gather(case):
    (f, g), nv = po.pearce_operands(4), 5
    whole = cpu_ref(terms(f, nv, 1), terms(g, nv, 2), nv)
    n = len(whole[0])
    cut = {"balanced": n // 2, "ragged": n // 7, "empty_rank": 0}[case]
    parts = _split(whole, [cut], trim=True)
    mp.spawn(_worker, args=(2, _free_port(), parts, whole, case), nprocs=2, join=True)
