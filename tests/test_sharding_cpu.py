"""N>1 host logic on CPU: world_size-2 gloo.  Each rank searches only the lists it owns (l % world) with the
numpy oracle, ONE all_gather of packed (id, distance) candidates, merge; the result must equal the unsharded
search.  This is the same code path bench.py drives over NCCL (knowhere_b200/sharding.py)."""
import os

import numpy as np
import pytest

from knowhere_b200 import datagen, sharding
from oracle import knowhere_oracle as ko


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nb, d, nlist, m, nq, k, nprobe = 3000, 32, 16, 8, 12, 5, 6
        xb, xq = datagen.clustered(nb, d, 42), datagen.clustered(nq, d, 43)
        rng = np.random.default_rng(0)
        cent = xb[rng.choice(nb, nlist, replace=False)].copy()
        assign = ko.pairwise_keys(xb, cent, ko.L2).argmin(1)
        full = {l: (np.nonzero(assign == l)[0].astype(np.int64), xb[assign == l].copy()) for l in range(nlist)}
        empty = (np.empty(0, np.int64), np.empty((0, d), np.float32))
        mine = {l: (full[l] if sharding.owner_of_list(l, world) == rank else empty) for l in range(nlist)}
        o = ko.IvfOracle("IVF_FLAT", ko.L2, cent, mine)           # same coarse quantizer on every rank
        ids, dis = o.search(xq, k, nprobe)
        mi, md = sharding.gather_and_merge(
            torch, dist, torch.from_numpy(ids), torch.from_numpy(dis),
            lambda gi, gd: sharding.merge_topk_numpy(gi.numpy(), gd.numpy(), "L2"), world)
        I0, D0 = ko.IvfOracle("IVF_FLAT", ko.L2, cent, full).search(xq, k, nprobe)
        ok = bool(np.array_equal(mi, I0) and np.allclose(md, D0))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_rank_list_sharding_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_pack_unpack_roundtrip():
    import torch
    ids = torch.tensor([[3, -1], [7, 9]], dtype=torch.int64)
    dist = torch.tensor([[-1.5, 3.4e38], [0.0, 2.25]], dtype=torch.float32)
    p = sharding.pack_candidates(torch, ids, dist)
    gi, gd = sharding.unpack_candidates(torch, p[None], 2)
    assert torch.equal(gi[0], ids) and torch.equal(gd[0], dist)
