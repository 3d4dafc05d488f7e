"""Multi-GPU list sharding on real GPUs (skipped on a 1-GPU box; the gloo CPU test covers the host logic).
Two processes, one NCCL communicator owned by the library (kb2_comm_*): the sharded IVF_PQ search is ONE collective
call (probe all-gather, bound all-reduce, candidate all-gather + merge kernel) and must equal the unsharded search."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import knowhere_b200 as kb
from knowhere_b200 import datagen
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
def bcast(b):
    box = [b]
    dist.broadcast_object_list(box, src=0)
    return box[0]
comm = kb.Comm(rank, world, rank, bcast)          # NCCL communicator owned by libknowhere_b200.so
ok_all = True
for (d, nlist, m, nq, nprobe, tag) in [(64, 128, 16, 500, 16, "lut engine"), (128, 64, 16, 4096, 16, "tc engine")]:
    nb, k = 60000, 10
    xb, xq = datagen.clustered(nb, d, 42), datagen.clustered(nq, d, 43)
    cfgb = {"nlist": nlist, "m": m}   # pure ADC: sharded merge == unsharded exactly (refine would add candidates)
    full = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
    full.build(xb)                                   # every rank trains (deterministic build); rank 0's quantizers are used
    cent, pq = full.ivf_export_centroids(m)
    ct, pt = torch.from_numpy(cent).to(dev), torch.from_numpy(pq).to(dev)
    dist.broadcast(ct, 0); dist.broadcast(pt, 0)
    torch.cuda.synchronize()
    ref_ix = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
    kb._check(kb.lib().kb2_ivf_import_begin(ref_ix.h, nlist, ct.data_ptr(), pt.data_ptr()))
    ref_ix.add(xb)
    sh = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
    sh.set_shard(rank, world)
    kb._check(kb.lib().kb2_ivf_import_begin(sh.h, nlist, ct.data_ptr(), pt.data_ptr()))
    sh.add(xb)
    sh.set_comm(comm)
    cfg = {"nprobe": nprobe}
    ref_ix.enable_kernel_timing(True)
    I0, D0 = ref_ix.search(xq, k, cfg)
    eng = ref_ix.last_stage_info()["engine"]
    mi, md = sh.search(xq, k, cfg)                   # collective: host buffers in, merged global top-k out on every rank
    same_rows = (mi == I0).all(1).mean()
    ok = same_rows > 0.97 and np.allclose(np.sort(md, 1), np.sort(D0, 1), rtol=1e-6)   # PQ ties may swap ids
    # device buffers through the same call
    xq_d = torch.from_numpy(xq).to(dev)
    torch.cuda.synchronize()
    di, dd = sh.search(xq_d, k, cfg)
    ok = ok and np.array_equal(di.cpu().numpy(), mi) and np.array_equal(dd.cpu().numpy(), md)
    print(f"rank {rank} [{tag}, engine {eng}]: merged==unsharded rows {same_rows:.4f} ok={ok}", flush=True)
    ok_all = ok_all and ok
    del sh, ref_ix, full
# HNSW graph-partition sharding: every rank builds a sub-graph over its row slice; the collective search merges the shards
n, d, M, k = 20000, 64, 16, 10
xb, xq = datagen.clustered(n, d, 5), datagen.clustered(200, d, 6)
hs = kb.Index("HNSW", "L2", d, {"M": M, "efConstruction": 100}, device=rank)
hs.set_shard(rank, world)
hs.add(xb)
assert hs.count() == n // world
hs.set_comm(comm)
hi, hd = hs.search(xq, k, {"ef": 64})
gt, _ = kb.brute_force_search(xb, xq, k, "L2", device=rank)
rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(hi, gt)])
mask = np.zeros(n, bool); mask[::3] = True
fi, _ = hs.search(xq, k, {"ef": 64}, bitset=np.packbits(mask, bitorder="little"))
ok_h = rec > 0.9 and not mask[fi[fi >= 0]].any() and (np.diff(hd, axis=1) >= 0).all()
print(f"rank {rank} [hnsw graph partitions]: recall {rec:.3f} ok={ok_h}", flush=True)
ok_all = ok_all and ok_h
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok_all else 3)
'''


def test_two_gpu_list_sharding(kb, tmp_path):
    if kb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = tmp_path / "mg.py"
    p.write_text(SCRIPT % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(p)],
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
