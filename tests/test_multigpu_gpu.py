"""Multi-GPU list sharding on real GPUs (skipped on a 1-GPU box; the gloo CPU test covers the host logic).
Two processes, NCCL: sharded IVF_PQ search + one all-gather + merge kernel == unsharded search."""
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
from knowhere_b200 import datagen, sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
nb, d, nlist, m, nq, k = 50000, 64, 128, 16, 500, 10
xb, xq = datagen.clustered(nb, d, 42), datagen.clustered(nq, d, 43)
cfgb = {"nlist": nlist, "m": m}   # pure ADC: sharded merge == unsharded exactly (refine would add candidates)
full = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
full.build(xb)                                   # each rank trains identically? no: broadcast rank 0's quantizers
cent, pq = full.ivf_export_centroids(m)
ct, pt = torch.from_numpy(cent).to(dev), torch.from_numpy(pq).to(dev)
dist.broadcast(ct, 0); dist.broadcast(pt, 0)
torch.cuda.synchronize()   # the library copies on its own stream: the broadcast must have landed
ref_ix = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
kb._check(kb.lib().kb2_ivf_import_begin(ref_ix.h, nlist, ct.data_ptr(), pt.data_ptr()))
ref_ix.add(xb)
sh = kb.Index("IVF_PQ", "L2", d, cfgb, device=rank)
sh.set_shard(rank, world)
kb._check(kb.lib().kb2_ivf_import_begin(sh.h, nlist, ct.data_ptr(), pt.data_ptr()))
sh.add(xb)
cfg = {"nprobe": 16}
I0, D0 = ref_ix.search(xq, k, cfg)
xq_d = torch.from_numpy(xq).to(dev)
ids, dis = sh.search(xq_d, k, cfg)
stream = torch.cuda.current_stream().cuda_stream
def merge_fn(gi, gd):
    return kb.merge_topk(gi, gd, "L2", device=rank, stream=stream)
mi, md = sharding.gather_and_merge(torch, dist, ids, dis, merge_fn, world)
same = (mi.cpu().numpy() == I0).all(1).mean()
ok = same > 0.97 and np.allclose(np.sort(md.cpu().numpy(), 1), np.sort(D0, 1), rtol=1e-5)   # PQ ties may swap ids
local_only = float((ids.cpu().numpy() == I0).mean())
print(f"rank {rank}: merged==unsharded {ok}; local-only agreement {local_only:.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 3)
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
