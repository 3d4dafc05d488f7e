"""Quick IVF_PQ throughput probe (development aid, not the judged bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
import time

import numpy as np
import torch

import knowhere_b200 as kb
from knowhere_b200 import datagen

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
nlist = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
nprobe = int(sys.argv[4]) if len(sys.argv) > 4 else 64
d, m, k = 128, 16, 10
dev = torch.device("cuda:0")
t = time.time()
xb = datagen.clustered_torch(n, d, 42, dev)
xq = datagen.clustered_torch(nq, d, 43, dev)
torch.cuda.synchronize()
print("gen", time.time() - t)
ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m, "nbits": 8, "refine": True, "refine_type": "flat"})
ix.set_stream(torch.cuda.current_stream().cuda_stream)
t = time.time()
ix.train(xb)
torch.cuda.synchronize()
print("train", time.time() - t)
t = time.time()
ix.add(xb)
torch.cuda.synchronize()
print("add", time.time() - t)
ix.enable_kernel_timing(True)
for rk in (1, 4):
    cfg = {"nprobe": nprobe, "refine_k": rk}
    t = time.time()
    ids, dist = ix.search(xq, k, cfg)
    torch.cuda.synchronize()
    print("first search (incl seal)", time.time() - t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ix.search(xq, k, cfg, out=(ids, dist))
    e0.record()
    for _ in range(5):
        ix.search(xq, k, cfg, out=(ids, dist))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    c = ix.last_counters()
    print(f"refine_k={rk}: {ms:.3f} ms/batch  qps={nq / ms * 1e3:.0f}  scan_kernel_ms={ix.last_kernel_ms():.3f} "
          f"codes={c['codes']} GBps_alg={c['code_bytes'] / ix.last_kernel_ms() / 1e6:.1f} launches={c['launches']}")
    # recall vs exact
    gt, _ = kb.brute_force_search(xb, xq[:1000].contiguous(), k, "L2")
    rec = datagen.recall(gt.cpu().numpy(), ids[:1000].cpu().numpy())
    print("recall@10 =", rec)
xh = xq.cpu().numpy()
t = time.time()
ids_h, dist_h = ix.search(xh, k, {"nprobe": nprobe, "refine_k": 4})
print("e2e host call", time.time() - t, "equal:", np.array_equal(ids_h, ids.cpu().numpy()))
