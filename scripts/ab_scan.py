"""A/B of scan-kernel variants on ONE box / ONE index (development aid)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowhere_b200 as kb
from knowhere_b200 import datagen
n, d, nq, k = 10_000_000, 128, 10000, 10
dev = torch.device("cuda:0")
xb = datagen.clustered_torch(n, d, 42, dev); xq = datagen.clustered_torch(nq, d, 43, dev)
ix = kb.Index("IVF_PQ", "L2", d, {"nlist": 4096, "m": 16, "nbits": 8, "refine": True, "refine_type": "flat"})
ix.set_stream(torch.cuda.current_stream().cuda_stream)
ix.build(xb)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dis = torch.empty((nq, k), dtype=torch.float32, device=dev)
cfg = {"nprobe": 64, "refine_k": 4}
ix.enable_kernel_timing(True)
base = None
for rep in range(2):
    for nt, acc, noshare in itertools.product((256,), (2,), (0, 1)):
        os.environ["KB2_SCAN_NT"] = str(nt); os.environ["KB2_SCAN_ACC"] = str(acc); os.environ["KB2_SCAN_FULLMERGE"] = str(noshare)
        for _ in range(3):
            ix.search(xq, k, cfg, out=(ids, dis))
        ks = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            ix.search(xq, k, cfg, out=(ids, dis)); ks.append(ix.last_kernel_ms())
        e1.record(); torch.cuda.synchronize()
        if base is None: base = ids.clone()
        print(f"rep{rep} NT={nt} ACC={acc} fullmerge={noshare}: step {e0.elapsed_time(e1)/8:.3f} ms  scan kernel {sum(ks)/len(ks):.3f} ms  same_ids={bool((ids==base).all())}", flush=True)
