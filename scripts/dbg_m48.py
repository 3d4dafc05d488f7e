import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knowhere_b200 as kb
from knowhere_b200 import datagen
m, d = int(sys.argv[1]), int(sys.argv[2])
nb, nlist, nq, k = 30000, 64, int(sys.argv[3]) if len(sys.argv) > 3 else 100, 10
xb, xq = datagen.clustered(nb, d, 42), datagen.clustered(nq, d, 43)
ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m}); ix.build(xb)
print("built", flush=True)
for np_ in (8, 1, 64):
    ids, dist = ix.search(xq, k, {"nprobe": np_})
    print("nprobe", np_, "ok", ids[0][:5], flush=True)
