"""which part of the GPU build lowers IVF_PQ recall vs the reference build? (coarse centroids / PQ codebooks / encoder)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knowhere_b200 as kb
from knowhere_b200 import datagen
from oracle import ref

nb, d, nlist, m, nq, k, nprobe = 40000, 128, 128, 16, 1000, 10, 16
xb, xq = datagen.clustered(nb, d, 42), datagen.clustered(nq, d, 43)
gt, _ = ref.flat_search(xb, xq, k, 0)

def rec(ids):
    return np.mean([len(set(a) & set(b)) for a, b in zip(gt, ids)]) / k

r = ref.RefIvf("IVF_PQ", d, 0, nlist, m, 8); r.train(xb); r.add(xb)
c_ref, pq_ref = r.centroids(), r.pq_centroids()
print("A ref coarse + ref pq + ref encode  (cpu search):", rec(r.search(xq, k, nprobe)[0]))
ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m}); ix.build(xb)
c_gpu, pq_gpu = ix.ivf_export_centroids(m)
print("B gpu coarse + gpu pq + gpu encode  (gpu search):", rec(ix.search(xq, k, {"nprobe": nprobe})[0]))

def gpu_encode(c, pq):
    c, pq = np.ascontiguousarray(c), np.ascontiguousarray(pq)
    i2 = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m})
    kb._check(kb.lib().kb2_ivf_import_begin(i2.h, nlist, c.ctypes.data, pq.ctypes.data))
    i2.add(xb)
    return rec(i2.search(xq, k, {"nprobe": nprobe})[0])

print("C ref coarse + ref pq + gpu encode:", gpu_encode(c_ref, pq_ref))
print("D gpu coarse + ref pq(!) + gpu encode:", gpu_encode(c_gpu, pq_ref))
# PQ retrained by the reference on residuals wrt GPU coarse centroids is not expressible here; instead:
print("E ref coarse + gpu pq(!) + gpu encode:", gpu_encode(c_ref, pq_gpu))
# list balance
ls_ref = np.array([len(i) for _, i, _ in r.lists()]); ls_gpu = np.array([ix.L.kb2_ivf_list_size(ix.h, l) for l in range(nlist)])
print("list size std ref", ls_ref.std(), "gpu", ls_gpu.std(), "max", ls_ref.max(), ls_gpu.max())
