"""compare quantisation error of the GPU-built IVF_PQ index with the reference-built one"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knowhere_b200 as kb
from knowhere_b200 import datagen
from oracle import ref

nb, d, nlist, m = 40000, 128, 128, 16
xb = datagen.clustered(nb, d, 42)

def recon_err(cent, pq, lists):
    err = 0.0
    for l, ids, codes in lists:
        if len(ids) == 0: continue
        rec = cent[l][None].repeat(len(ids), 0).reshape(len(ids), m, d // m).copy()
        for mm in range(m):
            rec[:, mm] += pq[mm][codes[:, mm]]
        err += ((xb[ids] - rec.reshape(len(ids), d)) ** 2).sum()
    return err / nb

r = ref.RefIvf("IVF_PQ", d, 0, nlist, m, 8); r.train(xb); r.add(xb)
e_ref = recon_err(r.centroids(), r.pq_centroids(), list(r.lists()))
coarse_ref = ((xb - r.centroids()[np.concatenate([np.full(len(i), l) for l, i, c in r.lists()])][np.argsort(np.concatenate([i for l, i, c in r.lists()]))]) ** 2).sum() / nb
ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m}); ix.build(xb)
c, pq = ix.ivf_export_centroids(m)
L = [(l,) + ix.ivf_export_list(l, m) for l in range(nlist)]
e_gpu = recon_err(c, pq, L)
asg = np.empty(nb, np.int64)
for l, ids, codes in L: asg[ids] = l
coarse_gpu = ((xb - c[asg]) ** 2).sum() / nb
print(f"reconstruction MSE: reference-built {e_ref:.4f}  gpu-built {e_gpu:.4f}   coarse MSE ref {coarse_ref:.4f} gpu {coarse_gpu:.4f}")
# GPU encode with REFERENCE codebooks: isolates encoder from trainer
ix2 = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m})
kb._check(kb.lib().kb2_ivf_import_begin(ix2.h, nlist, r.centroids().ctypes.data, r.pq_centroids().ctypes.data))
ix2.add(xb)
L2_ = [(l,) + ix2.ivf_export_list(l, m) for l in range(nlist)]
print("gpu encode with reference codebooks MSE", recon_err(r.centroids(), r.pq_centroids(), L2_))
