"""Debug aid: expected survivor counts of the tensor-core PQ filter (numpy emulation, bf16-rounded operands) against
the counters the GPU engine reports, per list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knowhere_b200 as kb
from knowhere_b200 import datagen
from oracle import ref

nb, d, nlist, m, nprobe, nq, k = 60000, 128, 64, 16, 16, 3000, 40
xb = datagen.clustered(nb, d, 42); xq = datagen.clustered(nq, d, 43)
r = ref.RefIvf("IVF_PQ", d, 0, nlist, m, 8)
r.train(xb); r.add(xb)
lists = list(r.lists())
c = {"survivors": -1}
if os.environ.get("TC_DEBUG_GPU", "1") == "1":
    ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m, "nbits": 8})
    ix.ivf_import(r.centroids(), r.pq_centroids(), lists, raw=None)
    os.environ["KB2_PQ_ENGINE"] = "tc"; os.environ["KB2_TC_VERBOSE"] = "1"
    ids, dist = ix.search(xq, k, {"nprobe": nprobe})
    c = ix.last_counters()
    print("gpu counters", c)
    os.environ["KB2_PQ_ENGINE"] = "lut"
    ids0, dist0 = ix.search(xq, k, {"nprobe": nprobe})
    print("engines equal:", np.array_equal(ids, ids0), np.array_equal(dist, dist0))
lists = [(a[-2], a[-1]) for a in lists]

def bf16(x):
    xi = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    xi = ((xi + 0x7FFF + ((xi >> 16) & 1)) >> 16) << 16
    return xi.astype(np.uint32).view(np.float32)

cent = r.centroids(); pq = r.pq_centroids().reshape(m, 256, d // m)
probes, pdis = r.coarse(xq, nprobe)
Rmax = np.sqrt((np.linalg.norm(pq, axis=2).max(axis=1) ** 2).sum()) * 1.0001
dec = []; t1s = []
for l in range(nlist):
    idl, codes = lists[l]
    codes = np.asarray(codes).reshape(-1, m)
    dl = pq[np.arange(m)[None, :], codes].reshape(len(idl), d) if len(idl) else np.zeros((0, d), np.float32)
    dec.append(dl); t1s.append((dl ** 2).sum(1) + 2 * dl @ cent[l])
qn = np.linalg.norm(xq, axis=1) * 1.0001
bound = np.full(nq, np.inf, np.float32)
for qi in range(nq):
    keys = []; seen = 0
    for j in range(min(8, nprobe)):
        l = probes[qi, j]
        if seen >= 3000: break
        if len(dec[l]) == 0: continue
        seen += len(dec[l])
        keys.append(((xq[qi] - cent[l]) ** 2).sum() + t1s[l] - 2 * dec[l] @ xq[qi])
    keys = np.sort(np.concatenate(keys))
    if seen >= 4 * k and len(keys) >= k: bound[qi] = keys[k - 1]
print("queries without bound:", int(np.isinf(bound).sum()))
exp_pass = np.zeros(nlist, np.int64); exact_pass = np.zeros(nlist, np.int64)
xq16 = bf16(xq)
for l in range(nlist):
    qs = np.where((probes == l).any(axis=1))[0]
    if len(qs) == 0 or len(dec[l]) == 0: continue
    S = bf16(dec[l]).astype(np.float64) @ xq16[qs].T.astype(np.float64)          # [codes, queries]
    base = ((xq[qs] - cent[l]) ** 2).sum(1)
    margin = 2 * 0.0085 * Rmax * qn[qs] * 1.01
    keyp = base[None, :] + t1s[l][:, None] - 2 * S
    ok = np.isfinite(bound[qs])
    exp_pass[l] = (keyp[:, ok] <= (bound[qs] + margin)[None, ok]).sum()
    Sx = dec[l].astype(np.float64) @ xq[qs].T.astype(np.float64)
    exact_pass[l] = ((base[None, :] + t1s[l][:, None] - 2 * Sx)[:, ok] <= bound[qs][None, ok]).sum()
print("expected filter survivors", int(exp_pass.sum()), "exact survivors", int(exact_pass.sum()), "gpu reported", c["survivors"])
print("per-list expected (top 8):", sorted(exp_pass.tolist(), reverse=True)[:8], " list sizes:", sorted([len(x) for x in dec], reverse=True)[:5])
