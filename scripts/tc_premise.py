"""CPU experiment behind the list-major tensor-core PQ filter: how many codes survive a per-query
bound taken from the P0 nearest lists, and how much a bf16 / tf32 error margin inflates that set."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from knowhere_b200 import datagen
from oracle import ref

n, d, nlist, m, nprobe, K, nq = 1_000_000, 128, 512, 16, 64, 40, 100
xb = datagen.clustered(n, d, 1); xq = datagen.clustered(nq, d, 2)
t = time.time()
ix = ref.RefIvf("IVF_PQ", d, 0, nlist, m=m)
ix.train(xb[:200000]); ix.add(xb)
print("built", time.time() - t)
cent = ix.centroids(); pq = ix.pq_centroids().reshape(m, 256, d // m)
lists = [ix.get_list(l) for l in range(nlist)]
probes, dis = ix.coarse(xq, nprobe)

def rnd(x, bits):  # round-to-nearest to `bits` explicit mantissa bits
    xi = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    sh = 23 - bits
    xi = ((xi + (1 << (sh - 1))) >> sh) << sh
    return xi.astype(np.uint32).view(np.float32)

Rmax = np.sqrt((np.linalg.norm(pq, axis=2).max(axis=1) ** 2).sum())
print("Rmax(global bound on |r^|)", Rmax, " mean |q|", np.linalg.norm(xq, axis=1).mean())
res = {}
for qi in range(nq):
    q = xq[qi]; keys = []; apx_b = []; apx_t = []; first = []; rn = []
    for pi, l in enumerate(probes[qi]):
        ids, codes = lists[l]
        if len(ids) == 0: continue
        codes = codes.reshape(-1, m)
        dec = pq[np.arange(m)[None, :], codes].reshape(len(ids), d)
        base = ((q - cent[l]) ** 2).sum()
        t1 = (dec ** 2).sum(1) + 2 * dec @ cent[l]
        keys.append(base + t1 - 2 * dec @ q)
        apx_b.append(base + t1 - 2 * rnd(dec, 7) @ rnd(q, 7))
        apx_t.append(base + t1 - 2 * rnd(dec, 10) @ rnd(q, 10))
        first.append(np.full(len(ids), pi)); rn.append(np.linalg.norm(dec, axis=1))
    keys = np.concatenate(keys); apx_b = np.concatenate(apx_b); apx_t = np.concatenate(apx_t)
    first = np.concatenate(first); rn = np.concatenate(rn)
    qn = np.linalg.norm(q)
    for P0 in (1, 2, 4, 8):
        sel = keys[first < P0]
        bound = np.sort(sel)[K - 1] if len(sel) >= K else np.inf
        ex = (keys <= bound).sum()
        mb = 2 * 2.1 * 2.0 ** -8 * qn
        mt = 2 * 2.1 * 2.0 ** -11 * qn
        r = res.setdefault(P0, [])
        r.append((ex, (apx_b <= bound + mb * Rmax).sum(), (apx_b <= bound + mb * rn).sum(),
                  (apx_t <= bound + mt * Rmax).sum(), len(keys),
                  np.abs(apx_b - keys).max() / (mb * Rmax), (keys[apx_b > bound + mb * Rmax] <= bound).sum()))
for P0, r in res.items():
    a = np.array(r, dtype=np.float64)
    print(f"P0={P0}: exact survivors mean {a[:,0].mean():.0f} max {a[:,0].max():.0f} | bf16 global-margin {a[:,1].mean():.0f} max {a[:,1].max():.0f}"
          f" | bf16 per-code-margin {a[:,2].mean():.0f} | tf32 global {a[:,3].mean():.0f} | scanned {a[:,4].mean():.0f}"
          f" | max err/margin {a[:,5].max():.3f} | missed {a[:,6].sum():.0f}")
