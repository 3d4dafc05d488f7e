"""Host-side check of the guarantee behind the tensor-core IVF_PQ engine, on the CPU:

 1. the bf16 contraction error never exceeds the margin the kernel adds to the bound:
        |S' - <q, r^>| <= 0.0085 |q| Rmax,   Rmax^2 = sum_m max_j |c_pq[m][j]|^2
 2. hence filtering with  key' <= B_q + |alpha| * margin  keeps every code whose exact key is <= B_q, for any valid
    upper bound B_q of the k-th best key (here: the k-th best over the query's nearest lists, as phase A computes it),
    and the exact top-k of the survivors equals the exact top-k of the full scan.

The index (centroids, codebooks, codes) comes from the compiled reference; the arithmetic is emulated in numpy
(round-to-nearest-even bf16 operands, wide accumulation)."""
import numpy as np
import pytest

from knowhere_b200 import datagen


def _bf16(x):
    xi = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    xi = ((xi + 0x7FFF + ((xi >> 16) & 1)) >> 16) << 16
    return xi.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("metric", [0, 1])
def test_bf16_filter_keeps_every_code_under_the_bound(ref, metric):
    nb, d, nlist, m, nprobe, nq, k = 20000, 128, 32, 16, 8, 64, 40
    xb = datagen.clustered(nb, d, 11)
    xq = datagen.clustered(nq, d, 12)
    r = ref.RefIvf("IVF_PQ", d, metric, nlist, m, 8)
    r.train(xb)
    r.add(xb)
    cent = r.centroids()
    pq = r.pq_centroids().reshape(m, 256, d // m)
    lists = [r.get_list(l) for l in range(nlist)]
    probes, pdis = r.coarse(xq, nprobe)
    rmax = float(np.sqrt((np.linalg.norm(pq, axis=2).max(axis=1) ** 2).sum()))
    alpha = 2.0 if metric == 0 else 1.0
    dec = []
    for l in range(nlist):
        codes = np.asarray(lists[l][1]).reshape(-1, m)
        dec.append(pq[np.arange(m)[None, :], codes].reshape(len(codes), d) if len(codes) else np.zeros((0, d), np.float32))
    worst = 0.0
    for qi in range(nq):
        q = xq[qi]
        margin = 0.0085 * float(np.linalg.norm(q)) * rmax
        keys, keys_apx, seen, near = [], [], 0, []
        for j, l in enumerate(probes[qi]):
            if len(dec[l]) == 0:
                continue
            s_exact = dec[l].astype(np.float64) @ q.astype(np.float64)
            s_apx = _bf16(dec[l]).astype(np.float64) @ _bf16(q).astype(np.float64)
            worst = max(worst, float(np.abs(s_apx - s_exact).max()) / margin)
            if metric == 0:
                t1 = (dec[l].astype(np.float64) ** 2).sum(1) + 2 * dec[l].astype(np.float64) @ cent[l].astype(np.float64)
                base = float(((q - cent[l]).astype(np.float64) ** 2).sum())
                ke, ka = base + t1 - 2 * s_exact, base + t1 - 2 * s_apx
            else:
                base = -float(q.astype(np.float64) @ cent[l].astype(np.float64))
                ke, ka = base - s_exact, base - s_apx
            keys.append(ke)
            keys_apx.append(ka)
            if seen < 3000:          # phase A: nearest lists until enough codes were seen
                near.append(ke)
                seen += len(ke)
        keys, keys_apx = np.concatenate(keys), np.concatenate(keys_apx)
        near = np.sort(np.concatenate(near))
        assert len(near) >= k
        bound = near[k - 1]
        survivors = keys_apx <= bound + alpha * margin
        # every code at or under the bound survives the filter ...
        assert survivors[keys <= bound].all()
        # ... so the exact top-k of the survivors is the exact top-k of everything
        topk_all = np.sort(keys)[:k]
        topk_surv = np.sort(keys[survivors])[:k]
        assert np.array_equal(topk_all, topk_surv)
        # and the filter is selective (this is what makes the re-evaluation cheap)
        assert survivors.sum() <= 0.2 * len(keys)
    # the margin is a rigorous bound with room to spare (measured: ~0.1 of it)
    assert worst < 0.5, worst
