"""Seeded synthetic data.

`uniform` reproduces the reference unit tests' GenDataSet distribution (tests/ut/utils.h:41-50:
uniform_real(0,100)); `clustered` is the structured low-rank mixture needed for PQ recall:
z ~ N(mu_c, I_r), x = z B + 0.05 eps."""
import numpy as np


def uniform(n, d, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, d), dtype=np.float32) * 100.0).astype(np.float32)


def clustered(n, d, seed, r=16, n_clusters=None, struct_seed=7):
    n_clusters = n_clusters or max(16, int(1000 * n / 200000))
    srng = np.random.default_rng(struct_seed)
    mu = srng.standard_normal((n_clusters, r)).astype(np.float32) * 3.0
    B = srng.standard_normal((r, d)).astype(np.float32)
    rng = np.random.default_rng(seed)
    c = rng.integers(0, n_clusters, n)
    z = mu[c] + rng.standard_normal((n, r)).astype(np.float32)
    x = z @ B + 0.05 * rng.standard_normal((n, d)).astype(np.float32)
    return np.ascontiguousarray(x, np.float32)


def clustered_torch(n, d, seed, device, r=16, n_clusters=None, struct_seed=7, chunk=1 << 20):
    """Same mixture generated on the GPU (torch is plumbing here: device memory + RNG)."""
    import torch
    n_clusters = n_clusters or max(16, int(1000 * n / 200000))
    g = torch.Generator(device=device)
    g.manual_seed(struct_seed)
    mu = torch.randn((n_clusters, r), generator=g, device=device) * 3.0
    B = torch.randn((r, d), generator=g, device=device)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        c = torch.randint(0, n_clusters, (m,), generator=g, device=device)
        z = mu[c] + torch.randn((m, r), generator=g, device=device)
        out[s:s + m] = z @ B + 0.05 * torch.randn((m, d), generator=g, device=device)
    return out


def recall(gt_ids, ids):
    """size(gt ∩ res) / (nq*k)  (reference tests/ut/utils.h:110-133)."""
    hit = 0
    for a, b in zip(gt_ids, ids):
        hit += len(set(a.tolist()) & set(b.tolist()) - {-1})
    return hit / float(gt_ids.shape[0] * gt_ids.shape[1])
