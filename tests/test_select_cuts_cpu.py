"""CPU models of the two histogram cuts added to the selection kernels (the guarantees the CUDA code relies on).

1. select_keys_hist_kernel fast path (kb2_flat.cuh): the slice is split into 256 disjoint chunks (one per thread), T is the upper
   edge of the bin of a 256-bin histogram of the CHUNK MINIMA at which the cumulative count reaches K.  Guarantee: at least K
   keys are <= T (so {keys <= T} contains the K smallest), whatever the data; on the coarse-stage geometry (best 80 of 4096)
   the set stays within the 128-entry slot, otherwise the kernel falls back to the level-wise histogram.
2. exact_eval_kernel trim (kb2_ivfpq_tc.cuh): 256 linear bins over [min, max] of a row's exact keys, keep every entry whose bin
   is <= the bin where the cumulative count reaches k'.  Guarantee: the k' smallest keys (with all their ties) are kept.
"""
import numpy as np

from knowhere_b200 import datagen


def f2ord(x):
    """order-preserving uint32 image of fp32 (kb2_common.cuh f2ord)"""
    u = np.asarray(x, np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def chunk_minimum_cut(keys, K):
    """returns T (ordered-int domain) as the kernel computes it, or None when it would fall back"""
    o = f2ord(keys).astype(np.uint64)
    n = len(o)
    pad = (-n) % 256
    op = np.concatenate([o, np.full(pad, 0xFFFFFFFF, np.uint64)]).reshape(-1, 256)   # thread t sees keys i = t (mod 256)
    cmin = op.min(0)
    finite = cmin < 0xFF800000      # image of +inf: chunks holding filtered keys only do not count
    if finite.sum() < K:
        return None
    vmin, ctop = int(cmin[finite].min()), int(cmin[finite].max())
    span = ctop - vmin
    shift = 0 if span < 256 else span.bit_length() - 8
    bins = ((cmin[finite] - vmin) >> shift).astype(np.int64)
    assert bins.max() <= 255
    hist = np.bincount(bins, minlength=256)
    b = int(np.searchsorted(np.cumsum(hist), K))          # first bin with cumulative count >= K
    return min(vmin + ((b + 1) << shift) - 1, ctop)


def test_chunk_minimum_cut_is_a_valid_bound():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.choice([512, 1000, 4096, 6656]))
        K = int(rng.choice([17, 48, 80, 128]))
        kind = trial % 4
        if kind == 0:
            keys = rng.standard_normal(n).astype(np.float32) * 10
        elif kind == 1:
            keys = -np.abs(rng.standard_cauchy(n)).astype(np.float32)     # inner-product keys with outliers
        elif kind == 2:
            keys = rng.integers(0, 5, n).astype(np.float32)               # heavy ties
        else:
            keys = rng.random(n).astype(np.float32)
            keys[rng.random(n) < 0.3] = np.inf                            # filtered entries
        T = chunk_minimum_cut(keys, K)
        if T is None:
            continue
        o = f2ord(keys).astype(np.uint64)
        emitted = o <= T
        assert emitted.sum() >= K
        kth = np.sort(o)[K - 1]
        assert (o[emitted].max() >= kth) and np.all(o[~emitted] > kth)     # everything left out is worse than the K-th best


def test_chunk_minimum_cut_fits_the_slot_at_the_coarse_geometry():
    """best 80 of 4096 centroid distances: the emitted set stays within the 128-entry slot (no fallback in the common case)"""
    xb = datagen.clustered(200_000, 128, 42)
    xq = datagen.clustered(200, 128, 43)
    rng = np.random.default_rng(1)
    cent = xb[rng.choice(len(xb), 4096, replace=False)]
    D = (xq ** 2).sum(1)[:, None] + (cent ** 2).sum(1)[None] - 2 * xq @ cent.T
    counts = []
    for row in D.astype(np.float32):
        T = chunk_minimum_cut(row, 80)
        counts.append(int((f2ord(row).astype(np.uint64) <= T).sum()))
    counts = np.array(counts)
    assert counts.min() >= 80
    assert (counts <= 128).mean() >= 0.99, counts.max()


def trim_cut(keys, k):
    """entries kept by the exact_eval trim: bin <= crossing bin of a 256-bin linear histogram over [min, max]"""
    keys = np.asarray(keys, np.float32)
    lo, hi = keys.min(), keys.max()
    sc = np.float32(256.0) / (hi - lo) if hi > lo else np.float32(0)
    bins = np.minimum(255, ((keys - lo) * sc).astype(np.int64))
    hist = np.bincount(bins, minlength=256)
    if hist.sum() < k:
        return np.ones(len(keys), bool)
    b = int(np.searchsorted(np.cumsum(hist), k))
    return bins <= b


def test_trim_cut_keeps_the_k_smallest_with_ties():
    rng = np.random.default_rng(2)
    for trial in range(300):
        n = int(rng.integers(41, 2048))
        k = int(rng.choice([10, 40, 96]))
        if n <= k:
            continue
        keys = (rng.standard_normal(n) * (1 + trial % 5)).astype(np.float32)
        if trial % 3 == 0:
            keys = np.round(keys, 1)                                      # ties
        keep = trim_cut(keys, k)
        kth = np.sort(keys)[k - 1]
        assert keep.sum() >= k
        assert np.all(keep[keys <= kth])                                   # the k smallest and every tie of the k-th
