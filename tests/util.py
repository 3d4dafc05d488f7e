import numpy as np


def assert_topk_parity(ids, dist, ref_ids, ref_dist, rtol=1e-4, atol=1e-6, what="", max_tie_rows=None):
    """Parity bar of BASELINE north_star: identical top-k id sets; distances within 1e-4 relative.
    An id mismatch is tolerated only if it is a tie at the k-th boundary within the distance
    tolerance (fp32 summation order differs between the CPU SIMD level and the GPU)."""
    ids, dist, ref_ids, ref_dist = map(np.asarray, (ids, dist, ref_ids, ref_dist))
    assert ids.shape == ref_ids.shape, what
    nq, k = ids.shape
    bad = 0
    for i in range(nq):
        a, b = ids[i], ref_ids[i]
        if np.array_equal(a, b):
            np.testing.assert_allclose(dist[i], ref_dist[i], rtol=rtol, atol=atol, err_msg=f"{what} q{i}")
            continue
        sa, sb = set(a.tolist()), set(b.tolist())
        if sa == sb:
            # same set, order differs only among (near-)equal distances
            np.testing.assert_allclose(np.sort(dist[i]), np.sort(ref_dist[i]), rtol=rtol, atol=atol,
                                       err_msg=f"{what} q{i}")
            continue
        # boundary swap: every id that differs must sit at the k-th distance within tolerance
        kth = ref_dist[i][-1]
        tol = rtol * max(abs(kth), 1.0) + atol
        for x in sa ^ sb:
            d = dist[i][list(a).index(x)] if x in sa else ref_dist[i][list(b).index(x)]
            assert abs(d - kth) <= tol * 4, f"{what} q{i}: id {x} differs and is not a boundary tie ({d} vs kth {kth})"
        bad += 1
    lim = max(1, nq // 50) if max_tie_rows is None else max_tie_rows
    assert bad <= lim, f"{what}: too many boundary-tie rows ({bad}/{nq})"


def recall_at_k(gt_ids, ids):
    hit = 0
    for a, b in zip(gt_ids, ids):
        hit += len((set(a.tolist()) & set(b.tolist())) - {-1})
    return hit / float(gt_ids.shape[0] * gt_ids.shape[1])
