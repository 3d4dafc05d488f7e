"""FLAT / BruteForce parity against the compiled reference (GPU).  Mirrors the reference's own
tests: tests/ut/test_bruteforce.cc:57-77 (self-query KAT) and tests/ut/test_search.cc:185-268."""
import numpy as np
import pytest

from knowhere_b200 import datagen
from tests.util import assert_topk_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["L2", "IP"])
def test_bruteforce_self_query_kat(kb, metric):
    # reference KAT: queries are the first nq base rows => ids[i][0]==i, L2 dist exactly 0
    xb = datagen.uniform(10000, 128, 42)
    if metric == "IP":
        xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = xb[:100].copy()
    ids, dist = kb.brute_force_search(xb, xq, 10, metric)
    assert (ids[:, 0] == np.arange(100)).all()
    if metric == "L2":
        assert (dist[:, 0] == 0.0).all()
    else:
        np.testing.assert_allclose(dist[:, 0], 1.0, atol=1e-5)


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nb,nq,d,k", [(10000, 100, 128, 10), (1000, 10, 128, 100), (5003, 37, 96, 7), (300, 5, 17, 3)])
def test_flat_matches_reference(kb, ref, metric, nb, nq, d, k):
    xb = datagen.uniform(nb, d, 42)
    xq = datagen.uniform(nq, d, 43)
    I0, D0 = ref.flat_search(xb, xq, k, metric)          # FLAT index path (baseline faiss SIMD)
    I1, D1 = ref.bruteforce_search(xb, xq, k, metric)    # BruteForce path (src/simd hooks)
    ix = kb.Index("FLAT", "L2" if metric == 0 else "IP", d)
    ix.add(xb)
    assert ix.count() == nb
    ids, dist = ix.search(xq, k)
    assert_topk_parity(ids, dist, I0, D0, what="FLAT vs IndexFlat")
    assert_topk_parity(ids, dist, I1, D1, what="FLAT vs BruteForce")
    ids2, dist2 = kb.brute_force_search(xb, xq, k, "L2" if metric == 0 else "IP")
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2)


def test_flat_fewer_rows_than_k(kb):
    xb = datagen.uniform(5, 16, 1)
    ix = kb.Index("FLAT", "L2", 16)
    ix.add(xb)
    ids, dist = ix.search(xb[:2].copy(), 8)
    assert (ids[:, 5:] == -1).all() and (dist[:, 5:] == np.finfo(np.float32).max).all()
    assert sorted(ids[0, :5].tolist()) == [0, 1, 2, 3, 4]


def test_flat_bitset(kb, ref):
    xb = datagen.uniform(4000, 64, 5)
    xq = datagen.uniform(20, 64, 6)
    mask = np.zeros(4000, bool)
    mask[::3] = True                      # filtered OUT
    bits = np.packbits(mask, bitorder="little")
    ix = kb.Index("FLAT", "L2", 64)
    ix.add(xb)
    ids, dist = ix.search(xq, 10, bitset=bits)
    assert not mask[ids].any()
    keep = np.nonzero(~mask)[0]
    I0, D0 = ref.flat_search(xb[keep], xq, 10, 0)
    assert_topk_parity(ids, dist, keep[I0], D0, what="FLAT bitset")


def test_flat_device_resident_io(kb):
    torch = pytest.importorskip("torch")
    xb = datagen.uniform(3000, 32, 9)
    xq = datagen.uniform(50, 32, 10)
    ix = kb.Index("FLAT", "IP", 32)
    ix.add(torch.from_numpy(xb).cuda())
    ids_h, dist_h = ix.search(xq, 5)
    ids_d, dist_d = ix.search(torch.from_numpy(xq).cuda(), 5)
    assert np.array_equal(ids_h, ids_d.cpu().numpy()) and np.array_equal(dist_h, dist_d.cpu().numpy())


def test_flat_range_search(kb, ref):
    xb = datagen.uniform(3000, 32, 11)
    xq = datagen.uniform(10, 32, 12)
    I, D = ref.flat_search(xb, xq, 20, 0)
    radius = float(np.median(D[:, -1]))
    lims0, ids0, dis0 = ref.flat_range_search(xb, xq, radius, 0)
    ix = kb.Index("FLAT", "L2", 32)
    ix.add(xb)
    lims, ids, dis = ix.range_search(xq, radius)
    assert np.array_equal(lims, lims0)
    for i in range(10):
        a = set(ids[lims[i]:lims[i + 1]].tolist())
        b = set(ids0[lims0[i]:lims0[i + 1]].tolist())
        assert a == b
        assert (np.diff(dis[lims[i]:lims[i + 1]]) >= 0).all()


def test_flat_serialize_roundtrip(kb):
    xb = datagen.uniform(2000, 24, 3)
    xq = datagen.uniform(8, 24, 4)
    ix = kb.Index("FLAT", "L2", 24)
    ix.add(xb)
    a = ix.search(xq, 5)
    ix2 = kb.Index.deserialize(ix.serialize())
    b = ix2.search(xq, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    np.testing.assert_array_equal(ix2.get_vector_by_ids([3, 1999]), xb[[3, 1999]])


@pytest.mark.parametrize("kind,cfg,scfg", [("FLAT", {}, {}), ("IVF_FLAT", {"nlist": 16}, {"nprobe": 16}),
                                           ("HNSW", {"M": 16, "efConstruction": 100}, {"ef": 64})])
def test_cosine_metric(kb, kind, cfg, scfg):
    # reference KAT (tests/ut/test_bruteforce.cc:57-77): self query under COSINE => first hit is itself, |dist-1| < 1e-5
    xb = datagen.clustered(4000, 48, 21) + 1.0
    ix = kb.Index(kind, "COSINE", 48, cfg)
    ix.build(xb)
    ids, dist = ix.search(xb[:50].copy(), 5, scfg)
    assert (ids[:, 0] == np.arange(50)).all()
    np.testing.assert_allclose(dist[:, 0], 1.0, atol=1e-5)
    assert (dist <= 1.0 + 1e-5).all() and (dist >= -1.0 - 1e-5).all()
    xn = xb / np.linalg.norm(xb, axis=1, keepdims=True)
    if kind != "HNSW":
        gt = np.argsort(-(xn[:50] @ xn.T), axis=1)[:, :5]
        assert (ids == gt).mean() > 0.98
    assert not ix.has_raw_data()
    bi, bd = kb.brute_force_search(xb, xb[:50].copy(), 5, "COSINE")
    assert (bi[:, 0] == np.arange(50)).all()
