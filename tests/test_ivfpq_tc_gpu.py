"""List-major tensor-core engine of the IVF_PQ scan (kb2_ivfpq_tc.cuh) against the query-major LUT engine.

The tensor-core contraction is only a filter: survivors are re-evaluated with the LUT kernel's own fp32 operations,
so both engines must return the same (id, distance) rows bit for bit; the LUT engine itself is pinned to the
compiled reference in test_ivf_gpu.py, and here the tc engine is compared with the reference directly as well."""
import os

import numpy as np
import pytest

from knowhere_b200 import datagen
from tests.util import assert_topk_parity

pytestmark = pytest.mark.gpu


def _search(ix, xq, k, cfg, engine, bitset=None):
    old = os.environ.get("KB2_PQ_ENGINE")
    os.environ["KB2_PQ_ENGINE"] = engine
    try:
        return ix.search(xq, k, cfg, bitset=bitset) if bitset is not None else ix.search(xq, k, cfg)
    finally:
        if old is None:
            os.environ.pop("KB2_PQ_ENGINE", None)
        else:
            os.environ["KB2_PQ_ENGINE"] = old


def _build(kb, metric, nb, nlist, refine=False, seed=42):
    xb = datagen.clustered(nb, 128, seed)
    ix = kb.Index("IVF_PQ", metric, 128, {"nlist": nlist, "m": 16, "nbits": 8, "refine": refine, "refine_type": "flat"})
    ix.train(xb[: min(nb, 40000)])
    ix.add(xb)
    return ix, xb


@pytest.mark.parametrize("metric", ["L2", "IP"])
@pytest.mark.parametrize("nb,nlist,nprobe,nq,k", [(60000, 64, 16, 3000, 10), (30000, 32, 32, 1000, 40), (8000, 128, 64, 700, 10)])
def test_ivfpq_tc_engine_matches_lut_engine(kb, metric, nb, nlist, nprobe, nq, k):
    ix, xb = _build(kb, metric, nb, nlist)
    xq = datagen.clustered(nq, 128, 43)
    cfg = {"nprobe": nprobe}
    i0, d0 = _search(ix, xq, k, cfg, "lut")
    i1, d1 = _search(ix, xq, k, cfg, "tc")
    c = ix.last_counters()
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32)), f"distances differ in {(d0 != d1).any(axis=1).sum()} rows"
    assert np.array_equal(i0, i1), f"ids differ in {(i0 != i1).any(axis=1).sum()} rows"
    # the tc pass really ran: it reports the (query, code) pairs it filtered
    assert c["codes"] > 0


def test_ivfpq_tc_engine_refine_bitset_and_reference(kb, ref):
    nb, nlist, nprobe, nq, k = 40000, 64, 16, 2000, 10
    xb = datagen.clustered(nb, 128, 7)
    xq = datagen.clustered(nq, 128, 8)
    r = ref.RefIvf("IVF_PQ", 128, 0, nlist, 16, 8, refine=True)
    r.train(xb)
    r.add(xb)
    ix = kb.Index("IVF_PQ", "L2", 128, {"nlist": nlist, "m": 16, "nbits": 8, "refine": True, "refine_type": "flat"})
    ix.ivf_import(r.centroids(), r.pq_centroids(), list(r.lists()), raw=xb)
    cfg = {"nprobe": nprobe, "refine_k": 4.0}
    I0, D0 = r.search(xq, k, nprobe, refine_k=4.0)
    i1, d1 = _search(ix, xq, k, cfg, "tc")
    assert_topk_parity(i1, d1, I0, D0, rtol=1e-4, atol=1e-3, what="IVF_PQ tc engine + refine", max_tie_rows=nq // 20)
    # bitset on a GPU-built index (labels == insertion rows): every third row filtered out
    ix2, _ = _build(kb, "L2", nb, nlist, refine=True, seed=7)
    mask = np.zeros(nb, bool)
    mask[::3] = True
    bits = np.packbits(mask, bitorder="little")
    a = _search(ix2, xq, k, cfg, "lut", bitset=bits)
    b = _search(ix2, xq, k, cfg, "tc", bitset=bits)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert not mask[b[0][b[0] >= 0]].any()


def test_ivfpq_tc_engine_flagged_queries_fall_back(kb):
    # tiny lists: the two nearest lists hold fewer than k codes for many queries -> no bound -> those queries are
    # redone by the LUT kernel inside the same call; results must still be identical
    ix, xb = _build(kb, "L2", 6000, 512, seed=5)
    xq = datagen.clustered(600, 128, 6)
    cfg = {"nprobe": 64}
    i0, d0 = _search(ix, xq, 40, cfg, "lut")
    os.environ["KB2_TC_P0"] = "1"     # phase A may only look at the nearest list (~12 codes): no bound for most queries
    try:
        i1, d1 = _search(ix, xq, 40, cfg, "tc")
    finally:
        os.environ.pop("KB2_TC_P0", None)
    c = ix.last_counters()
    assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert c["flagged"] > 0


def test_ivfpq_tc_engine_on_shards(kb):
    """List sharding (kb2_index_set_shard): lists of the other shard have length 0 — no work items, phase A walks further
    down the probe list.  Each shard must answer identically through both engines, and the merged shards must equal the
    unsharded search."""
    nb, nlist, nq, k, world = 40000, 64, 2000, 10, 2
    xb = datagen.clustered(nb, 128, 21)
    xq = datagen.clustered(nq, 128, 22)
    full, _ = _build(kb, "L2", nb, nlist, seed=21)
    cent, pq = full.ivf_export_centroids(16)
    cfg = {"nprobe": 16}
    I0, D0 = _search(full, xq, k, cfg, "tc")
    ids, dis = [], []
    for rank in range(world):
        sh = kb.Index("IVF_PQ", "L2", 128, {"nlist": nlist, "m": 16, "nbits": 8})
        sh.set_shard(rank, world)
        kb._check(kb.lib().kb2_ivf_import_begin(sh.h, nlist, cent.ctypes.data, pq.ctypes.data))
        sh.add(xb)
        a = _search(sh, xq, k, cfg, "lut")
        b = _search(sh, xq, k, cfg, "tc")
        assert sh.last_counters()["codes"] > 0
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), f"shard {rank}"
        ids.append(b[0]); dis.append(b[1])
    mi, md = kb.merge_topk(np.stack(ids), np.stack(dis), "L2")
    assert_topk_parity(mi, md, I0, D0, rtol=1e-6, atol=1e-6, what="sharded tc merge", max_tie_rows=nq // 10)


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_ivfpq_tc_engine_m48_dsub2(kb, ref, metric):
    """The <G=3, dsub=2> instance of the engine (m = 48, d = 96: BASELINE C5's geometry, IP and L2): bit-identical to the LUT
    engine, and the reference's IndexIVFPQ on the same imported index agrees (IVFPQScanner_impl.h:110-185)."""
    nb, d, m, nlist, nprobe, nq, k = 60000, 96, 48, 64, 16, 3000, 10
    xb = datagen.clustered(nb, d, 42)
    xq = datagen.clustered(nq, d, 43)
    mt = 0 if metric == "L2" else 1
    r = ref.RefIvf("IVF_PQ", d, mt, nlist, m, 8)
    r.train(xb)
    r.add(xb)
    ix = kb.Index("IVF_PQ", metric, d, {"nlist": nlist, "m": m, "nbits": 8})
    ix.ivf_import(r.centroids(), r.pq_centroids(), list(r.lists()))
    cfg = {"nprobe": nprobe}
    ix.enable_kernel_timing(True)
    i1, d1 = _search(ix, xq, k, cfg, "tc")
    assert ix.last_stage_info()["engine"] == "tc" and ix.last_counters()["codes"] > 0
    i0, d0 = _search(ix, xq, k, cfg, "lut")
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32)), f"distances differ in {(d0 != d1).any(axis=1).sum()} rows"
    assert np.array_equal(i0, i1)
    I0, D0 = r.search(xq, k, nprobe)
    assert_topk_parity(i1, d1, I0, D0, rtol=1e-4, atol=1e-3, what=f"IVF_PQ m48 {metric} tc engine", max_tie_rows=nq // 10)
    # int8-valued data (C5 is int8 widened to fp32) through the typed entry points, with a bitset
    s = 127.0 / np.abs(xb).max()
    xb8 = np.clip(np.round(xb * s), -127, 127).astype(np.int8)
    xq8 = np.clip(np.round(xq * s), -127, 127).astype(np.int8)
    ix8 = kb.Index("IVF_PQ", metric, d, {"nlist": nlist, "m": m, "nbits": 8})
    ix8.build(xb8)
    mask = np.zeros(nb, bool)
    mask[::4] = True
    bits = np.packbits(mask, bitorder="little")
    a = _search(ix8, xq8, k, cfg, "lut", bitset=bits)
    b = _search(ix8, xq8, k, cfg, "tc", bitset=bits)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert not mask[b[0][b[0] >= 0]].any()
