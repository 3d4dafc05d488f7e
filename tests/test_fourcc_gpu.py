"""faiss fourcc wire format on the GPU: a reference-written stream loads straight into the GPU index and answers like the
reference; a GPU-built index serialises to a stream the reference loads and searches with the same result."""
import numpy as np
import pytest

from knowhere_b200 import datagen
from tests.util import assert_topk_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,m,refine", [("IVF_FLAT", 0, False), ("IVF_PQ", 16, False), ("IVF_PQ", 16, True)])
def test_load_reference_stream_ivf(kb, ref, kind, m, refine, tmp_path):
    nb, d, nlist, nprobe, k = 20000, 64, 32, 8, 10
    xb = datagen.clustered(nb, d, 42)
    xq = datagen.clustered(100, d, 43)
    r = ref.RefIvf(kind, d, 0, nlist, m, 8, refine=refine)
    r.train(xb)
    r.add(xb)
    blob = r.write()
    ix = kb.Index.deserialize_faiss(blob)
    assert ix.count() == nb and ix.meta()["type"] == kind
    rk = 4 if refine else 1
    ids, dist = ix.search(xq, k, {"nprobe": nprobe, "refine_k": rk})
    I0, D0 = r.search(xq, k, nprobe, refine_k=float(rk) if refine else 0.0)
    assert_topk_parity(ids, dist, I0, D0, rtol=1e-4, atol=1e-3, what=f"{kind} loaded from a faiss stream", max_tie_rows=100)
    # file variant (DeserializeFromFile)
    f = tmp_path / "idx.bin"
    f.write_bytes(blob)
    ix2 = kb.Index.deserialize_from_file(str(f))
    ids2, _ = ix2.search(xq, k, {"nprobe": nprobe, "refine_k": rk})
    assert np.array_equal(ids, ids2)
    # GPU index -> faiss stream -> reference search
    out = ix.serialize_faiss()
    I1, D1, n1 = ref.read_and_search(out, xq, k, nprobe=nprobe, refine_k=float(rk) if refine else 0.0)
    assert n1 == nb
    assert_topk_parity(ids, dist, I1, D1, rtol=1e-4, atol=1e-3, what=f"{kind} written as a faiss stream", max_tie_rows=100)


def test_gpu_built_ivfpq_served_by_reference(kb, ref):
    nb, d, nlist, m = 30000, 64, 64, 16
    xb = datagen.clustered(nb, d, 1)
    xq = datagen.clustered(50, d, 2)
    ix = kb.Index("IVF_PQ", "L2", d, {"nlist": nlist, "m": m, "refine": True, "refine_type": "flat"})
    ix.build(xb)
    ids, dist = ix.search(xq, 10, {"nprobe": 16, "refine_k": 4})
    I1, D1, n1 = ref.read_and_search(ix.serialize_faiss(), xq, 10, nprobe=16, refine_k=4.0)
    assert n1 == nb
    assert_topk_parity(ids, dist, I1, D1, rtol=1e-4, atol=1e-3, what="GPU-built IVF_PQ+refine served by faiss", max_tie_rows=5)


def test_flat_and_hnsw_streams(kb, ref):
    n, d, M = 5000, 32, 8
    xb = datagen.clustered(n, d, 7)
    xq = datagen.clustered(40, d, 8)
    fi = kb.Index.deserialize_faiss(ref.flat_write(xb, 0))
    ids, dist = fi.search(xq, 5)
    I0, D0 = ref.flat_search(xb, xq, 5, 0)
    assert np.array_equal(ids, I0)
    I1, D1, _ = ref.read_and_search(fi.serialize_faiss(), xq, 5)
    assert np.array_equal(I1, I0)
    h = ref.RefHnsw(d, M, 0, 60)
    h.add(xb)
    hx = kb.Index.deserialize_faiss(h.write())
    ids, dist = hx.search(xq, 5, {"ef": 32})
    I2, D2, _ = h.search(xq, 5, 32)
    assert (ids == I2).all(1).mean() > 0.9
    back = ref.hnsw_read_meta(hx.serialize_faiss(), True)
    assert np.array_equal(back["neighbors"], h.export()["neighbors"]) and np.array_equal(back["xb"], xb)
