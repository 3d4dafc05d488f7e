"""faiss fourcc wire format, host side — no GPU needed: streams written by the compiled reference
(faiss::write_index: IxF2, IwFl, IwPQ, IxRF, IHNf) are parsed by kb2_fourcc.h; the parse is re-emitted with our writer and
handed BACK to the reference (faiss::read_index), which must search it exactly like the original."""
import numpy as np
import pytest

from knowhere_b200 import datagen


def _kb():
    import knowhere_b200 as kb
    if not hasattr(kb.lib(), "kb2_faiss_describe"):
        pytest.skip("library built without the fourcc API")
    return kb


def test_fourcc_flat_roundtrip(ref):
    kb = _kb()
    xb = datagen.clustered(500, 24, 1)
    xq = datagen.clustered(7, 24, 2)
    for metric, name in ((0, "L2"), (1, "IP")):
        blob = ref.flat_write(xb, metric)
        meta = kb.faiss_describe(blob)
        assert meta == {"type": "FLAT", "dim": 24, "rows": 500, "metric_type": name}
        again = kb.faiss_rewrite(blob)
        I0, D0, n0 = ref.read_and_search(blob, xq, 5)
        I1, D1, n1 = ref.read_and_search(again, xq, 5)
        assert n0 == n1 == 500 and np.array_equal(I0, I1) and np.array_equal(D0, D1)


@pytest.mark.parametrize("kind,m,refine", [("IVF_FLAT", 0, False), ("IVF_PQ", 8, False), ("IVF_PQ", 8, True)])
def test_fourcc_ivf_roundtrip(ref, kind, m, refine):
    kb = _kb()
    nb, d, nlist = 3000, 32, 16
    xb = datagen.clustered(nb, d, 3)
    xq = datagen.clustered(20, d, 4)
    r = ref.RefIvf(kind, d, 0, nlist, m, 8, refine=refine)
    r.train(xb)
    r.add(xb)
    blob = r.write()
    meta = kb.faiss_describe(blob)
    assert meta["type"] == kind and meta["dim"] == d and meta["rows"] == nb and meta["nlist"] == nlist
    if kind == "IVF_PQ":
        assert meta["m"] == m and meta["refine"] is refine
    again = kb.faiss_rewrite(blob)
    rk = 4.0 if refine else 0.0
    I0, D0, _ = ref.read_and_search(blob, xq, 10, nprobe=4, refine_k=rk)
    I1, D1, n1 = ref.read_and_search(again, xq, 10, nprobe=4, refine_k=rk)
    assert n1 == nb and np.array_equal(I0, I1) and np.array_equal(D0, D1)
    # and it is the search the live reference object gives
    I2, D2 = r.search(xq, 10, 4, refine_k=rk)
    assert np.array_equal(I0, I2)


def test_fourcc_hnsw_roundtrip(ref):
    kb = _kb()
    n, d, M = 1500, 16, 8
    xb = datagen.clustered(n, d, 5)
    h = ref.RefHnsw(d, M, 0, 40)
    h.add(xb)
    blob = h.write()
    meta = kb.faiss_describe(blob)
    assert meta["type"] == "HNSW" and meta["rows"] == n and meta["M"] == M
    again = kb.faiss_rewrite(blob)
    a, b = ref.hnsw_read_meta(blob, True), ref.hnsw_read_meta(again, True)
    g = h.export()
    assert a["entry_point"] == b["entry_point"] == g["entry_point"] and a["max_level"] == b["max_level"] == g["max_level"]
    assert np.array_equal(a["neighbors"], b["neighbors"]) and np.array_equal(b["neighbors"], g["neighbors"])
    assert np.array_equal(b["xb"], xb)


def test_fourcc_rejects_garbage(ref):
    kb = _kb()
    with pytest.raises(kb.KnowhereError) as e:
        kb.faiss_describe(b"IxZZ" + b"\0" * 64)
    assert e.value.status == 7            # not_implemented: unknown fourcc
    blob = ref.flat_write(datagen.clustered(50, 8, 1), 0)
    with pytest.raises(kb.KnowhereError) as e:
        kb.faiss_describe(blob[: len(blob) // 2])
    assert e.value.status == 19           # invalid_binary_set: truncated
