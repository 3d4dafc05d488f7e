"""Secondary measurements (not the judged bench line): IVF_FLAT (BASELINE configs[1]) and HNSW (configs[3] shape,
scaled in n) with roofline + reference-CPU numbers, written as JSON lines for profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import knowhere_b200 as kb
from knowhere_b200 import datagen
from oracle import ref

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rec(gt, ids):
    return float(np.mean([len(set(a) & set(b)) for a, b in zip(gt, ids)]) / gt.shape[1])


def ivf_flat(n=1_000_000, d=128, nlist=1024, nprobe=32, nq=1000, k=10):
    xb = datagen.clustered_torch(n, d, 42, dev)
    xq = datagen.clustered_torch(nq, d, 43, dev)
    ix = kb.Index("IVF_FLAT", "L2", d, {"nlist": nlist})
    ix.set_stream(stream)
    t0 = time.time(); ix.build(xb); torch.cuda.synchronize(); tb = time.time() - t0
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dis = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cfg = {"nprobe": nprobe}
    ix.enable_kernel_timing(True)
    ms = timeit(lambda: ix.search(xq, k, cfg, out=(ids, dis)))
    kms, c = ix.last_kernel_ms(), ix.last_counters()
    gt, _ = kb.brute_force_search(xb, xq, k, "L2", stream=stream)
    out = dict(workload=f"IVF_FLAT L2 {n}x{d} nlist={nlist} nprobe={nprobe} batch={nq} k={k}", qps=nq / ms * 1e3, ms_per_batch=ms,
               scan_kernel_ms=kms, algorithmic_bytes=c["code_bytes"], achieved_GBps=c["code_bytes"] / kms / 1e6,
               frac_of_hbm_peak=c["code_bytes"] / kms / 1e6 / PEAK, recall_at_10=rec(gt.cpu().numpy(), ids.cpu().numpy()), build_s=tb,
               launches=c["launches"])
    # reference CPU on the same index
    r = ref.RefIvf("IVF_FLAT", d, 0, nlist)
    cent, _ = ix.ivf_export_centroids(0)
    r.import_state(cent, None, ((l,) + ix.ivf_export_list(l, d * 4) for l in range(nlist)))
    xq_h = xq.cpu().numpy()
    r.search(xq_h[:64], k, nprobe)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); I, D = r.search(xq_h, k, nprobe); ts.append(time.perf_counter() - t0)
    out.update(cpu_reference_qps=nq / np.median(ts), cpu_threads=os.cpu_count(), cpu_recall=rec(gt.cpu().numpy(), I),
               ids_equal_to_cpu=float((ids.cpu().numpy() == I).mean()))
    print(json.dumps(out), flush=True)


def hnsw(n=100_000, d=768, M=16, efc=200, ef=128, nq=1000, k=10):
    xb = datagen.clustered(n, d, 42)
    xq = datagen.clustered(nq, d, 43)
    t0 = time.time()
    h = ref.RefHnsw(d, M, 1, efc); h.add(xb); tb = time.time() - t0
    g = h.export()
    ix = kb.Index("HNSW", "IP", d, {"M": M, "efConstruction": efc})
    ix.set_stream(stream)
    ix.hnsw_import(xb, g["levels"], g["offsets"], g["neighbors"], g["cum"], g["entry_point"], g["max_level"])
    xq_d = torch.from_numpy(xq).to(dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dis = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cfg = {"ef": ef}
    ix.enable_kernel_timing(True)
    ms = timeit(lambda: ix.search(xq_d, k, cfg, out=(ids, dis)))
    kms, c = ix.last_kernel_ms(), ix.last_counters()
    ndis, nhops = ix.hnsw_last_stats()
    gt, _ = kb.brute_force_search(xb, xq, k, "IP")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); I, D, st = h.search(xq, k, ef); ts.append(time.perf_counter() - t0)
    out = dict(workload=f"HNSW IP {n}x{d} M={M} efConstruction={efc} ef={ef} batch={nq} k={k} (graph built by the reference)",
               qps=nq / ms * 1e3, ms_per_batch=ms, kernel_ms=kms, ndis_per_query=ndis / nq, nhops_per_query=nhops / nq,
               algorithmic_bytes=c["code_bytes"], achieved_GBps=c["code_bytes"] / kms / 1e6,
               frac_of_hbm_peak=c["code_bytes"] / kms / 1e6 / PEAK, recall_at_10=rec(gt, ids.cpu().numpy()),
               cpu_reference_qps=nq / np.median(ts), cpu_threads=os.cpu_count(), cpu_recall=rec(gt, I),
               cpu_ndis_per_query=st[0] / nq, rows_identical_to_cpu=float((ids.cpu().numpy() == I).all(1).mean()),
               reference_build_s=tb)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "ivf_flat"):
        ivf_flat()
    if which in ("all", "hnsw"):
        hnsw()
