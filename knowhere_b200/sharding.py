"""Host-side logic of the multi-GPU path: list ownership, candidate packing for the single
all-gather, and the gather+merge step.  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the merge
itself is a callable (the CUDA merge kernel in production, a numpy merge in the gloo tests)."""
import numpy as np


def owner_of_list(list_no, world):
    """inverted list l lives on rank l % world (kb2_index_set_shard)"""
    return list_no % world


def pack_candidates(torch, ids, dist):
    """[nq,k] int64 ids + [nq,k] float32 distances -> one [nq,2k] int64 tensor (one collective, not two)"""
    k = ids.shape[1]
    out = torch.empty((ids.shape[0], 2 * k), dtype=torch.int64, device=ids.device)
    out[:, :k] = ids
    out[:, k:] = dist.contiguous().view(torch.int32).to(torch.int64)
    return out


def unpack_candidates(torch, gathered, k):
    """[world,nq,2k] -> ([world,nq,k] int64, [world,nq,k] float32)"""
    ids = gathered[:, :, :k].contiguous()
    dist = gathered[:, :, k:].to(torch.int32).contiguous().view(torch.float32)
    return ids, dist


def gather_and_merge(torch, dist_mod, ids, dist, merge_fn, world):
    """ONE all_gather of the packed per-shard top-k, then merge_fn([world,nq,k] ids, dist) -> ([nq,k],[nq,k])"""
    k = ids.shape[1]
    pack = pack_candidates(torch, ids, dist)
    if dist_mod.get_backend() == "nccl":
        gathered = torch.empty((world,) + tuple(pack.shape), dtype=torch.int64, device=pack.device)
        dist_mod.all_gather_into_tensor(gathered, pack)
    else:  # gloo (CPU tests): same collective, list flavour
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist_mod.all_gather(parts, pack)
        gathered = torch.stack(parts)
    g_ids, g_dist = unpack_candidates(torch, gathered, k)
    return merge_fn(g_ids, g_dist)


def merge_topk_numpy(ids, dist, metric):
    """reference merge used by the CPU tests: order by (distance, id), -1 entries last"""
    world, nq, k = ids.shape
    oi = np.full((nq, k), -1, np.int64)
    od = np.full((nq, k), np.finfo(np.float32).max if metric == "L2" else -np.finfo(np.float32).max, np.float32)
    for q in range(nq):
        cand = [(float(dist[w, q, j]) if metric == "L2" else -float(dist[w, q, j]), int(ids[w, q, j]))
                for w in range(world) for j in range(k) if ids[w, q, j] >= 0]
        cand.sort()
        for j, (key, i) in enumerate(cand[:k]):
            oi[q, j] = i
            od[q, j] = key if metric == "L2" else -key
    return oi, od
