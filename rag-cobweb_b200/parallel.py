"""Multi-GPU plumbing for batched predict: one process per GPU,
torch.distributed over NCCL (gloo in the CPU tests).

The path shards over independent units (queries): the node store is replicated (one broadcast
after build), each rank answers a contiguous shard of the batch, and the [Q/g, k] results come
back with one all-gather -- no reduction on the data path.  For stores beyond one HBM the
sentences (and the nodes on their paths) are partitioned instead and per-rank top-k lists are
merged after the same all-gather (`merge_topk`).  ifit does not shard (every insert updates the
root): replicas only.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_store(tree, src=0):
    """Replicate the node store of rank `src` on every rank (tensors are reallocated to the
    source's capacity first)."""
    s = tree.store
    dev = s.hdr.device
    meta = torch.tensor([s.cap, s.pool_cap], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src)
    cap, pool_cap = int(meta[0].item()), int(meta[1].item())
    if (cap, pool_cap) != (s.cap, s.pool_cap):
        s.cap = 0  # drop contents, reallocate at the source's size
        s.pool_cap = 0
        s._alloc(cap, pool_cap)
        s.cap, s.pool_cap = cap, pool_cap
    for name in ("hdr", "mean", "m2", "count", "parent", "child_off", "child_cnt", "child_cap", "child_pool", "n_sent",
                 "free_list"):
        dist.broadcast(getattr(s, name), src)
    s._struct = None


def gather_results(ids, vals, counts=None):
    """All-gather per-rank [q_r, k] results into the full batch order.  Shards may differ in
    size by one row (shard_bounds): rows are padded to the largest shard for the collective."""
    world = dist.get_world_size()
    if world == 1:
        return ids, vals
    q_r = torch.tensor([ids.shape[0]], dtype=torch.int64, device=ids.device)
    sizes = [torch.zeros_like(q_r) for _ in range(world)]
    dist.all_gather(sizes, q_r)
    sizes = [int(s.item()) for s in sizes]
    m, k = max(sizes), ids.shape[1]
    pid = torch.full((m, k), -1, dtype=ids.dtype, device=ids.device)
    pva = torch.full((m, k), float("-inf"), dtype=vals.dtype, device=vals.device)
    pid[: ids.shape[0]], pva[: vals.shape[0]] = ids, vals
    gi = torch.empty((world * m, k), dtype=ids.dtype, device=ids.device)
    gv = torch.empty((world * m, k), dtype=vals.dtype, device=vals.device)
    dist.all_gather_into_tensor(gi, pid)
    dist.all_gather_into_tensor(gv, pva)
    keep = torch.cat([torch.arange(r * m, r * m + sizes[r], device=ids.device) for r in range(world)])
    return gi[keep], gv[keep]


def merge_topk(ids, vals, k):
    """Merge candidate lists [q, c] (c = ranks * k) into the k best per query by
    (score desc, id asc) -- the order the single-GPU top-k uses."""
    key = torch.where(ids >= 0, vals, torch.full_like(vals, float("-inf")))
    # stable two-pass sort: by id, then by score
    o1 = torch.argsort(ids.to(torch.int64) + (ids < 0) * (1 << 40), dim=1, stable=True)
    key1 = torch.gather(key, 1, o1)
    o2 = torch.argsort(-key1, dim=1, stable=True)
    order = torch.gather(o1, 1, o2)[:, :k]
    return torch.gather(ids, 1, order), torch.gather(key, 1, order)


def gather_candidates(ids, vals):
    """All-gather [q, k] candidate lists of every rank along the candidate axis -> [q, world*k]."""
    world = dist.get_world_size()
    if world == 1:
        return ids, vals
    gi = [torch.empty_like(ids) for _ in range(world)]
    gv = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(gi, ids)
    dist.all_gather(gv, vals)
    return torch.cat(gi, 1), torch.cat(gv, 1)


def partition_sentences(leaf_row_sorted_pos, world, rank):
    """Sentence positions (already in tree order) owned by `rank` in the store-sharded mode."""
    return shard_bounds(len(leaf_row_sorted_pos), world, rank)
