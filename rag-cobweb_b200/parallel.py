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
    s.derive()  # var / tf rows from the received m2 / count


class ResultGather:
    """All-gather of the per-rank [q_r, k] results into the full batch order: ONE collective on one packed buffer.

    The shard sizes follow from shard_bounds(total, world, r) on every rank, so nothing is exchanged to learn them and
    nothing is read back to the host; ids (int32) and scores (float32 bit patterns) travel in one [rows, 2k] int32
    buffer.  Buffers are allocated once and re-used by every call."""

    def __init__(self, total, k, world=None, device="cuda"):
        self.world = world if world is not None else dist.get_world_size()
        self.k, self.total = k, total
        self.sizes = [shard_bounds(total, self.world, r)[1] - shard_bounds(total, self.world, r)[0] for r in range(self.world)]
        self.m = max(self.sizes) if self.sizes else 0
        self.send = torch.empty((self.m, 2 * k), dtype=torch.int32, device=device)
        self.recv = torch.empty((self.world * self.m, 2 * k), dtype=torch.int32, device=device)
        self.keep = None
        if min(self.sizes) != self.m:  # uneven shards: rows of the padded layout that are real
            self.keep = torch.cat([torch.arange(r * self.m, r * self.m + self.sizes[r], device=device)
                                   for r in range(self.world)])

    def __call__(self, ids, vals):
        if self.world == 1:
            return ids, vals
        n, k = ids.shape[0], self.k
        self.send[:n, :k] = ids
        self.send[:n, k:] = vals.view(torch.int32)
        dist.all_gather_into_tensor(self.recv, self.send)
        out = self.recv if self.keep is None else self.recv[self.keep]
        return out[:, :k], out[:, k:].view(torch.float32)


def gather_results(ids, vals, total):
    """One-off form of ResultGather: `total` = rows of the whole batch (this rank holds its shard_bounds share)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ids, vals
    return ResultGather(total, ids.shape[1], device=ids.device)(ids, vals)


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
    """All-gather [q, k] candidate lists of every rank along the candidate axis -> [q, world*k]; ids and score bits
    packed into one buffer, one collective."""
    world = dist.get_world_size()
    if world == 1:
        return ids, vals
    q, k = ids.shape
    send = torch.cat([ids, vals.view(torch.int32)], 1).contiguous()
    recv = torch.empty((world, q, 2 * k), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(recv.view(world * q, 2 * k), send)
    gi = recv[:, :, :k].permute(1, 0, 2).reshape(q, world * k)
    gv = recv[:, :, k:].permute(1, 0, 2).reshape(q, world * k).contiguous().view(torch.float32)
    return gi, gv
