"""Flat structure-of-arrays node store resident in HBM.

Replaces the reference's one-Python-object-per-concept representation
(src/cobweb/CobwebTorchNode.py:31-55) by torch CUDA tensors that the sm_100a kernels address
through the plain-C ``cw_store`` struct (include/cobweb_b200.h).  torch is plumbing here:
allocation, growth, host<->device copies.  All arithmetic on the arrays happens in the kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class NodeStore:
    def __init__(self, d, prior_var, flags, cap=4096, pool_cap=None, device=None):
        _lib.require_cuda()
        if not (1 <= d <= _lib.MAX_D):
            raise ValueError(f"embedding dim {d} outside 1..{_lib.MAX_D}")
        self.d = int(d)
        self.prior_var = float(prior_var)
        self.flags = int(flags)
        self.device = torch.device(device if device is not None else "cuda")
        self.cap = 0
        self.pool_cap = 0
        self._alloc(max(int(cap), 1024), int(pool_cap or 0))
        self.clear()

    # ------------------------------------------------------------------ allocation
    def _alloc(self, cap, pool_cap):
        dev, d = self.device, self.d
        pool_cap = max(pool_cap, 8 * cap + 4 * _lib.IFIT_POOL_SLACK)
        new = dict(
            mean=torch.zeros((cap, d), dtype=torch.float32, device=dev),
            m2=torch.zeros((cap, d), dtype=torch.float32, device=dev),
            # derived rows (compute_var / its log), kept in step with m2 / count by the ifit kernel
            var=torch.zeros((cap, d), dtype=torch.float32, device=dev),
            tf=torch.zeros((cap, d), dtype=torch.float32, device=dev),
            count=torch.zeros(cap, dtype=torch.float32, device=dev),
            parent=torch.full((cap,), -2, dtype=torch.int32, device=dev),
            child_off=torch.zeros(cap, dtype=torch.int32, device=dev),
            child_cnt=torch.zeros(cap, dtype=torch.int32, device=dev),
            child_cap=torch.zeros(cap, dtype=torch.int32, device=dev),
            n_sent=torch.zeros(cap, dtype=torch.int32, device=dev),
            free_list=torch.zeros(cap, dtype=torch.int32, device=dev),
            child_pool=torch.zeros(pool_cap, dtype=torch.int32, device=dev),
        )
        if self.cap:
            n = self.cap
            for k in ("mean", "m2", "var", "tf", "count", "parent", "child_off", "child_cnt", "child_cap", "n_sent", "free_list"):
                new[k][:n] = getattr(self, k)
            new["child_pool"][: self.pool_cap] = self.child_pool
        else:
            self.hdr = torch.zeros(_lib.HDR_WORDS, dtype=torch.int32, device=dev)
            self.scratch = torch.zeros(_lib.SCRATCH_WORDS, dtype=torch.int32, device=dev)
        for k, v in new.items():
            setattr(self, k, v)
        self.cap, self.pool_cap = cap, pool_cap
        self._struct = None

    def struct(self):
        if self._struct is None:
            s = _lib.CwStore()
            s.D, s.cap, s.pool_cap, s.flags, s.prior_var = self.d, self.cap, self.pool_cap, self.flags, self.prior_var
            for k in ("mean", "m2", "count", "parent", "child_off", "child_cnt", "child_cap", "child_pool", "n_sent",
                      "free_list", "hdr", "scratch", "var", "tf"):
                setattr(s, k, getattr(self, k).data_ptr())
            self._struct = s
        return C.byref(self._struct)

    def clear(self):
        """CobwebTorchTree.clear (CobwebTorchTree.py:43-50)."""
        _lib.check(_lib.load().cw_store_init(self.struct(), _lib.stream_ptr()), "cw_store_init")

    # ------------------------------------------------------------------ header
    def header(self):
        """Synchronises the stream and returns the header words as numpy int32."""
        return self.hdr.cpu().numpy()

    @staticmethod
    def _u64(h, lo):
        return (int(np.uint32(h[lo + 1])) << 32) | int(np.uint32(h[lo]))

    def counters(self):
        h = self.header()
        return dict(scores=self._u64(h, _lib.HDR_N_SCORES), rows=self._u64(h, _lib.HDR_N_ROWS),
                    levels=self._u64(h, _lib.HDR_N_LEVELS))

    def ifit_phase_cycles(self):
        """Cycles the lead CTA spent per ifit phase since the store was created (cw_ifit.cu MARK()).  All zero unless the
        library was built with the timers (`CW_IFIT_FINE_TIMERS=1 python rag-cobweb_b200/build.py --force`): they sit on
        the critical path of every level-step and cost ~2 % of the insert rate."""
        base = 16 + 4 * _lib.MAX_CHILDREN + 1 + 3
        w = self.scratch[base:base + 20].cpu().numpy().view(np.int64)
        names = ["apply", "S1", "lists+slices", "phaseA", "S2", "decA", "phaseB", "S3", "decB", "-"]
        return dict(zip(names, w.tolist()))

    @property
    def root(self):
        return int(self.header()[_lib.HDR_ROOT])

    def reserve(self, n_inserts, h=None):
        """Make room for n_inserts more instances (amortised doubling; compacts the child pool
        when most of it is leaked chunks)."""
        h = self.header() if h is None else h
        n_used, pool_used, max_child = int(h[_lib.HDR_N_USED]), int(h[_lib.HDR_POOL_USED]), int(h[_lib.HDR_MAX_CHILD])
        free_top = int(h[_lib.HDR_FREE_TOP])
        need_nodes = n_used - free_top + 2 * n_inserts + 2 * _lib.IFIT_NODE_SLACK
        need_pool = pool_used + 24 * n_inserts + 2 * _lib.IFIT_POOL_SLACK + 16 * max_child
        if need_pool > self.pool_cap and pool_used > 0:
            self.compact_pool()
            pool_used = int(self.header()[_lib.HDR_POOL_USED])
            need_pool = pool_used + 24 * n_inserts + 2 * _lib.IFIT_POOL_SLACK + 16 * max_child
        cap, pool_cap = self.cap, self.pool_cap
        if need_nodes > cap:
            cap = max(need_nodes, int(cap * 1.5))
        if need_pool > pool_cap:
            pool_cap = max(need_pool, int(pool_cap * 1.5))
        if cap != self.cap or pool_cap != self.pool_cap:
            self._alloc(cap, pool_cap)

    def compact_pool(self):
        """Rewrite child lists contiguously (drops chunks leaked by list growth)."""
        h = self.header()
        n_used = int(h[_lib.HDR_N_USED])
        cnt = self.child_cnt[:n_used].to(torch.int64)
        alive = self.parent[:n_used] > -2
        cnt = torch.where(alive, cnt, torch.zeros_like(cnt))
        capn = torch.where(cnt > 0, torch.clamp(2 * cnt, min=4), torch.zeros_like(cnt))
        new_off = torch.cumsum(capn, 0) - capn
        total = int(capn.sum().item())
        src_start = torch.repeat_interleave(self.child_off[:n_used].to(torch.int64), cnt)
        dst_start = torch.repeat_interleave(new_off, cnt)
        within = torch.arange(int(cnt.sum().item()), device=self.device) - torch.repeat_interleave(
            torch.cumsum(cnt, 0) - cnt, cnt)
        new_pool = torch.zeros_like(self.child_pool)
        new_pool[dst_start + within] = self.child_pool[src_start + within]
        self.child_pool = new_pool
        self.child_off[:n_used] = new_off.to(torch.int32)
        self.child_cap[:n_used] = capn.to(torch.int32)
        self.hdr[_lib.HDR_POOL_USED] = total
        self._struct = None

    # ------------------------------------------------------------------ host copies
    def topology(self):
        """Host copy of the topology arrays (one sync): dict of numpy arrays."""
        h = self.header()
        n_used, pool_used = int(h[_lib.HDR_N_USED]), int(h[_lib.HDR_POOL_USED])
        return dict(
            root=int(h[_lib.HDR_ROOT]), n_used=n_used,
            parent=self.parent[:n_used].cpu().numpy(), count=self.count[:n_used].cpu().numpy(),
            child_off=self.child_off[:n_used].cpu().numpy(), child_cnt=self.child_cnt[:n_used].cpu().numpy(),
            child_pool=self.child_pool[: max(pool_used, 1)].cpu().numpy(), n_sent=self.n_sent[:n_used].cpu().numpy(),
        )

    def rows(self, ids):
        idx = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=self.device)
        return self.mean[idx].cpu().numpy(), self.m2[idx].cpu().numpy()

    def load_arrays(self, parent, count, n_sent, mean, m2):
        """Replace the contents by nodes in an order where parent[i] < i and siblings appear in
        child-list order (BFS / pre-order dumps).  Node i gets id i."""
        parent = np.asarray(parent, np.int64)
        n = len(parent)
        self.reserve(0, h=np.array([0, n] + [0] * (_lib.HDR_WORDS - 2), np.int32))
        cnt = np.bincount(parent[parent >= 0], minlength=n).astype(np.int64)
        capn = np.where(cnt > 0, np.maximum(2 * cnt, 4), 0)
        off = np.cumsum(capn) - capn
        if int(capn.sum()) + 4 * _lib.IFIT_POOL_SLACK > self.pool_cap:
            self._alloc(self.cap, int(capn.sum()) + 4 * _lib.IFIT_POOL_SLACK)
        pool = np.zeros(self.pool_cap, np.int32)
        kids = np.nonzero(parent >= 0)[0]
        order = kids[np.argsort(parent[kids], kind="stable")]  # children grouped by parent, in id order
        par_sorted = parent[order]
        within = np.arange(len(order)) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        pool[off[par_sorted] + within] = order
        dev = self.device
        for lo in range(0, n, 1 << 16):  # in row chunks: the sources may be memory-mapped snapshots of many GB
            hi = min(n, lo + (1 << 16))
            self.mean[lo:hi] = torch.from_numpy(np.array(mean[lo:hi], dtype=np.float32, copy=True)).to(dev)
            self.m2[lo:hi] = torch.from_numpy(np.array(m2[lo:hi], dtype=np.float32, copy=True)).to(dev)
        self.count[:n] = torch.as_tensor(np.asarray(count, np.float32), device=dev)
        self.parent[:n] = torch.as_tensor(parent.astype(np.int32), device=dev)
        self.child_off[:n] = torch.as_tensor(off.astype(np.int32), device=dev)
        self.child_cnt[:n] = torch.as_tensor(cnt.astype(np.int32), device=dev)
        self.child_cap[:n] = torch.as_tensor(capn.astype(np.int32), device=dev)
        self.n_sent[:n] = torch.as_tensor(np.asarray(n_sent if n_sent is not None else np.zeros(n), np.int32), device=dev)
        self.child_pool.copy_(torch.as_tensor(pool, device=dev))
        hdr = np.zeros(_lib.HDR_WORDS, np.int32)
        hdr[_lib.HDR_ROOT], hdr[_lib.HDR_N_USED], hdr[_lib.HDR_POOL_USED] = 0, n, int(capn.sum())
        hdr[_lib.HDR_MAX_CHILD] = int(cnt.max()) if n else 0
        self.hdr.copy_(torch.as_tensor(hdr, device=dev))
        self.derive(n)

    def derive(self, n=None):
        """Rebuild the derived rows (var, tf) of node rows [0, n) after mean / m2 / count were written from outside."""
        n = int(self.header()[_lib.HDR_N_USED]) if n is None else int(n)
        _lib.check(_lib.load().cw_store_derive(self.struct(), n, _lib.stream_ptr()), "cw_store_derive")

    def bytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in ("mean", "m2", "var", "tf", "count", "parent", "child_off", "child_cnt", "child_cap", "child_pool",
                             "n_sent", "free_list"))
