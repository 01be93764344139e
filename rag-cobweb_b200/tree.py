"""CobwebTorchTree with the reference's signatures (src/cobweb/CobwebTorchTree.py), backed by
the HBM node store and the sm_100a kernels of libcobweb_b200.so.

Reference surface kept: __init__(shape, use_info, acuity_cutoff, use_kl, prior_var, alpha,
device) :23, clear :43, ifit :123, categorize :291, compute_var :336, compute_score :344,
dump_json :67, load_json :94, analyze_structure :366, attributes root / shape / device /
prior_var.  Additive: ifit_batch, categorize_batch (the batched entry points the kernels are
built for).
"""
import json
import math

import numpy as np
import torch

from . import _lib, constants, serialize, topology
from .store import NodeStore

OP_NAMES = ["best", "new", "merge", "split", "leaf", "fringe"]


def default_prior_var():
    """1 / (2 e pi) as the reference evaluates it (CobwebTorchTree.py:35-40): python double
    2*e times the fp32 pi tensor, reciprocal in fp32."""
    return float(np.float32(1.0) / (np.float32(2 * math.e) * np.float32(math.pi)))


class SentenceList(list):
    """node.sentence_id: a list whose growth is mirrored into the store's n_sent counter, which
    is what the best-first kernel tests for `if curr.sentence_id` (CobwebTorchTree.py:267)."""

    def __init__(self, tree, node_id, items=()):
        super().__init__(items)
        self._tree, self._node = tree, node_id

    def append(self, sid):
        super().append(sid)
        self._tree._note_sentence(self._node, sid)


class CobwebNode:
    """Handle onto one row of the node store (stands in for CobwebTorchNode objects,
    src/cobweb/CobwebTorchNode.py:9).  Reads go to the device on access."""

    __slots__ = ("tree", "node_id")

    def __init__(self, tree, node_id):
        self.tree, self.node_id = tree, int(node_id)

    def __eq__(self, other):
        return isinstance(other, CobwebNode) and other.tree is self.tree and other.node_id == self.node_id

    def __hash__(self):
        return hash(("CobwebNode", id(self.tree), self.node_id))

    def __repr__(self):
        return f"CobwebNode(id={self.node_id})"

    @property
    def id(self):
        return str(self.node_id)

    @property
    def concept_id(self):
        return self.node_id

    @property
    def count(self):
        return self.tree.store.count[self.node_id].clone()

    @property
    def mean(self):
        return self.tree.store.mean[self.node_id].clone()

    @property
    def meanSq(self):
        return self.tree.store.m2[self.node_id].clone()

    @property
    def var(self):
        return self.tree.compute_var(self.meanSq, self.count)

    @property
    def std(self):
        return torch.sqrt(self.var)

    @property
    def parent(self):
        p = int(self.tree.store.parent[self.node_id].item())
        return CobwebNode(self.tree, p) if p >= 0 else None

    @property
    def children(self):
        s = self.tree.store
        n = int(s.child_cnt[self.node_id].item())
        if n == 0:
            return []
        off = int(s.child_off[self.node_id].item())
        return [CobwebNode(self.tree, c) for c in s.child_pool[off:off + n].cpu().tolist()]

    @property
    def sentence_id(self):
        return self.tree._sentence_list(self.node_id)

    @sentence_id.setter
    def sentence_id(self, value):
        self.tree._set_sentence_list(self.node_id, list(value or []))

    def log_prob(self, instance):
        """CobwebTorchNode.log_prob (CobwebTorchNode.py:100-104); accessor-level (torch ops on
        the device row), the batched kernels do not go through here."""
        x = self.tree._as_device_vec(instance)
        var = self.var
        return -(0.5 * torch.log(var) + 0.5 * math.log(2 * math.pi) + 0.5 * torch.square(x - self.mean) / var).sum()

    def depth(self):
        d, p = 0, self.parent
        while p is not None:
            d, p = d + 1, p.parent
        return d

    def num_concepts(self):
        return 1 + sum(c.num_concepts() for c in self.children)


class CobwebTorchTree:
    IFIT_CHUNK = 32768  # instances per kernel launch (bounds a launch to about a second)

    def __init__(self, shape, use_info=True, acuity_cutoff=False, use_kl=True, prior_var=None, alpha=1e-8,
                 device=None, greedy_mode=None):
        """greedy_mode (additive): None = the module switch constants.COBWEB_GREEDY_MODE at construction time (the
        reference reads src/utils/constants.py the same way)."""
        _lib.require_cuda()
        if isinstance(shape, torch.Size) or isinstance(shape, (tuple, list)):
            dims = tuple(int(v) for v in shape)
        else:
            dims = (int(shape),)
        if len(dims) != 1:
            raise ValueError(f"cobweb-b200 supports flat embeddings, got shape {dims}")
        dev = "cuda" if device in (None, "cuda") else device
        if not str(dev).startswith("cuda"):
            raise _lib.CobwebB200Error(f"device {device!r}: this engine runs on CUDA only (no CPU fallback)")
        self.device = dev
        self.shape = torch.Size(dims)
        self.use_info, self.acuity_cutoff, self.use_kl = bool(use_info), bool(acuity_cutoff), bool(use_kl)
        self.alpha = torch.tensor(alpha, dtype=torch.float32, device=self.device)
        self.pi_tensor = torch.tensor(math.pi, dtype=torch.float32, device=self.device)
        pv = default_prior_var() if prior_var is None else float(prior_var)
        self.prior_var = torch.tensor(pv, dtype=torch.float32, device=self.device)
        self.greedy_mode = bool(constants.COBWEB_GREEDY_MODE if greedy_mode is None else greedy_mode)
        flags = self._flags()
        self.store = NodeStore(dims[0], pv, flags, device=self.device)
        self._sent = {}     # node id -> SentenceList
        self._sent_stale = False   # a wrapper added sentences since _sent was built (CobwebWrapper._record)
        self._sent_loader = None
        self._frontier = None
        self.last_trace = None

    def _flags(self):
        return ((_lib.CW_USE_INFO if self.use_info else 0) | (_lib.CW_USE_KL if self.use_kl else 0) |
                (_lib.CW_ACUITY_CUTOFF if self.acuity_cutoff else 0) | (_lib.CW_GREEDY if self.greedy_mode else 0))

    # ------------------------------------------------------------------ basics
    @property
    def d(self):
        return self.shape[0]

    @property
    def root(self):
        return CobwebNode(self, self.store.root)

    def clear(self):
        self.store.clear()
        self._sent = {}
        self._sent_stale = False

    def __str__(self):
        return f"CobwebTorchTree(D={self.d}, nodes={self.num_nodes()})"

    def num_nodes(self):
        t = self.store.topology()
        return int(len(topology.bfs_order(t["root"], t["child_off"], t["child_cnt"], t["child_pool"])[0]))

    def _as_device_vec(self, x):
        return torch.as_tensor(np.asarray(x.detach().cpu() if torch.is_tensor(x) else x, dtype=np.float32),
                               device=self.device).reshape(-1)

    def _as_device_mat(self, X):
        if torch.is_tensor(X):
            X = X.detach().to(device=self.device, dtype=torch.float32)
        else:
            X = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32), device=self.device)
        X = X.reshape(-1, self.d) if X.dim() != 2 else X
        if X.shape[1] != self.d:
            raise ValueError(f"instance dim {X.shape[1]} != tree dim {self.d}")
        return X.contiguous()

    # ------------------------------------------------------------------ sentence bookkeeping
    def _sync_sentences(self):
        """Bring node.sentence_id lists up to date with the wrapper's sentence -> leaf map."""
        if self._sent_stale and self._sent_loader is not None:
            self._sent = self._sent_loader()
        self._sent_stale = False

    def _sentence_list(self, node_id):
        self._sync_sentences()
        lst = self._sent.get(node_id)
        if lst is None:
            lst = self._sent[node_id] = SentenceList(self, node_id)
        return lst

    def _set_sentence_list(self, node_id, items):
        self._sent[node_id] = SentenceList(self, node_id, items)
        self.store.n_sent[node_id] = len(items)

    def _note_sentence(self, node_id, sid):
        self.store.n_sent[node_id] += 1

    # ------------------------------------------------------------------ ifit
    def ifit_batch(self, X, tag_sentences=False, trace=False):
        """Insert the rows of X in order (CobwebTorchTree.ifit per row, CobwebTorchTree.py:123-233)
        on the device.  Returns the leaf node id per row (int32 tensor on the device); with
        trace=True also (ops int8 numpy, offsets int64 numpy) of the decisions taken."""
        X = self._as_device_mat(X)
        n = X.shape[0]
        leaves = torch.empty(n, dtype=torch.int32, device=self.device)
        L = _lib.load()
        tr_parts, off_parts = [], []
        pos = 0
        while pos < n:
            chunk = min(n - pos, self.IFIT_CHUNK)
            self.store.reserve(chunk)
            if trace:
                tcap = 64 * chunk + 1024
                tr = torch.zeros(tcap, dtype=torch.int8, device=self.device)
                toff = torch.zeros(chunk + 1, dtype=torch.int64, device=self.device)
                trp, toffp = tr.data_ptr(), toff.data_ptr()
            else:
                tcap, trp, toffp = 0, None, None
            _lib.check(L.cw_ifit(self.store.struct(), X[pos:].data_ptr(), chunk, leaves[pos:].data_ptr(), trp, toffp,
                                 tcap, int(bool(tag_sentences)), _lib.stream_ptr()), "cw_ifit")
            h = self.store.header()
            done, status = int(h[_lib.HDR_DONE]), int(h[_lib.HDR_STATUS])
            if trace:
                o = toff[: done + 1].cpu().numpy()
                if o[-1] > tcap:
                    raise _lib.CobwebB200Error("ifit trace buffer overflow")
                tr_parts.append(tr[: int(o[-1])].cpu().numpy())
                off_parts.append(o)
            pos += done
            if status == _lib.CW_E_CAPACITY:
                self.store.reserve(max(chunk, 4096) * 2, h=h)
                self.store.hdr[_lib.HDR_STATUS] = 0
            elif status != 0 or done < chunk:
                msg = (f"a node exceeded {_lib.MAX_CHILDREN} children (unsupported fan-out)" if status == _lib.CW_E_FANOUT
                       else f"cw_ifit kernel status {status}" if status else "cw_ifit stopped early without a status")
                err = _lib.CobwebB200Error(msg)
                err.completed, err.leaves = pos, leaves[:pos]   # rows [0, pos) are in the tree
                raise err
        if trace:
            ops = np.concatenate(tr_parts) if tr_parts else np.zeros(0, np.int8)
            offs, base = [np.zeros(1, np.int64)], 0
            for o in off_parts:
                offs.append(o[1:] + base)
                base += int(o[-1])
            self.last_trace = (ops, np.concatenate(offs))
            return leaves, ops, np.concatenate(offs)
        return leaves

    def ifit(self, instance):
        """Incrementally fit one instance; returns its concept (CobwebTorchTree.py:123)."""
        leaf = self.ifit_batch(self._as_device_vec(instance).reshape(1, -1))
        return CobwebNode(self, int(leaf.item()))

    # ------------------------------------------------------------------ categorize
    def categorize_batch(self, Q, retrieve_k=None, use_best=True, greedy=False, max_nodes=float("inf")):
        """Best-first search for a batch (CobwebTorchTree._cobweb_categorize, CobwebTorchTree.py:235-289).
        retrieve_k=None -> dict(best=[nq] node ids); else dict(leaves=[nq,k] (-1 padded), nfound=[nq]).
        Both carry lp_calls=[nq] (rows scored per query).  Tensors stay on the device."""
        Q = self._as_device_mat(Q)
        nq = Q.shape[0]
        k = 0 if retrieve_k is None else int(retrieve_k)
        L = _lib.load()
        h = self.store.header()
        n_live_bound = int(h[_lib.HDR_N_USED]) + 1
        n_ctas = max(1, min(L.cw_categorize_ctas(), nq, (1 << 31) // (n_live_bound * 4)))
        need = n_ctas * n_live_bound * 4
        if self._frontier is None or self._frontier.numel() < need:
            self._frontier = torch.empty(need, dtype=torch.int32, device=self.device)
        leaves = torch.full((nq, max(k, 1)), -1, dtype=torch.int32, device=self.device)
        nfound = torch.zeros(nq, dtype=torch.int32, device=self.device)
        best = torch.zeros(nq, dtype=torch.int32, device=self.device)
        calls = torch.zeros(nq, dtype=torch.int64, device=self.device)
        mn = (1 << 62) if (max_nodes is None or max_nodes == float("inf")) else int(max_nodes)
        _lib.check(L.cw_categorize(self.store.struct(), Q.data_ptr(), nq, k, mn, int(bool(greedy)), int(bool(use_best)),
                                   n_ctas, self._frontier.data_ptr(), n_live_bound, leaves.data_ptr(),
                                   nfound.data_ptr(), best.data_ptr(), calls.data_ptr(), _lib.stream_ptr()),
                   "cw_categorize")
        out = dict(lp_calls=calls)
        if k:
            out.update(leaves=leaves, nfound=nfound)
        else:
            out.update(best=best)
        return out

    def categorize(self, instance, use_best=True, greedy=False, max_nodes=float("inf"), retrieve_k=None):
        """CobwebTorchTree.categorize (CobwebTorchTree.py:291).  Like the reference, asking for
        more leaves than can be retrieved raises IndexError (:289)."""
        r = self.categorize_batch(self._as_device_vec(instance).reshape(1, -1), retrieve_k, use_best, greedy, max_nodes)
        if retrieve_k is None:
            return CobwebNode(self, int(r["best"].item()))
        nf = int(r["nfound"].item())
        if nf < retrieve_k:
            raise IndexError("list index out of range")
        return [CobwebNode(self, i) for i in r["leaves"][0, :retrieve_k].cpu().tolist()]

    # ------------------------------------------------------------------ formulas (accessor level)
    def compute_var(self, meanSq, count):
        """CobwebTorchTree.compute_var (CobwebTorchTree.py:336-342)."""
        if self.acuity_cutoff:
            return torch.clamp(meanSq / count, self.prior_var)
        return meanSq / count + self.prior_var

    def compute_score(self, mu1, var1, mu2, var2):
        """CobwebTorchTree.compute_score (CobwebTorchTree.py:344-364) on torch tensors; the ifit
        kernel evaluates the same expression internally."""
        if self.use_info:
            if self.use_kl:
                score = (torch.log(var2) - torch.log(var1)).sum()
                score += ((var1 + torch.pow(mu1 - mu2, 2)) / var2).sum()
                score -= mu1.numel()
                score /= 2
            else:
                score = 0.5 * (torch.log(var2) - torch.log(var1)).sum()
        else:
            score = -(1 / (2 * torch.sqrt(self.pi_tensor) * torch.sqrt(var1))).sum()
            score += (1 / (2 * torch.sqrt(self.pi_tensor) * torch.sqrt(var2))).sum()
        return score

    # ------------------------------------------------------------------ structure export
    def bfs(self):
        """dict(order, parent (BFS idx), depth, count, nchild, nsent) -- numpy, BFS order."""
        t = self.store.topology()
        order, parent_b, depth = topology.bfs_order(t["root"], t["child_off"], t["child_cnt"], t["child_pool"])
        return dict(order=order, parent=parent_b.astype(np.int32), depth=depth, count=t["count"][order],
                    nchild=t["child_cnt"][order], nsent=t["n_sent"][order])

    def analyze_structure(self):
        """CobwebTorchTree.analyze_structure (CobwebTorchTree.py:366-401)."""
        b = self.bfs()
        leaf_count = int((b["nchild"] == 0).sum())
        print(f"\nTotal number of leaf nodes: {leaf_count}\n")
        print("Number of nodes at each level:")
        for level, cnt in enumerate(np.bincount(b["depth"])):
            print(f"  Level {level}: {cnt} node(s)")
        print("\nParent nodes by number of children:")
        hist = np.bincount(b["nchild"])
        for nc in range(1, len(hist)):
            if hist[nc]:
                print(f" {hist[nc]} parent(s) with {nc} child(ren)")

    # ------------------------------------------------------------------ JSON
    def _sentence_ids_by_node(self):
        self._sync_sentences()
        return {nid: list(lst) for nid, lst in self._sent.items() if len(lst)}

    def dump_json(self):
        """CobwebTorchTree.dump_json (CobwebTorchTree.py:67-81): same nested document."""
        b = self.bfs()
        mean, m2 = self.store.rows(b["order"])
        sent = self._sentence_ids_by_node()
        sids = [sent.get(int(n), []) for n in b["order"]]
        return serialize.dump_tree_json(self._params(), b["parent"], b["count"], mean, m2, sids)

    def load_json(self, json_string):
        """CobwebTorchTree.load_json (CobwebTorchTree.py:94-121).  Child order is kept as
        written (the reference's loader reverses it)."""
        params, parent, count, mean, m2, sids = serialize.load_tree_json(json_string)
        self.use_info, self.acuity_cutoff, self.use_kl = params["use_info"], params["acuity_cutoff"], params["use_kl"]
        self.shape = torch.Size(params["shape"])
        self.alpha = torch.tensor(params["alpha"], dtype=torch.float32, device=self.device)
        self.prior_var = torch.tensor(params["prior_var"], dtype=torch.float32, device=self.device)
        self.store = NodeStore(self.shape[0], float(self.prior_var.item()), self._flags(), cap=len(parent) + 1024,
                               device=self.device)
        nsent = [len(s) if s else 0 for s in sids]
        self.store.load_arrays(parent, count, nsent, mean, m2)
        self._sent = {i: SentenceList(self, i, s) for i, s in enumerate(sids) if s}

    def _params(self):
        return dict(use_info=self.use_info, acuity_cutoff=self.acuity_cutoff, use_kl=self.use_kl, shape=list(self.shape),
                    alpha=self.alpha.item(), prior_var=self.prior_var.item())

    def save_snapshot(self, path, leaf_of_sentence=None, extra=None):
        """Additive: binary snapshot of the tree (serialize.write_snapshot) -- the same content as dump_json as raw arrays,
        streamed from the device in row chunks.  leaf_of_sentence: node ids (as returned by ifit) per sentence."""
        b = self.bfs()
        order = torch.as_tensor(b["order"].astype(np.int64), device=self.device)
        pos = np.full(int(b["order"].max()) + 1, -1, np.int64)
        pos[b["order"]] = np.arange(len(b["order"]))

        def rows(lo, hi):
            return self.store.mean[order[lo:hi]].cpu().numpy(), self.store.m2[order[lo:hi]].cpu().numpy()

        leaf = None if leaf_of_sentence is None else pos[np.asarray(leaf_of_sentence, np.int64)]
        serialize.write_snapshot(path, self._params(), b["parent"], b["count"], b["nsent"], rows, leaf, extra)

    def load_snapshot(self, path):
        """Replace this tree by a snapshot; returns (leaf_of_sentence as node ids of THIS store, extra)."""
        snap = serialize.read_snapshot(path)
        p = snap["params"]
        self.use_info, self.acuity_cutoff, self.use_kl = p["use_info"], p["acuity_cutoff"], p["use_kl"]
        self.shape = torch.Size(p["shape"])
        self.alpha = torch.tensor(p["alpha"], dtype=torch.float32, device=self.device)
        self.prior_var = torch.tensor(p["prior_var"], dtype=torch.float32, device=self.device)
        n = len(snap["parent"])
        self.store = NodeStore(self.shape[0], float(self.prior_var.item()), self._flags(), cap=n + 1024, device=self.device)
        self.store.load_arrays(snap["parent"], snap["count"], snap["n_sent"], snap["mean"], snap["m2"])
        self._sent, self._sent_stale = {}, False
        return snap["leaf_of_sentence"], snap["extra"]

    def load_arrays(self, parent, count, n_sent, mean, m2):
        """Replace the tree by flat arrays (topologically ordered; see NodeStore.load_arrays)."""
        self.store.load_arrays(parent, count, n_sent, mean, m2)
        self._sent = {}
