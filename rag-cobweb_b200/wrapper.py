"""CobwebWrapper with the reference's signatures (src/cobweb/CobwebWrapper.py), backed by the
device node store and kernels.

Reference surface kept: __init__(corpus, corpus_embeddings, encode_func) :13, add_sentences :52,
build_prediction_index :91, cobweb_predict_indexed :210, cobweb_rank_scores :267,
set_level_weights :335, set_weight_schedule :348, get_level_weights :410,
get_weight_schedule_info :414, force_rebuild_index :422, cobweb_predict_fast :428,
cobweb_predict :435, print_tree :463, dump_json :484, load_json :501, __len__ :557, attributes
sentences / sentence_to_node / tree / device / max_init_search.  visualize_subtrees (graphviz
rendering) is out of scope and raises NotImplementedError.
Additive: predict_fast_batch / predict_batch / rank_scores_batch for query batches.

Deliberate differences (DESIGN.md "Tie-breaking"): no 1e-6*randn noise and no in-leaf
random.shuffle -- exact ties are broken by ascending sentence id.
"""
import ctypes as C
import json
import os

import numpy as np
import torch

from . import _lib, topology
from .tree import CobwebNode, CobwebTorchTree


class DenseIndex:
    """Device-resident prediction index (build_prediction_index, CobwebWrapper.py:91-208)."""

    SCORE_BUDGET_BYTES = 16 << 30  # node-score scratch per query chunk
    # where the node scores are computed: "fp32" = FP32 pipe, (x*r + mb)^2 per triple (cw_dense.cu);
    # "tf32x3" = tcgen05 contraction with hi/lo-split TF32 operands (cw_tensor.cu, ~1e-6 relative to "fp32") as a
    # pre-filter for top-kc candidates, followed by the exact re-score (cw_rescore.cu): top-k ids and scores are
    # bit-identical to "fp32"; raw node / leaf score matrices in this mode are the approximate ones
    # "tf32x3f" = the same pre-filter with the path sums fused into the score kernel (cumulative ancestor sums +
    # leaf scores in the epilogue, candidates filtered against a per-query bound from a sample of the leaves): the
    # [nodes, queries] score matrix and the path kernel disappear for the leaves; same exact re-score, same results
    MODES = ("fp32", "tf32x3", "tf32x3f")
    FUSED_CAP = 1024  # candidate slots per query of the filtering epilogue
    FUSED_MIN_QUERIES = 256
    TENSOR_MIN_NODES = 16384  # smaller indexes are answered on the FP32 pipe (fewer launches, no read-back; same result)
    EPS_SCALE = 2.0 ** -18  # bound of |tf32x3 - fp32| leaf score relative to the operand magnitudes (cw_dense_rescore)

    def __init__(self, tree, leaf_of_sentence, level_weights=None, sentence_ids=None):
        """leaf_of_sentence[i] = leaf node id of sentence i.  sentence_ids (optional, sorted global
        ids) restricts the index to a shard of the sentences: only the nodes on their root->leaf
        paths are indexed and scored, returned ids stay global (store-sharded mode)."""
        L = _lib.load()
        self.tree = tree
        dev = tree.device
        t = tree.store.topology()
        order, parent_b, depth = topology.bfs_order(t["root"], t["child_off"], t["child_cnt"], t["child_pool"])
        leaf_of_sentence = np.asarray(leaf_of_sentence)
        self.sentence_ids = None
        if sentence_ids is not None:
            self.sentence_ids = np.asarray(sentence_ids, np.int64)
            leaf_of_sentence = leaf_of_sentence[self.sentence_ids]
            order, parent_b, depth = topology.restrict_to_paths(order, parent_b, depth, leaf_of_sentence, t["n_used"])
        self.order_host = order
        self._topo = (parent_b, depth, leaf_of_sentence, level_weights, t["n_used"])
        self.nn = len(order)
        self.max_depth = int(depth.max()) + 1
        d = tree.d
        self.n_ntiles = (self.nn + _lib.TILE_N - 1) // _lib.TILE_N
        self.n_ktiles = (d + _lib.TILE_K - 1) // _lib.TILE_K
        # rows of the node-major score matrix (covers the 128-row tiles of the FP32 kernel and the 256-row
        # tiles of the tensor-core kernel)
        self.ld = (self.nn + _lib.TC_TILE_N - 1) // _lib.TC_TILE_N * _lib.TC_TILE_N
        tile_elems = self.n_ntiles * self.n_ktiles * _lib.TILE_K * _lib.TILE_N
        self.R = torch.empty(tile_elems, dtype=torch.float32, device=dev)
        self.MB = torch.empty(tile_elems, dtype=torch.float32, device=dev)
        self.sumlog = torch.zeros(self.ld, dtype=torch.float32, device=dev)
        self.order = torch.as_tensor(order.astype(np.int32), device=dev)
        self.n_pos = len(leaf_of_sentence)
        ix = _lib.CwIndex()
        ix.D, ix.nn, ix.n_ntiles, ix.n_ktiles = d, self.nn, self.n_ntiles, self.n_ktiles
        ix.R, ix.MB, ix.sumlog = self.R.data_ptr(), self.MB.data_ptr(), self.sumlog.data_ptr()
        if self.n_pos:
            p = topology.sentence_paths(order, parent_b, depth, leaf_of_sentence, level_weights, n_slots=t["n_used"])
            self.max_len = p["max_len"]
            self.path_idx = torch.as_tensor(np.ascontiguousarray(p["path_idx"].T), device=dev)  # [n_pos, max_len]
            self.level_w = torch.as_tensor(p["level_w"], dtype=torch.float64, device=dev)
            if self.sentence_ids is not None:
                p["pos_rec"][:, 3] = self.sentence_ids[p["pos_rec"][:, 3]]  # local position ids -> global ids
            self.pos_rec = torch.as_tensor(p["pos_rec"], device=dev)  # [n_pos, 4]
            ix.n_pos, ix.max_len = self.n_pos, self.max_len
            ix.path_idx, ix.pos_rec, ix.level_w = self.path_idx.data_ptr(), self.pos_rec.data_ptr(), self.level_w.data_ptr()
        self.ix = ix
        _lib.check(L.cw_index_build(tree.store.struct(), self.order.data_ptr(), self.nn, C.byref(ix), _lib.stream_ptr()),
                   "cw_index_build")
        self._ws = None
        self.tx = None
        self.mode = "fp32"

    def set_mode(self, mode):
        """Select the scoring kernel for predict / predict_host / node_scores; builds the tensor-core
        operands (cw_tc_index_build) on first use."""
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {self.MODES}")
        if mode == "tf32x3f":
            self.set_mode("tf32x3")
            if not self.n_pos:
                return self
            if getattr(self, "fx", None) is None:
                self._build_fused()
            self.mode = mode
            return self
        if mode == "tf32x3" and self.tx is None:
            L, dev, d = _lib.load(), self.tree.device, self.tree.d
            tx = _lib.CwTcIndex()
            tx.D, tx.nn = d, self.nn
            tx.n_ntiles = (self.nn + _lib.TC_TILE_N - 1) // _lib.TC_TILE_N
            tx.n_slabs = (d + _lib.TC_SLAB_D - 1) // _lib.TC_SLAB_D
            self.tcB = torch.empty(L.cw_tc_b_bytes(self.nn, d), dtype=torch.uint8, device=dev)
            self.hconst = torch.empty(tx.n_ntiles * _lib.TC_TILE_N, dtype=torch.float32, device=dev)
            tx.B, tx.hconst = self.tcB.data_ptr(), self.hconst.data_ptr()
            _lib.check(L.cw_tc_index_build(self.tree.store.struct(), self.order.data_ptr(), self.nn, self.sumlog.data_ptr(),
                                           C.byref(tx), _lib.stream_ptr()), "cw_tc_index_build")
            # constants of the re-score margin: hmax = max_b sum_d mean^2/var (= -2 h - sumlog), lmax = max_b |sumlog|,
            # wfac = max over path lengths of sum_j |level_w[j]| / len
            sl = self.sumlog[: self.nn].double()
            tx.hmax = float((-2.0 * self.hconst[: self.nn].double() - sl).max().clamp_min(0.0)) * (1.0 + 1e-6) + 1e-6
            tx.lmax = float(sl.abs().max())
            tx.wfac, tx.eps_scale = 1.0, self.EPS_SCALE
            if self.n_pos:
                rec = self.pos_rec.cpu().numpy()
                cw = np.cumsum(np.abs(self.level_w.cpu().numpy()))
                lens = np.unique(rec[:, 0])
                tx.wfac = float(max(1.0, (cw[lens - 1] / lens).max()))
                pos_of_sid = np.full(int(rec[:, 3].max()) + 1, -1, np.int32)
                pos_of_sid[rec[:, 3]] = np.arange(len(rec), dtype=np.int32)
                self.pos_of_sid = torch.as_tensor(pos_of_sid, device=dev)
                tx.pos_of_sid = self.pos_of_sid.data_ptr()
            self.rows = torch.empty((self.nn, d, 2), dtype=torch.float32, device=dev)
            _lib.check(L.cw_rescore_rows_build(self.tree.store.struct(), self.order.data_ptr(), self.nn, self.rows.data_ptr(),
                                               _lib.stream_ptr()), "cw_rescore_rows_build")
            tx.rows = self.rows.data_ptr()
            self.tx = tx
        self.mode = mode
        return self

    def _build_fused(self):
        """Operands of the fused mode: separate score-kernel operand sets for the internal and the leaf rows
        (topology.fused_layout), per-leaf records, the flat index over the sampled leaves."""
        L, dev, d = _lib.load(), self.tree.device, self.tree.d
        parent_b, depth, leaf_of_sentence, level_weights, n_used = self._topo
        F = topology.fused_layout(self.order_host, parent_b, depth, leaf_of_sentence, level_weights, n_slots=n_used,
                                  sentence_ids=self.sentence_ids, tile=_lib.TC_TILE_N)
        order_np = np.asarray(self.order_host, np.int64)
        fx = {"F": F, "n_int": len(F["int_rows"]), "n_leaf": len(F["leaf_rows"]), "n_s": F["n_sample_tiles"]}

        def tc_index(rows):
            tx = _lib.CwTcIndex()
            tx.D, tx.nn = d, len(rows)
            tx.n_ntiles = (len(rows) + _lib.TC_TILE_N - 1) // _lib.TC_TILE_N
            tx.n_slabs = (d + _lib.TC_SLAB_D - 1) // _lib.TC_SLAB_D
            B = torch.empty(L.cw_tc_b_bytes(len(rows), d), dtype=torch.uint8, device=dev)
            h = torch.empty(tx.n_ntiles * _lib.TC_TILE_N, dtype=torch.float32, device=dev)
            od = torch.as_tensor(order_np[rows].astype(np.int32), device=dev)
            sl = self.sumlog[torch.as_tensor(rows, device=dev)].contiguous()
            tx.B, tx.hconst = B.data_ptr(), h.data_ptr()
            _lib.check(L.cw_tc_index_build(self.tree.store.struct(), od.data_ptr(), len(rows), sl.data_ptr(), C.byref(tx),
                                           _lib.stream_ptr()), "cw_tc_index_build")
            torch.cuda.current_stream().synchronize()  # od / sl may be freed after this
            return tx, B, h

        if fx["n_int"]:
            fx["tx_int"], fx["B_int"], fx["h_int"] = tc_index(F["int_rows"])
            fx["int_parent"] = torch.as_tensor(F["int_parent"], device=dev)
            fx["int_w"] = torch.as_tensor(F["int_w"], device=dev)
        fx["tx_leaf"], fx["B_leaf"], fx["h_leaf"] = tc_index(F["leaf_rows"])
        n_pad = fx["tx_leaf"].n_ntiles * _lib.TC_TILE_N
        rec = torch.zeros((n_pad, 4), dtype=torch.float32, device=dev)
        nl = fx["n_leaf"]
        rec[:nl, 0] = fx["h_leaf"][:nl]
        rec[:nl, 1] = torch.as_tensor(F["leaf_w"], device=dev)
        rec[:nl, 2] = torch.as_tensor(F["leaf_inv_len"], device=dev)
        par = torch.full((n_pad,), -1, dtype=torch.int32, device=dev)
        par[:nl] = torch.as_tensor(F["leaf_parent"], device=dev)
        rec[:, 3] = par.view(torch.float32)
        fx["leaf_rec"] = rec
        fx["sent_off"] = torch.as_tensor(F["sent_off"], device=dev)
        fx["sent_ids"] = torch.as_tensor(F["sent_ids"], device=dev)
        if fx["n_s"]:
            flat = _lib.CwIndex()
            flat.D, flat.nn, flat.n_pos, flat.max_len = d, fx["n_s"] * _lib.TC_TILE_N, len(F["flat_pos_rec"]), 1
            fx["flat_path"] = torch.as_tensor(F["flat_path"], device=dev)
            fx["flat_rec"] = torch.as_tensor(F["flat_pos_rec"], device=dev)
            fx["flat_w"] = torch.ones(1, dtype=torch.float64, device=dev)
            flat.path_idx, flat.pos_rec, flat.level_w = fx["flat_path"].data_ptr(), fx["flat_rec"].data_ptr(), fx["flat_w"].data_ptr()
            fx["flat"] = flat
        self.fx = fx

    def _fused_candidates(self, q, nq, kc, ws):
        """Top-kc candidates (sentence ids + approximate leaf scores, best first) of the fused mode into
        ws["cand_sid"] / ws["cand_val"]; returns the overflow flags [nq]."""
        L, fx, dev = _lib.load(), self.fx, self.tree.device
        T, ldq, st = _lib.TC_TILE_N, ws["ldq"], _lib.stream_ptr()
        n_int_pad = (fx["n_int"] + T - 1) // T * T
        n_s_rows = fx["n_s"] * T
        scores = ws["scores"]
        if scores.shape[0] < n_int_pad + n_s_rows:
            ws["fused_extra"] = ws.get("fused_extra") if ws.get("fused_extra") is not None and ws["fused_extra"].shape[0] >= n_s_rows \
                else torch.empty((max(n_s_rows, 1), ldq), dtype=torch.float32, device=dev)
            LS = ws["fused_extra"]
        else:
            LS = scores[n_int_pad:n_int_pad + max(n_s_rows, 1)]
        S = scores[:max(n_int_pad, 1)]
        cap = self.FUSED_CAP
        if ws.get("f_cap_q", 0) < nq:
            ws["f_val"] = torch.empty((ws["cap_q"], cap), dtype=torch.float32, device=dev)
            ws["f_row"] = torch.empty((ws["cap_q"], cap), dtype=torch.int32, device=dev)
            ws["f_cnt"] = torch.zeros(ws["cap_q"], dtype=torch.int32, device=dev)
            ws["f_ovf"] = torch.zeros(ws["cap_q"], dtype=torch.int32, device=dev)
            ws["f_samp_sid"] = torch.empty((ws["cap_q"], ws["cand_sid"].shape[1]), dtype=torch.int32, device=dev)
            ws["f_samp_val"] = torch.empty((ws["cap_q"], ws["cand_sid"].shape[1]), dtype=torch.float32, device=dev)
            ws["f_cap_q"] = ws["cap_q"]
        tx_leaf = fx["tx_leaf"]
        _lib.check(L.cw_tc_build_queries(C.byref(tx_leaf), q.data_ptr(), nq, ws["xt"].data_ptr(), st), "cw_tc_build_queries")
        if fx["n_int"]:
            tx_int = fx["tx_int"]
            _lib.check(L.cw_tc_score_tiles(C.byref(tx_int), ws["xt"].data_ptr(), nq, 0, 0, tx_int.n_ntiles, S.data_ptr(), ldq,
                                           None, None, 0, None, 0, None, None, None, st), "cw_tc_score_tiles (internal)")
            off = fx["F"]["level_off"]
            for lvl in range(len(off) - 1):
                _lib.check(L.cw_tc_cumsum_level(S.data_ptr(), ldq, int(off[lvl]), int(off[lvl + 1]), fx["int_parent"].data_ptr(),
                                                fx["int_w"].data_ptr(), st), "cw_tc_cumsum_level")
        tau = torch.full((nq,), float("-inf"), dtype=torch.float32, device=dev)
        samp_sid = samp_val = None
        if fx["n_s"]:
            _lib.check(L.cw_tc_score_tiles(C.byref(tx_leaf), ws["xt"].data_ptr(), nq, 1, 0, fx["n_s"], LS.data_ptr(), ldq,
                                           S.data_ptr(), fx["leaf_rec"].data_ptr(), fx["n_leaf"], None, 0, None, None, None, st),
                       "cw_tc_score_tiles (sample)")
            samp_sid, samp_val = ws["f_samp_sid"], ws["f_samp_val"]
            if kc <= 32:
                _lib.check(L.cw_dense_rows_topk(LS.data_ptr(), ldq, nq, n_s_rows, fx["sent_off"].data_ptr(),
                                                fx["sent_ids"].data_ptr(), kc, samp_sid.data_ptr(), samp_val.data_ptr(),
                                                ws["scratch"].data_ptr(), st), "cw_dense_rows_topk (sample)")
            else:
                _lib.check(L.cw_dense_paths_topk(C.byref(fx["flat"]), LS.data_ptr(), ldq, nq, kc, None, samp_sid.data_ptr(),
                                                 samp_val.data_ptr(), ws["scratch"].data_ptr(), st), "cw_dense_paths_topk (sample)")
            sv = samp_val.view(-1)[: nq * kc].view(nq, kc)
            ss = samp_sid.view(-1)[: nq * kc].view(nq, kc)
            tau = torch.where(ss[:, kc - 1] >= 0, sv[:, kc - 1], tau)
        ws["f_cnt"][:nq].zero_()
        _lib.check(L.cw_tc_score_tiles(C.byref(tx_leaf), ws["xt"].data_ptr(), nq, 2, fx["n_s"], tx_leaf.n_ntiles - fx["n_s"], None,
                                       ldq, S.data_ptr(), fx["leaf_rec"].data_ptr(), fx["n_leaf"], tau.data_ptr(), cap,
                                       ws["f_cnt"].data_ptr(), ws["f_val"].data_ptr(), ws["f_row"].data_ptr(), st),
                   "cw_tc_score_tiles (filter)")
        _lib.check(L.cw_tc_select(nq, kc, samp_sid.data_ptr() if samp_sid is not None else None,
                                  samp_val.data_ptr() if samp_val is not None else None, cap, ws["f_cnt"].data_ptr(),
                                  ws["f_val"].data_ptr(), ws["f_row"].data_ptr(), fx["sent_off"].data_ptr(),
                                  fx["sent_ids"].data_ptr(), ws["cand_sid"].data_ptr(), ws["cand_val"].data_ptr(),
                                  ws["f_ovf"].data_ptr(), st), "cw_tc_select")
        return ws["f_ovf"][:nq]

    def _node_scores_call(self, q, nq, ws):
        L = _lib.load()
        if self.mode in ("tf32x3", "tf32x3f"):
            _lib.check(L.cw_dense_node_scores_tc(C.byref(self.tx), q.data_ptr(), nq, ws["xt"].data_ptr(),
                                                 ws["scores"].data_ptr(), ws["ldq"], _lib.stream_ptr()), "cw_dense_node_scores_tc")
        else:
            _lib.check(L.cw_dense_node_scores(C.byref(self.ix), q.data_ptr(), nq, ws["xt"].data_ptr(),
                                              ws["scores"].data_ptr(), ws["ldq"], _lib.stream_ptr()), "cw_dense_node_scores")

    def bytes(self):
        b = (self.R.numel() + self.MB.numel() + self.sumlog.numel()) * 4
        if self.tx is not None:
            b += self.tcB.numel() + self.hconst.numel() * 4 + self.rows.numel() * 4
        return b

    def chunk_queries(self):
        return int(max(256, min(65535, self.SCORE_BUDGET_BYTES // (self.ld * 4)) // 256 * 256))  # whole 256-query tiles

    def workspace(self, nq, k):
        """Device work buffers for chunks of up to nq queries and top-k up to k (grown on demand)."""
        ws = self._ws
        if ws and ws["cap_q"] >= nq and ws["cap_k"] >= k:
            return ws
        L, dev = _lib.load(), self.tree.device
        nq = max(nq, ws["cap_q"]) if ws else nq
        k = max(k, ws["cap_k"], 1) if ws else max(k, 1)
        self._ws = ws = None
        ws = dict(
            cap_q=nq, cap_k=k,
            q=torch.empty((nq, self.tree.d), dtype=torch.float32, device=dev),
            # query operands: k-major tiles (FP32 kernel) or swizzled hi/lo images (tensor-core kernel)
            xt=torch.empty(max(L.cw_xt_floats(nq, self.tree.d) * 4, L.cw_tc_a_bytes(nq, self.tree.d)), dtype=torch.uint8,
                           device=dev),
            ldq=int(L.cw_score_ldq(nq)),
            scores=torch.empty((self.ld, int(L.cw_score_ldq(nq))), dtype=torch.float32, device=dev),  # node-major
            sid=torch.empty((nq, k), dtype=torch.int32, device=dev),
            val=torch.empty((nq, k), dtype=torch.float32, device=dev),
        )
        c0 = self.candidates(k)
        kc = max(k, c0, 32 if c0 else 0, self.candidates(k, 1))  # 32: room for the adaptive first level
        ws["scratch"] = torch.empty(max(1, nq * L.cw_topk_chunks(max(self.n_pos, 1)) * kc * 2), dtype=torch.int32, device=dev)
        ws["cand_sid"] = torch.empty((nq, kc), dtype=torch.int32, device=dev)
        ws["cand_val"] = torch.empty((nq, kc), dtype=torch.float32, device=dev)
        ws["fail"] = torch.zeros(1 + nq, dtype=torch.int32, device=dev)
        self._ws = ws
        return ws

    def candidates(self, k, level=0):
        """Candidates per query the tensor-core pre-filter hands to the exact re-score: level 0 = first attempt,
        level 1 = second attempt for the queries the first one flagged (0 = no such level: this k is served by the
        FP32 path -- k too large, or paths too long for the re-score kernel's shared memory)."""
        if k < 1 or k > 32 or not self.n_pos or level > 1:
            return 0
        # enough that the weakest candidate sits below (k-th best) - 2 eps for (nearly) every query: the first level
        # starts at 24 (k <= 10: at cfg3 and cfg4 the gap between ranks 10 and 24 exceeds the margin for all but
        # ~1 query in 10^4, tools/tc_gap_probe.py; 16 would save 0.9 ms of path kernel per 10k queries but the
        # second pass for its 0.1 % flagged queries costs more) and is raised by _adapt() when more than 0.5 % of
        # a batch had to be escalated; lists up to 32 use the fast insertion path
        kc = min(max(self._kc0, k + 6), 32) if k <= 16 else 2 * k
        if level == 1:
            if kc >= _lib.RESCORE_MAX_KC:
                return 0
            kc = _lib.RESCORE_MAX_KC
        if kc * self.max_len > 65535 or _lib.load().cw_rescore_smem_bytes(self.tree.d, self.max_len, kc) > 200 * 1024:
            return 0
        return kc

    def _work_struct(self, ws, k):
        w = _lib.CwDenseWork()
        w.Q_dev, w.xt_scratch, w.node_scores, w.ldq = ws["q"].data_ptr(), ws["xt"].data_ptr(), ws["scores"].data_ptr(), ws["ldq"]
        w.out_sid_dev, w.out_score_dev, w.scratch = ws["sid"].data_ptr(), ws["val"].data_ptr(), ws["scratch"].data_ptr()
        w.cand_sid, w.cand_score, w.fail = ws["cand_sid"].data_ptr(), ws["cand_val"].data_ptr(), ws["fail"].data_ptr()
        w.kc, w.kc2 = self.candidates(k), self.candidates(k, 1)
        return w

    def node_scores(self, Q):
        """[nq, nn] node log-likelihood scores in index (BFS) order (CobwebWrapper.py:283-287)."""
        nq = Q.shape[0]
        ws = self.workspace(nq, 0)
        self._node_scores_call(Q, nq, ws)
        return ws["scores"][: self.nn, :nq].T

    def predict(self, Q, k, want_leaf_scores=False, mode=None, _level=0):
        """Device batch -> (sids [nq,k] int32, scores [nq,k], leaf_scores [nq,L] or None).  mode overrides
        self.mode for this call.  In "tf32x3" mode top-k goes pre-filter -> exact re-score; flagged queries are
        answered again with more candidates (self.n_escalated) and what is still flagged on the FP32 pipe
        (self.n_fallback); leaf_scores, if requested, come from the same node scores as the top-k of that mode
        (approximate for "tf32x3")."""
        L = _lib.load()
        if want_leaf_scores and self.sentence_ids is not None:
            raise ValueError("leaf scores are indexed by global sentence id; not available on a sentence shard")
        mode = mode or self.mode
        if mode != self.mode:
            prev = self.mode
            self.set_mode(mode)
            try:
                return self.predict(Q, k, want_leaf_scores)
            finally:
                self.mode = prev
        tensor = mode in ("tf32x3", "tf32x3f")
        if tensor and k > 0 and not want_leaf_scores and self.nn < self.TENSOR_MIN_NODES:
            return self.predict(Q, k, mode="fp32")  # launch-bound at this size: the FP32 path is the shorter one
        kc = self.candidates(k, _level) if (tensor and not want_leaf_scores) else 0
        if tensor and kc == 0 and k > 0 and not want_leaf_scores:
            return self.predict(Q, k, mode="fp32")  # this k / depth is not served by the re-score kernel
        # small batches: the fused pipeline's extra launches (one per tree level) cost more than the path kernel saves
        fused = (mode == "tf32x3f" and kc > 0 and _level == 0 and getattr(self, "fx", None) is not None and
                 Q.shape[0] >= self.FUSED_MIN_QUERIES)
        if fused and self.fx["n_s"] == 0 and self.fx["n_leaf"] > self.FUSED_CAP:
            fused = False  # too few leaf tiles to sample a threshold from, too many leaves for the candidate buffer
        nq_total = Q.shape[0]
        step = self.chunk_queries()
        sids = torch.empty((nq_total, max(k, 1)), dtype=torch.int32, device=Q.device)
        vals = torch.empty((nq_total, max(k, 1)), dtype=torch.float32, device=Q.device)
        leaf = torch.empty((nq_total, self.n_pos), dtype=torch.float32, device=Q.device) if want_leaf_scores else None
        redo = []
        for lo in range(0, nq_total, step):
            nq = min(step, nq_total - lo)
            ws = self.workspace(min(step, nq_total), k)
            q = Q[lo:lo + nq]
            ovf = None
            if fused:
                ovf = self._fused_candidates(q, nq, kc, ws)
            else:
                self._node_scores_call(q, nq, ws)
            if kc:
                if not fused:
                    _lib.check(L.cw_dense_paths_topk(C.byref(self.ix), ws["scores"].data_ptr(), ws["ldq"], nq, kc, None,
                                                     ws["cand_sid"].data_ptr(), ws["cand_val"].data_ptr(),
                                                     ws["scratch"].data_ptr(), _lib.stream_ptr()), "cw_dense_paths_topk")
                tx = self.tx
                _lib.check(L.cw_dense_rescore(self.tree.store.struct(), C.byref(self.ix), tx.rows, tx.pos_of_sid, q.data_ptr(),
                                              nq, kc, ws["cand_sid"].data_ptr(), ws["cand_val"].data_ptr(), k, tx.hmax, tx.lmax,
                                              tx.wfac, tx.eps_scale, sids[lo:lo + nq].data_ptr(), vals[lo:lo + nq].data_ptr(),
                                              ws["fail"].data_ptr(), _lib.stream_ptr()), "cw_dense_rescore")
                nf = int(ws["fail"][0])  # one 4-byte read-back per chunk
                flagged = ws["fail"][1:1 + nf].long()
                if ovf is not None and bool(ovf.any()):  # candidate buffer overflow: treat like a flagged query
                    flagged = torch.unique(torch.cat([flagged, torch.nonzero(ovf).view(-1)]))
                if flagged.numel():
                    redo.append(flagged + lo)
            else:
                _lib.check(L.cw_dense_paths_topk(C.byref(self.ix), ws["scores"].data_ptr(), ws["ldq"], nq, k,
                                                 leaf[lo:lo + nq].data_ptr() if leaf is not None else None,
                                                 sids[lo:lo + nq].data_ptr(), vals[lo:lo + nq].data_ptr(),
                                                 ws["scratch"].data_ptr(), _lib.stream_ptr()), "cw_dense_paths_topk")
        if kc and _level == 0:
            self._adapt(nq_total, sum(int(r.numel()) for r in redo))
        if redo:
            idx = torch.cat(redo)
            if self.candidates(k, _level + 1):
                self.n_escalated += int(idx.numel())
                s2, v2, _ = self.predict(Q[idx].contiguous(), k, mode="tf32x3" if fused else None, _level=_level + 1)
            else:
                self.n_fallback += int(idx.numel())
                s2, v2, _ = self.predict(Q[idx].contiguous(), k, mode="fp32")
            sids[idx], vals[idx] = s2, v2
        return sids, vals, leaf

    _kc0 = 24        # first-level candidates per query (k <= 16), adapted to the escalation rate

    def _adapt(self, n_queries, n_flagged):
        if n_queries >= 64 and n_flagged > 0.005 * n_queries and self._kc0 < 32:
            self._kc0 += 8

    n_escalated = 0  # queries whose first candidate list could not be decided and were re-run with more candidates
    n_fallback = 0   # queries answered by the FP32 path because the re-score margin did not hold at any level

    def predict_host(self, Q_host, k, out_sid=None, out_val=None):
        """Host batch (numpy / pinned tensor) -> host ids/scores through the single C-ABI call
        cw_predict_dense_host (H2D + kernels + D2H + sync inside)."""
        L = _lib.load()
        Qh = Q_host if torch.is_tensor(Q_host) else torch.from_numpy(np.ascontiguousarray(Q_host, np.float32))
        nq_total = Qh.shape[0]
        step = self.chunk_queries()
        if out_sid is None:
            out_sid = torch.empty((nq_total, k), dtype=torch.int32)
            out_val = torch.empty((nq_total, k), dtype=torch.float32)
        if self.mode == "tf32x3f" and self.candidates(k) > 0 and self.nn >= self.TENSOR_MIN_NODES:
            # fused mode: pinned host batch -> device, the device pipeline, ids/scores back (no single C call yet)
            for lo in range(0, nq_total, step):
                nq = min(step, nq_total - lo)
                ws = self.workspace(min(step, nq_total), k)
                qd = ws["q"][:nq]
                qd.copy_(Qh[lo:lo + nq], non_blocking=True)
                sd, vd, _ = self.predict(qd, k)
                out_sid[lo:lo + nq].copy_(sd, non_blocking=True)
                out_val[lo:lo + nq].copy_(vd, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_sid, out_val
        tensor = self.mode in ("tf32x3", "tf32x3f") and self.candidates(k) > 0 and self.nn >= self.TENSOR_MIN_NODES
        nfb = (C.c_int32 * 2)(0, 0)
        for lo in range(0, nq_total, step):
            nq = min(step, nq_total - lo)
            ws = self.workspace(min(step, nq_total), k)
            w = self._work_struct(ws, k)
            _lib.check(L.cw_predict_dense_host(C.byref(self.ix), C.byref(self.tx) if tensor else None,
                                               self.tree.store.struct(), Qh[lo:lo + nq].data_ptr(), nq, k, C.byref(w),
                                               out_sid[lo:lo + nq].data_ptr(), out_val[lo:lo + nq].data_ptr(), nfb,
                                               _lib.stream_ptr()), "cw_predict_dense_host")
            self.n_escalated += nfb[0]
            self.n_fallback += nfb[1]
            if tensor:
                self._adapt(nq, nfb[0])
        return out_sid, out_val


class _RankScores(torch.autograd.Function):
    """cobweb_rank_scores with a gradient w.r.t. the queries (cw_rank_scores_bwd)."""

    @staticmethod
    def forward(ctx, Q, index):
        _, _, leaf = index.predict(Q.detach(), 0, want_leaf_scores=True, mode="fp32")
        ctx.index = index
        ctx.save_for_backward(Q.detach())
        return leaf

    @staticmethod
    def backward(ctx, grad_leaf):
        (Q,) = ctx.saved_tensors
        ix = ctx.index
        nq = Q.shape[0]
        out = torch.empty_like(Q)
        step = max(32, min(nq, ix.chunk_queries()))
        for lo in range(0, nq, step):
            n = min(step, nq - lo)
            gs = torch.empty((ix.nn, n), dtype=torch.float32, device=Q.device)
            g = grad_leaf[lo:lo + n].contiguous().to(torch.float32)
            _lib.check(_lib.load().cw_rank_scores_bwd(C.byref(ix.ix), Q[lo:lo + n].data_ptr(), n, g.data_ptr(), gs.data_ptr(),
                                                      n, out[lo:lo + n].data_ptr(), _lib.stream_ptr()), "cw_rank_scores_bwd")
        return out, None


class CobwebWrapper:
    def __init__(self, corpus=None, corpus_embeddings=None, encode_func=lambda x: x):
        _lib.require_cuda()
        self.encode_func = encode_func
        self.sentences = []
        self.device = "cuda"
        self.max_init_search = 100000
        self._index = None
        self._level_weights = None
        self._weight_schedule = None
        self._schedule_params = {}
        self.max_depth = 0
        self._leaf_of_sentence = np.zeros(0, np.int32)

        if corpus_embeddings is not None:
            if isinstance(corpus_embeddings, list):
                corpus_embeddings = torch.tensor(corpus_embeddings)
            embedding_shape = corpus_embeddings.shape[1:]
        elif corpus and len(corpus) > 0:
            sample_emb = self.encode_func([corpus[0]])
            embedding_shape = sample_emb.shape[1:]
        else:
            raise ValueError("CobwebWrapper needs a corpus or corpus_embeddings to size the tree")
        self.tree = CobwebTorchTree(shape=embedding_shape, device=self.device)
        if corpus_embeddings is not None:
            if corpus is None:
                corpus = [None] * len(corpus_embeddings)
            self.add_sentences(corpus, corpus_embeddings)
        elif corpus is not None and len(corpus) > 0:
            self.add_sentences(corpus)

    # ------------------------------------------------------------------ build
    def add_sentences(self, new_sentences, new_vectors=None):
        """CobwebWrapper.add_sentences (CobwebWrapper.py:52-80): one ifit per row, in order; the
        leaf -> sentence-id bookkeeping is done by the kernel (n_sent) and the id map below."""
        if new_vectors is None:
            new_embeddings = self.encode_func(new_sentences)
        else:
            new_embeddings = new_vectors
            if isinstance(new_embeddings, list):
                new_embeddings = torch.tensor(new_embeddings)
            if new_embeddings.shape[1] != self.tree.shape[0]:
                print(f"[Warning] Provided vector dim {new_embeddings.shape[1]} != tree dim {self.tree.shape[0]}, re-encoding...")
                new_embeddings = self.encode_func(new_sentences)
        n = len(new_sentences)
        X = self.tree._as_device_mat(new_embeddings)[:n]
        leaves = self.tree.ifit_batch(X, tag_sentences=True)
        self.sentences.extend(new_sentences)
        self._leaf_of_sentence = np.concatenate([self._leaf_of_sentence, leaves.cpu().numpy()])
        self.tree._sent = {}
        self._invalidate_prediction_index()

    @property
    def sentence_to_node(self):
        """sentence id -> concept handle (CobwebWrapper.py:77)."""
        return {i: CobwebNode(self.tree, int(n)) for i, n in enumerate(self._leaf_of_sentence)}

    def _sync_sentence_lists(self):
        if self.tree._sent:
            return
        order = np.argsort(self._leaf_of_sentence, kind="stable")
        for sid in order:
            self.tree._sent.setdefault(int(self._leaf_of_sentence[sid]), []).append(int(sid))

    def _invalidate_prediction_index(self):
        self._index = None
        self._shard_key = None
        self._shard_index = None

    @property
    def _prediction_index_valid(self):
        return self._index is not None

    def build_prediction_index(self):
        """CobwebWrapper.build_prediction_index (CobwebWrapper.py:91-208)."""
        if self._index is not None:
            return
        ix = DenseIndex(self.tree, self._leaf_of_sentence, self._level_weights)
        self._index = ix.set_mode(self._resolve_mode(ix))
        self.max_depth = max(self.max_depth, self._index.max_depth)

    # Scoring mode of the dense index (DenseIndex.MODES; every mode returns the same ids and scores).  "auto": the
    # fused tensor-core mode for indexes of AUTO_TENSOR_NODES nodes and more -- it builds two more operand copies, which
    # 180 GB of HBM is there for -- and the FP32 pipe below (launch-bound at that size).
    dense_mode = os.environ.get("COBWEB_B200_DENSE_MODE", "auto")
    AUTO_TENSOR_NODES = 16384

    def _resolve_mode(self, index):
        if self.dense_mode != "auto":
            return self.dense_mode
        return "tf32x3f" if index.nn >= self.AUTO_TENSOR_NODES else "fp32"

    def set_dense_mode(self, mode):
        """Additive: where cobweb_predict_fast / predict_fast_batch compute node scores -- "auto" (default), "fp32"
        (FP32 pipe), "tf32x3" (tcgen05 tensor cores, split-TF32 operands), "tf32x3f" (the same, fused).  See
        DenseIndex.MODES."""
        if mode != "auto" and mode not in DenseIndex.MODES:
            raise ValueError(f"mode must be 'auto' or one of {DenseIndex.MODES}")
        self.dense_mode = mode
        if self._index is not None:
            self._index.set_mode(self._resolve_mode(self._index))
        if getattr(self, "_shard_index", None) is not None:
            self._shard_index.set_mode(self._resolve_mode(self._shard_index))

    def force_rebuild_index(self):
        self._invalidate_prediction_index()
        self.build_prediction_index()

    # ------------------------------------------------------------------ level weights
    def set_level_weights(self, weights):
        self._level_weights = weights
        self._weight_schedule = None
        self._invalidate_prediction_index()

    def set_weight_schedule(self, schedule_type, max_depth=10, **kwargs):
        if self._prediction_index_valid:
            max_depth = self.max_depth
        self._weight_schedule = schedule_type
        self._schedule_params = kwargs
        self._level_weights = topology.generate_weight_schedule(schedule_type, max_depth, **kwargs)
        self._invalidate_prediction_index()

    def get_level_weights(self):
        return self._level_weights if self._level_weights is not None else [1.0, 1.0, 1.0, 1.0]

    def get_weight_schedule_info(self):
        return {"schedule_type": self._weight_schedule, "schedule_params": self._schedule_params,
                "current_weights": self.get_level_weights()}

    def get_prediction_index_info(self):
        ix = self._index
        return {"index_valid": ix is not None, "total_nodes": ix.nn if ix else 0,
                "leaf_paths_cached": ix.n_pos if ix else 0, "means_cached": ix is not None, "vars_cached": ix is not None}

    # ------------------------------------------------------------------ dense predict
    def _embed(self, input, is_embedding):
        emb = input if is_embedding else self.encode_func([input])[0]
        return self.tree._as_device_vec(emb).reshape(1, -1)

    def predict_fast_batch(self, Q, k=5):
        """Batched cobweb_predict_fast(return_ids=True): device tensors (ids [nq,k], scores [nq,k])."""
        self.build_prediction_index()
        Q = self.tree._as_device_mat(Q)
        k = min(int(k), self._index.n_pos)
        if k > _lib.MAX_K:
            _, _, leaf = self._index.predict(Q, 0, want_leaf_scores=True, mode="fp32")
            vals, ids = torch.sort(leaf, dim=1, descending=True, stable=True)
            return ids[:, :k].to(torch.int32), vals[:, :k]
        sids, vals, _ = self._index.predict(Q, k)
        return sids, vals

    def predict_fast_sharded(self, Q, k=5, world=None, rank=None):
        """Store-sharded dense predict (trees beyond one HBM): this rank indexes and
        scores only its contiguous share of the sentences (in tree order) plus the nodes on their
        paths, then per-rank top-k lists are all-gathered and merged.  Same ids and scores as
        predict_fast_batch.  With world/rank given explicitly and no process group, returns this
        shard's candidates (used by the single-GPU test that emulates the ranks one after another)."""
        import torch.distributed as dist
        from . import parallel
        live = dist.is_available() and dist.is_initialized()
        if world is None:
            world, rank = (dist.get_world_size(), dist.get_rank()) if live else (1, 0)
        key = (world, rank, len(self.sentences))
        if getattr(self, "_shard_key", None) != key:
            t = self.tree.store.topology()
            order, parent_b, depth = topology.bfs_order(t["root"], t["child_off"], t["child_cnt"], t["child_pool"])
            row_of = np.full(t["n_used"], -1, np.int64)
            row_of[order] = np.arange(len(order))
            tree_order = np.lexsort((np.arange(len(self._leaf_of_sentence)), row_of[self._leaf_of_sentence]))
            lo, hi = parallel.shard_bounds(len(tree_order), world, rank)
            self._shard_index = DenseIndex(self.tree, self._leaf_of_sentence, self._level_weights,
                                           sentence_ids=np.sort(tree_order[lo:hi]))
            self._shard_index.set_mode(self._resolve_mode(self._shard_index))
            self._shard_key = key
        Q = self.tree._as_device_mat(Q)
        kk = min(int(k), self._shard_index.n_pos, _lib.MAX_K)
        ids, vals, _ = self._shard_index.predict(Q, kk)
        if kk < k:  # pad so every rank contributes the same number of candidates
            pad = k - kk
            ids = torch.cat([ids, torch.full((ids.shape[0], pad), -1, dtype=ids.dtype, device=ids.device)], 1)
            vals = torch.cat([vals, torch.full((vals.shape[0], pad), float("-inf"), device=vals.device)], 1)
        if live and world > 1:
            ci, cv = parallel.gather_candidates(ids, vals)
            return parallel.merge_topk(ci, cv, k)
        return ids, vals

    def rank_scores_batch(self, Q):
        """Batched cobweb_rank_scores: [nq, L] leaf scores indexed by sentence id.  Differentiable
        w.r.t. Q when Q is a CUDA tensor that requires grad (the training use of the reference,
        src/training/cobweb_query_train.py:104-126)."""
        self.build_prediction_index()
        if torch.is_tensor(Q) and Q.requires_grad:
            Qd = Q.to(device=self.device, dtype=torch.float32)
            Qd = Qd.reshape(1, -1) if Qd.dim() == 1 else Qd
            return _RankScores.apply(Qd.contiguous(), self._index)
        _, _, leaf = self._index.predict(self.tree._as_device_mat(Q), 0, want_leaf_scores=True, mode="fp32")
        return leaf

    def cobweb_predict_indexed(self, input, k=5, return_ids=False, is_embedding=False):
        """CobwebWrapper.cobweb_predict_indexed (CobwebWrapper.py:210-265), noise-free."""
        self.build_prediction_index()
        if len(self.sentences) == 0:
            return []
        ids, _ = self.predict_fast_batch(self._embed(input, is_embedding), k)
        out = []
        for sid in ids[0].cpu().tolist():
            if 0 <= sid < len(self.sentences):
                out.append(sid if return_ids else self.sentences[sid])
        return out

    def cobweb_predict_fast(self, input, k=5, return_ids=False, is_embedding=False):
        return self.cobweb_predict_indexed(input, k, return_ids, is_embedding)

    def cobweb_rank_scores(self, input, is_embedding=False):
        """CobwebWrapper.cobweb_rank_scores (CobwebWrapper.py:267-294); differentiable w.r.t. a tensor
        input that requires grad, like the reference's torch expression."""
        self.build_prediction_index()
        if len(self.sentences) == 0:
            return torch.empty(0, device=self.device)
        x = input if is_embedding else self.encode_func([input])[0]
        if torch.is_tensor(x) and x.requires_grad:
            return self.rank_scores_batch(x.reshape(1, -1))[0]
        return self.rank_scores_batch(self.tree._as_device_vec(x).reshape(1, -1))[0]

    # ------------------------------------------------------------------ best-first predict
    def predict_batch(self, Q, k=5):
        """Batched cobweb_predict: returns (leaves [nq,k] node ids on the host, nfound [nq],
        lp_calls [nq])."""
        r = self.tree.categorize_batch(Q, retrieve_k=k, use_best=True, max_nodes=self.max_init_search)
        return r["leaves"].cpu().numpy(), r["nfound"].cpu().numpy(), r["lp_calls"].cpu().numpy()

    def cobweb_predict(self, input, k=5, return_ids=False, is_embedding=False):
        """CobwebWrapper.cobweb_predict (CobwebWrapper.py:435-461)."""
        emb = input if is_embedding else self.encode_func([input])[0]
        leaves = self.tree.categorize(emb, use_best=True, max_nodes=self.max_init_search, retrieve_k=k)
        self._sync_sentence_lists()
        results = []
        for leaf in leaves:
            for sid in sorted(self.tree._sent.get(leaf.node_id, [])):
                if sid is None or sid >= len(self.sentences):
                    continue
                results.append(sid if return_ids else self.sentences[sid])
        return results

    # ------------------------------------------------------------------ misc
    def print_tree(self):
        self._sync_sentence_lists()
        b = self.tree.bfs()
        kids = [[] for _ in b["order"]]
        for i in range(1, len(b["order"])):
            kids[b["parent"][i]].append(i)

        def rec(i, depth):
            nid = int(b["order"][i])
            print(f"{'  ' * depth}- Node ID {nid} Sentence ID: {list(self.tree._sent.get(nid, []))}")
            for c in kids[i]:
                rec(c, depth + 1)

        print("\nCobweb Sentence Clustering Tree:")
        rec(0, 0)

    def dump_json(self, save_path=None):
        """CobwebWrapper.dump_json (CobwebWrapper.py:484-497)."""
        self._sync_sentence_lists()
        state = {"tree": json.loads(self.tree.dump_json()), "sentences": self.sentences,
                 "embedding_dim": self.tree.shape[0]}
        if save_path:
            with open(save_path, "w") as f:
                json.dump(state, f, indent=2)
        return json.dumps(state, indent=2)

    @staticmethod
    def load_json(json_data, encode_func=lambda x: x):
        """CobwebWrapper.load_json (CobwebWrapper.py:500-555).  The reference's version crashes on
        list-valued sentence ids; this one restores the id -> leaf map."""
        data = json.loads(json_data) if isinstance(json_data, str) else json_data
        w = CobwebWrapper.__new__(CobwebWrapper)
        w.encode_func = encode_func
        w.device = "cuda"
        w.sentences = data.get("sentences", [])
        w.max_init_search = data.get("max_init_search", 100000)
        w._index, w._level_weights, w._weight_schedule, w._schedule_params, w.max_depth = None, None, None, {}, 0
        w.tree = CobwebTorchTree(shape=(int(data["embedding_dim"]),), device=w.device)
        w.tree.load_json(json.dumps(data["tree"]))
        leaf = np.full(len(w.sentences), -1, np.int32)
        for nid, lst in w.tree._sent.items():
            for sid in lst:
                if 0 <= sid < len(leaf):
                    leaf[sid] = nid
        w._leaf_of_sentence = leaf
        return w

    def visualize_subtrees(self, directory, num_leaves=6):
        raise NotImplementedError("graphviz rendering is outside the hot-path scope (SURVEY.md section 2, row 3)")

    def __len__(self):
        return len(self.sentences)
