"""Batched retrieval evaluation with the reference's metric definitions
(src/utils/benchmark_utils.py: get_eval_ks :620, evaluate_retrieval :710-833, retrieve_torch_dot :602-614).

The reference evaluates one query at a time (retrieve_fn(query, top_k) -> list of documents) and
accumulates recall@k / mrr@k / ndcg@k over get_eval_ks(top_k); here the whole query set goes through
the batched predict entry points and the same numbers come out of one vectorised pass.

ndcg@k follows the reference's call exactly: sklearn ndcg_score([sorted(relevance, reverse=True)], [relevance])
with relevance the 0/1 hit list.  With a single relevant document that is 1.0 when the target is ranked
first and otherwise the tie-averaged gain of the k-1 zero-scored positions, (1/(k-1)) * sum_{i=2..k} 1/log2(i+1)
-- independent of the actual rank (a quirk of passing the sorted list as y_true).  The conventional
1/log2(rank+1) is reported additionally as ndcg_std@k.
"""
import time

import numpy as np
import torch


def get_eval_ks(top_k):
    """benchmark_utils.get_eval_ks (:620-623)."""
    return sorted(k for k in [2, 3, 5, 10, 20, 50, 100] if k <= top_k)


def _first_hit_rank(retrieved, targets):
    """retrieved [nq, top_k] ids (-1 = empty), targets [nq] -> 0-based rank of the first hit, or top_k if none."""
    retrieved = np.asarray(retrieved)
    hit = retrieved == np.asarray(targets)[:, None]
    rank = np.where(hit.any(1), hit.argmax(1), retrieved.shape[1])
    return rank, (retrieved >= 0).sum(1)


def metrics_from_ids(name, retrieved, targets, top_k, seconds):
    """The dict evaluate_retrieval returns (:822-833), from a [nq, >=top_k] id matrix."""
    retrieved = np.asarray(retrieved)[:, :top_k]
    n = len(retrieved)
    rank, n_valid = _first_hit_rank(retrieved, targets)
    out = {}
    disc = 1.0 / np.log2(np.arange(2, top_k + 2))  # discount of positions 1..top_k
    for k in get_eval_ks(top_k):
        within = rank < k
        kk = np.minimum(k, n_valid)  # len(retrieved[:k]) when fewer than k documents came back
        # tie-averaged gain of the kk-1 zero-scored positions 2..kk (sklearn _tie_averaged_dcg)
        tail = np.where(kk > 1, (np.cumsum(disc)[np.maximum(kk, 1) - 1] - disc[0]) / np.maximum(kk - 1, 1), 0.0)
        ndcg = np.where(rank == 0, 1.0, tail)
        out[f"recall@{k}"] = round(float(within.sum()) / n, 4)
        out[f"mrr@{k}"] = round(float((within / (rank + 1.0)).sum()) / n, 4)
        out[f"ndcg@{k}"] = round(float((within * ndcg).sum()) / n, 4)
        out[f"ndcg_std@{k}"] = round(float((within * disc[np.minimum(rank, top_k - 1)]).sum()) / n, 4)
    out["time_taken"] = round(float(seconds), 2)
    out["method"] = name
    out["avg_latency_ms"] = round(1000.0 * float(seconds) / max(n, 1), 2)
    return out


def evaluate_cobweb(wrapper, queries, target_ids, top_k=10, mode="fast", name=None):
    """evaluate_retrieval for a CobwebWrapper: mode "fast" = cobweb_predict_fast semantics (dense index),
    "basic" = cobweb_predict semantics (best-first search), both batched on the device.  target_ids are
    sentence ids (the reference compares document strings; ids are the same test on a corpus without
    duplicate texts)."""
    Q = wrapper.tree._as_device_mat(queries)
    torch.cuda.synchronize()
    t0 = time.time()
    if mode == "fast":
        ids, _ = wrapper.predict_fast_batch(Q, top_k)
        ids = ids.cpu().numpy()
    elif mode == "basic":
        leaves, nfound, _ = wrapper.predict_batch(Q, top_k)
        wrapper._sync_sentence_lists()
        ids = np.full((len(Q), top_k), -1, np.int64)
        for i, row in enumerate(np.asarray(leaves)):
            flat = [s for leaf in row[: int(nfound[i])] for s in wrapper.tree._sent.get(int(leaf), [])]
            ids[i, : min(top_k, len(flat))] = flat[:top_k]
    else:
        raise ValueError("mode must be 'fast' or 'basic'")
    torch.cuda.synchronize()
    return metrics_from_ids(name or f"Cobweb {mode}", ids, target_ids, top_k, time.time() - t0)


def retrieve_dot_batch(corpus_embs, queries, k):
    """retrieve_torch_dot (:602-614) for a batch on the device: X @ Q^T and top-k -- the brute-force inner-product
    baseline the reference's tables list next to FAISS IndexFlatIP (identical metrics).  Library GEMM:
    a baseline, not part of the engine."""
    X = torch.as_tensor(corpus_embs, dtype=torch.float32, device="cuda")
    Q = torch.as_tensor(queries, dtype=torch.float32, device="cuda")
    return torch.topk(Q @ X.T, k, dim=1).indices


def evaluate_dot(corpus_embs, queries, target_ids, top_k=10, name="Torch Dot (GPU, batched)"):
    torch.cuda.synchronize()
    t0 = time.time()
    ids = retrieve_dot_batch(corpus_embs, queries, top_k).cpu().numpy()
    return metrics_from_ids(name, ids, target_ids, top_k, time.time() - t0)
