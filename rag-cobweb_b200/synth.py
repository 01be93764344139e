"""Seeded synthetic embeddings of the shapes BASELINE.json names.

unit      rows ~ N(0, I) L2-normalised (MiniLM / roberta sentence-transformer shape; the
          reference normalises ST embeddings, src/utils/benchmark_utils.py:339)
whitened  64-centre Gaussian mixture, ~zero mean / unit variance per dim (PCA+ICA shape)
queries   q_j = X[t_j] + 0.05 * N(0, I) (unit: re-normalised); ground truth = t_j
"""
import numpy as np


def corpus(n, d, kind="unit", seed=0):
    rng = np.random.default_rng(seed)
    if kind == "unit":
        x = rng.standard_normal((n, d), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        return np.ascontiguousarray(x, dtype=np.float32)
    if kind == "whitened":
        centres = rng.standard_normal((64, d), dtype=np.float32)
        z = rng.integers(0, 64, size=n)
        x = 0.7 * centres[z] + 0.7 * rng.standard_normal((n, d), dtype=np.float32)
        return np.ascontiguousarray(x, dtype=np.float32)
    raise ValueError(f"unknown corpus kind {kind!r}")


def queries(x, q, kind="unit", seed=1, targets=None):
    """Return (Q[q, d] float32, targets[q] int64)."""
    rng = np.random.default_rng(seed)
    n, d = x.shape
    if targets is None:
        targets = np.arange(q) % n if q <= n else rng.integers(0, n, size=q)
    targets = np.asarray(targets, dtype=np.int64)
    qs = x[targets] + 0.05 * rng.standard_normal((len(targets), d), dtype=np.float32)
    if kind == "unit":
        qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    return np.ascontiguousarray(qs, dtype=np.float32), targets
