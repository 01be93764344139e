"""Small end-to-end exercise of every kernel (the command of the compute-sanitizer runs kept under profiles/):
ifit (cluster kernel), best-first categorize, FP32 dense predict, the fused tcgen05 pipeline (fp16x3 internal rows,
fp16 filter, finish, device-side fallback, audit), the exact small-batch path, rank-score gradient, whitening."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebWrapper, DenseIndex, synth  # noqa: E402

DenseIndex.TENSOR_MIN_NODES = 0
for n, d, kind in ((300, 64, "unit"), (700, 100, "whitened")):
    x = synth.corpus(n, d, kind, 0)
    x[20:24] = x[3]
    w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
    q, _ = synth.queries(x, 300, kind, 1)
    w.set_dense_mode("fp32")
    ids, vals = w.predict_fast_batch(q, 10)
    w.rank_scores_batch(q[:5])
    lv, nf, calls = w.predict_batch(q[:40], 5)
    w.tree.categorize_batch(q[:8])
    w.set_dense_mode("fused")
    w._index.audit_every = 64
    fi, fv = w.predict_fast_batch(q, 10)
    assert torch.equal(fi, ids) and torch.equal(fv, vals), "fused != fp32"
    hs, hv = w._index.predict_host(q, 10)
    assert np.array_equal(hs.numpy(), ids.cpu().numpy())
    si, sv = w.predict_fast_batch(q[:7], 10)
    assert torch.equal(si, ids[:7]) and torch.equal(sv, vals[:7]), "small path != fp32"
    qg = torch.from_numpy(q[:4]).cuda().requires_grad_(True)
    w.rank_scores_batch(qg).sum().backward()
    print(n, d, kind, "ok", int(calls.sum()), w._index.stats)
