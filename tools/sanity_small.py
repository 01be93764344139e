"""Small end-to-end exercise of every kernel (for compute-sanitizer runs)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebWrapper, synth
for n, d, kind in ((300, 64, "unit"), (200, 300, "whitened")):
    x = synth.corpus(n, d, kind, 0)
    w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
    q, _ = synth.queries(x, 40, kind, 1)
    ids, _ = w.predict_fast_batch(q, 10)
    w.rank_scores_batch(q[:5])
    lv, nf, calls = w.predict_batch(q, 5)
    w.tree.categorize_batch(q[:8])
    hs, hv = w._index.predict_host(q, 10)
    assert np.array_equal(hs.numpy(), ids.cpu().numpy())
    print(n, d, kind, "ok", int(calls.sum()))
