"""Identity check of the tensor modes against the FP32 path on a large whitened (cancellation-heavy) corpus.
usage: python tools/fused_check.py [n_docs] [dim] [n_queries] [kind]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
kind = sys.argv[4] if len(sys.argv) > 4 else "whitened"
x = synth.corpus(n, d, kind, seed=0)
q, _ = synth.queries(x, nq, kind, seed=1)
w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
Q = torch.as_tensor(q, device="cuda")
w.set_dense_mode("fp32")
w.build_prediction_index()
ix = w._index
print(f"{n}x{d} {kind}: {ix.nn} nodes, depth {ix.max_len}")
for k in (10, 1):
    ix.set_mode("fp32")
    i32, v32, _ = ix.predict(Q, k)
    for mode in ("tf32x3", "tf32x3f"):
        ix.set_mode(mode)
        e0, f0 = ix.n_escalated, ix.n_fallback
        ix.predict(Q, k)
        torch.cuda.synchronize()
        t0 = time.time()
        a, b, _ = ix.predict(Q, k)
        torch.cuda.synchronize()
        ms = (time.time() - t0) * 1e3
        print(f"k={k} {mode}: ids identical {bool(torch.equal(a, i32))}, scores bit-identical {bool(torch.equal(b, v32))}, "
              f"{ms:.2f} ms = {nq / ms * 1e3:.0f} q/s, escalated {ix.n_escalated - e0} fallback {ix.n_fallback - f0} (2 passes)")
