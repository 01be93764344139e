"""Per-stage timing of the fused tensor predict (CUDA events around each stage of DenseIndex._fused_candidates)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, _lib, synth  # noqa: E402

n, d, nq, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
k = 10
x = synth.corpus(n, d, kind, seed=0)
q, _ = synth.queries(x, nq, kind, seed=1)
w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
w.build_prediction_index()
ix = w._index.set_mode("tf32x3f")
Q = torch.as_tensor(q, device="cuda")
ix.predict(Q, k)
torch.cuda.synchronize()
marks = []
orig_check = _lib.check


def check(rc, what=""):
    orig_check(rc, what)
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((what, e))


_lib.check = check
import rag_cobweb_b200.wrapper as W  # noqa: E402
W._lib.check = check
for rep in range(3):
    marks.clear()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    ix.predict(Q, k)
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    torch.cuda.synchronize()
print(f"{n}x{d} {kind}, {nq} queries: fused predict {e0.elapsed_time(e1):.3f} ms; n_int {ix.fx['n_int']}, n_leaf {ix.fx['n_leaf']}, "
      f"sample tiles {ix.fx['n_s']}, levels {len(ix.fx['F']['level_off']) - 1}")
prev = e0
agg = {}
for what, e in marks:
    agg[what] = agg.get(what, 0.0) + prev.elapsed_time(e)
    prev = e
for what, ms in agg.items():
    print(f"  {what or '?'}: {ms:.3f} ms")
print(f"  tail (python, sync, torch ops): {prev.elapsed_time(e1):.3f} ms")
print("appended per query: mean %.1f max %d" % (float(ix._ws['f_cnt'][:nq].float().mean()), int(ix._ws['f_cnt'][:nq].max())))
