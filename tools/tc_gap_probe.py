"""How far apart are rank k and rank k' of the dense leaf scores, compared with the tf32x3-vs-fp32
score difference?  Decides the candidate count / margin of an exact re-score stage."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, synth  # noqa: E402

n, d, nq, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
k = 10
x = synth.corpus(n, d, kind, seed=0)
q, _ = synth.queries(x, nq, kind, seed=1)
w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
w.build_prediction_index()
ix = w._index
Q = torch.as_tensor(q, device="cuda")
ix.set_mode("fp32")
_, _, l32 = ix.predict(Q, 0, want_leaf_scores=True)
ix.set_mode("tf32x3")
_, _, ltc = ix.predict(Q, 0, want_leaf_scores=True)
diff = (ltc - l32).abs()
print(f"{n}x{d} {kind}: leaf score |tf32x3 - fp32| max {diff.max().item():.3e} mean {diff.mean().item():.3e}; "
      f"score magnitude ~{l32.abs().mean().item():.1f}")
s32 = torch.sort(l32, dim=1, descending=True).values
stc = torch.sort(ltc, dim=1, descending=True).values
for kp in (12, 16, 24, 32, 48, 64, 128):
    gap = (s32[:, k - 1] - stc[:, kp]).cpu().numpy()
    qs = np.quantile(gap, [0.0, 0.001, 0.01, 0.05, 0.5])
    print(f"k'={kp}: exact[k-1] - approx[k'] quantiles min/0.1%/1%/5%/50% = " + " ".join(f"{v:.3e}" for v in qs))
