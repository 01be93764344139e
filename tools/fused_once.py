"""Two fused predict calls on one tree (warm-up + one): the command the ncu captures of the fused kernels profile.
  python tools/fused_once.py [docs] [dim] [unit|whitened] [queries] [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, DenseIndex, synth  # noqa: E402

n, d, kind, nq, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 768,
                     sys.argv[3] if len(sys.argv) > 3 else "unit", int(sys.argv[4]) if len(sys.argv) > 4 else 10000,
                     int(sys.argv[5]) if len(sys.argv) > 5 else 10)
DenseIndex.TENSOR_MIN_NODES = 0
x = synth.corpus(n, d, kind, seed=0)
w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=torch.from_numpy(x).cuda())
q, _ = synth.queries(x, nq, kind, seed=1)
qd = torch.from_numpy(q).cuda()
w.set_dense_mode("fused")
w.build_prediction_index()
ix = w._index
for _ in range(2):
    ids, vals, _ = ix.predict(qd, k, small=False)
torch.cuda.synchronize()
print("stats", ix.stats, "stages", ix.profile_stages(qd[: ix.fused_workspace(nq, k)["cap_q"]], k))
