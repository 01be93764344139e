"""time from "sentences added" to "index usable again" with the device-side topology.
  python tools/index_rebuild.py [docs] [dim] [add] """
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, synth  # noqa: E402

n, d, add = (int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 768,
             int(sys.argv[3]) if len(sys.argv) > 3 else 1000)
x = synth.corpus(n + add, d, "unit", seed=0)
w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=torch.from_numpy(x[:n]).cuda())
q, _ = synth.queries(x, 2048, "unit", seed=1)
qd = torch.from_numpy(q).cuda()
w.predict_fast_batch(qd, 10)
torch.cuda.synchronize()
t0 = time.time()
w.add_sentences([None] * add, torch.from_numpy(x[n:]).cuda())
torch.cuda.synchronize()
t_add = time.time() - t0
for rep in range(3):
    w._invalidate_prediction_index()
    torch.cuda.synchronize()
    t0 = time.time()
    w.build_prediction_index()
    torch.cuda.synchronize()
    t_build = time.time() - t0
    t0 = time.time()
    ids, vals = w.predict_fast_batch(qd, 10)
    torch.cuda.synchronize()
    t_q = time.time() - t0
    print(f"rebuild {t_build * 1e3:.1f} ms (nodes {w._index.nn}, mode {w._index.mode}), first batch of 2048 queries {t_q * 1e3:.1f} ms", flush=True)
w.set_dense_mode("fp32")
ids32, vals32 = w.predict_fast_batch(qd, 10)
print(f"ifit of {add} sentences {t_add * 1e3:.0f} ms; ids identical to the FP32 form on the rebuilt index: "
      f"{bool(torch.equal(ids, ids32))} / scores {bool(torch.equal(vals, vals32))}; new sentences retrievable: "
      f"{float(np.mean([n + i in row for i, row in enumerate(w.predict_fast_batch(torch.from_numpy(x[n:n + 256]).cuda(), 10)[0].cpu().numpy())])):.3f}")
