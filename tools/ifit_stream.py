"""BASELINE configs[4] at full size: streaming ifit of N PCA/ICA-whitened-shape 256-d embeddings, inserts/s.
usage: python tools/ifit_stream.py [n=1000000] [d=256] -> one JSON line"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebTorchTree, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
x = torch.from_numpy(synth.corpus(n, d, "whitened", seed=0)).cuda()
t = CobwebTorchTree((d,))
torch.cuda.synchronize()
t0 = time.time()
marks = []
step = 100000
for lo in range(0, n, step):
    t.ifit_batch(x[lo:lo + step], tag_sentences=True)
    torch.cuda.synchronize()
    marks.append(time.time() - t0)
dt = marks[-1]
c = t.store.counters()
b = t.bfs()
print(json.dumps({"workload": f"configs[4] streaming ifit, {n} whitened {d}-d inserts", "inserts_per_s": n / dt, "seconds": dt,
                  "seconds_per_100k": [round(m - (marks[i - 1] if i else 0.0), 2) for i, m in enumerate(marks)],
                  "nodes": int(len(b["order"])), "depth": int(b["depth"].max()) + 1 if "depth" in b else None,
                  "levels_per_insert": c["levels"] / n, "rows_per_insert": c["rows"] / n,
                  "us_per_level_step": dt / c["levels"] * 1e6, "store_bytes": t.store.bytes()}))
