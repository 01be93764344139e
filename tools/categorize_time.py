"""Best-first categorize throughput on a tree built here: python tools/categorize_time.py [n] [d] [nq] [k]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebTorchTree, synth
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 30000, int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq, k = int(sys.argv[3]) if len(sys.argv) > 3 else 10000, int(sys.argv[4]) if len(sys.argv) > 4 else 10
x = torch.from_numpy(synth.corpus(n, d, "unit", 0)).cuda()
t = CobwebTorchTree((d,))
t.ifit_batch(x, tag_sentences=True)
q = x[:nq] + 0.05 * torch.randn(nq, d, device="cuda")
for mn in (100, 1000):
    t.categorize_batch(q, retrieve_k=k, max_nodes=mn)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        out = t.categorize_batch(q, retrieve_k=k, max_nodes=mn)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 3
    print(f"{n}x{d}, {nq} queries, k={k}, max_nodes={mn}: {nq / dt:.0f} q/s ({dt * 1e3:.2f} ms)")
