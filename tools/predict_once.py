import sys, numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebWrapper, synth
n, d, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synth.corpus(n, d, "unit", 0)
w = CobwebWrapper(corpus=[None]*n, corpus_embeddings=torch.from_numpy(x).cuda())
q, _ = synth.queries(x, nq, "unit", 1); qd = torch.from_numpy(q).cuda()
w.build_prediction_index()
for _ in range(3): w._index.predict(qd, 10)
torch.cuda.synchronize()
