import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebTorchTree, synth
n, d, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
x = torch.from_numpy(synth.corpus(n, d, kind, 0)).cuda()
t = CobwebTorchTree((d,))
t.IFIT_CHUNK = n
torch.cuda.synchronize(); t0 = time.time()
t.ifit_batch(x, tag_sentences=True)
torch.cuda.synchronize(); print(f"{n/(time.time()-t0):.0f} inserts/s")
