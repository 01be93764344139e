import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebTorchTree, synth, _lib
n, d, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
x = torch.from_numpy(synth.corpus(n, d, kind, 0)).cuda()
L = _lib.load()
for ncta in [int(v) for v in sys.argv[4:]]:
    L.cw_set_ifit_cluster(ncta)
    t = CobwebTorchTree((d,))
    torch.cuda.synchronize(); t0 = time.time()
    t.ifit_batch(x, tag_sentences=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    c = t.store.counters(); ph = t.store.ifit_phase_cycles(); tot = sum(ph.values())
    print(f"ncta={ncta}: {n/dt:.0f} inserts/s, {dt/c['levels']*1e6:.1f} us/level, rows/insert {c['rows']/n:.0f}", {k: round(100*v/max(tot,1)) for k, v in ph.items() if v})
