"""Ad-hoc GPU probe: ifit throughput and dense-kernel time at a given shape."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebWrapper, synth, _lib
import ctypes as C

n, d, nq, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
x = synth.corpus(n, d, kind, 0)
xd = torch.from_numpy(x).cuda()
torch.cuda.synchronize()
t0 = time.time()
w = CobwebWrapper(corpus=[None]*n, corpus_embeddings=xd)
torch.cuda.synchronize()
dt = time.time() - t0
c = w.tree.store.counters()
print(f"ifit {n}x{d} {kind}: {dt:.2f}s = {n/dt:.0f} inserts/s; levels/insert {c['levels']/n:.2f} rows/insert {c['rows']/n:.1f} scores/insert {c['scores']/n:.1f} us/level {dt/c['levels']*1e6:.2f}")
ph = w.tree.store.ifit_phase_cycles(); tot=sum(ph.values()); print("ifit phases (% of lead-CTA cycles):", {k: round(100*v/tot,1) for k,v in ph.items()}, "cycles/level", round(tot/c["levels"]))
b = w.tree.bfs(); print("nodes", len(b['order']), "depth", b['depth'].max(), "max children", b['nchild'].max(), "mean children", b['nchild'][b['nchild']>0].mean())
t0=time.time(); w.build_prediction_index(); torch.cuda.synchronize(); print(f"index build {time.time()-t0:.2f}s nn={w._index.nn} max_len={w._index.max_len}")
q,_ = synth.queries(x, nq, kind, 1); qd = torch.from_numpy(q).cuda()
ix = w._index
for name, fn in [("node_scores", lambda: ix.node_scores(qd[:ix.chunk_queries()])), ("predict k=10", lambda: ix.predict(qd, 10))]:
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); 
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    nqq = min(nq, ix.chunk_queries()) if name=="node_scores" else nq
    print(f"{name}: {ms:.2f} ms for {nqq} queries -> {nqq/ms*1e3:.0f} q/s; {2*2*nqq*ix.nn*d/ms/1e9:.1f} TFLOP/s (4 flop/elem)")
# ffma peak
L=_lib.load(); sink=torch.zeros(4, device='cuda')
for threads in (256, 512, 1024):
    blocks=148*(2048//threads); iters=20000
    L.cw_ffma_peak(blocks, threads, 10, sink.data_ptr(), None); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); L.cw_ffma_peak(blocks, threads, iters, sink.data_ptr(), None); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1); print(f"ffma peak threads={threads}: {2*blocks*threads*iters*64/ms/1e9:.1f} TFLOP/s")
# best-first
t0=time.time(); r = w.tree.categorize_batch(qd[:2048], retrieve_k=10, max_nodes=100000); torch.cuda.synchronize(); dt=time.time()-t0
print(f"best-first 2048 queries: {dt*1e3:.1f} ms -> {2048/dt:.0f} q/s, mean lp_calls {r['lp_calls'].float().mean().item():.0f}")
