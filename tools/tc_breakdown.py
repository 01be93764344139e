"""Per-kernel timing of the tf32x3 dense predict (CUDA events): tensor scores, path/top-k at k and kc, re-score."""
import ctypes as C
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
ix = w._index.set_mode("tf32x3")
Q = torch.as_tensor(q, device="cuda")
L = _lib.load()
ws = ix.workspace(nq, k)
kc = ix.candidates(k)
tx = ix.tx
sids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
vals = torch.empty((nq, k), dtype=torch.float32, device="cuda")


def timed(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)


def paths(kk, so, vo):
    _lib.check(L.cw_dense_paths_topk(C.byref(ix.ix), ws["scores"].data_ptr(), ws["ldq"], nq, kk, None, so.data_ptr(),
                                     vo.data_ptr(), ws["scratch"].data_ptr(), _lib.stream_ptr()))


def rescore():
    _lib.check(L.cw_dense_rescore(w.tree.store.struct(), C.byref(ix.ix), tx.rows, tx.pos_of_sid, Q.data_ptr(), nq, kc,
                                  ws["cand_sid"].data_ptr(), ws["cand_val"].data_ptr(), k, tx.hmax, tx.lmax, tx.wfac,
                                  tx.eps_scale, sids.data_ptr(), vals.data_ptr(), ws["fail"].data_ptr(), _lib.stream_ptr()))


print(f"{n}x{d} {kind}, {nq} queries, {ix.nn} nodes, depth {ix.max_len}, kc={kc}")
timed("tensor node scores", lambda: ix._node_scores_call(Q, nq, ws))
timed(f"paths + top-{k}", lambda: paths(k, ws["sid"], ws["val"]))
timed(f"paths + top-{kc}", lambda: paths(kc, ws["cand_sid"], ws["cand_val"]))
timed("re-score", rescore)
print("flagged:", int(ws["fail"][0]))
timed("predict (all)", lambda: ix.predict(Q, k))
