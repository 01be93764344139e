"""GPU probe of the tensor-core scoring kernel (cw_tensor.cu): accuracy against the FP32-pipe
kernel and a binary64 numpy evaluation, then timing of both kernels.
usage: python tools/tc_probe.py [n_docs] [dim] [n_queries] [kind]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_cobweb_b200 import CobwebWrapper, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    kind = sys.argv[4] if len(sys.argv) > 4 else "unit"
    k = 10
    x = synth.corpus(n, d, kind, seed=0)
    q, targets = synth.queries(x, nq, kind, seed=1)
    t0 = time.time()
    w = CobwebWrapper(corpus=[None] * n, corpus_embeddings=x)
    w.build_prediction_index()
    ix = w._index
    torch.cuda.synchronize()
    print(f"tree {n}x{d} ({kind}): {ix.nn} nodes, depth {ix.max_len}, built in {time.time() - t0:.1f}s", flush=True)
    Q = torch.as_tensor(q, device="cuda")

    ix.set_mode("fp32")
    s32 = ix.node_scores(Q).clone()
    ids32, v32, _ = ix.predict(Q, k)
    ids32, v32 = ids32.clone(), v32.clone()
    ix.set_mode("tf32x3")
    torch.cuda.synchronize()
    print("tensor operands built", flush=True)
    stc = ix.node_scores(Q).clone()
    torch.cuda.synchronize()
    idstc, vtc, _ = ix.predict(Q, k)
    torch.cuda.synchronize()

    # binary64 truth for a sample of queries
    mean, m2 = w.tree.store.rows(ix.order_host)
    cnt = w.tree.store.count[torch.as_tensor(np.asarray(ix.order_host, np.int64), device="cuda")].cpu().numpy()
    pv = float(w.tree.store.prior_var)
    var = np.where(cnt[:, None] > 0, m2.astype(np.float64) / np.maximum(cnt[:, None], 1) + pv, pv)
    qs = q[: min(nq, 16)].astype(np.float64)
    truth = np.stack([-0.5 * (np.log(var).sum(1) + (((qq[None, :] - mean.astype(np.float64)) ** 2) / var).sum(1)) for qq in qs])
    a32 = s32[: len(qs)].cpu().numpy().astype(np.float64)
    atc = stc[: len(qs)].cpu().numpy().astype(np.float64)
    scale = np.abs(truth)
    print(f"vs binary64: fp32 kernel max rel {np.max(np.abs(a32 - truth) / scale):.3e} max abs {np.max(np.abs(a32 - truth)):.3e}; "
          f"tf32x3 max rel {np.max(np.abs(atc - truth) / scale):.3e} max abs {np.max(np.abs(atc - truth)):.3e}", flush=True)
    dd = (stc - s32).abs()
    print(f"tf32x3 vs fp32 kernel, all {stc.numel()} scores: max abs {dd.max().item():.3e}, "
          f"max rel {(dd / s32.abs()).max().item():.3e}", flush=True)
    print(f"re-score: candidates/query {ix.candidates(k)}, queries sent to the FP32 fallback {ix.n_fallback}; "
          f"ids identical {bool((idstc == ids32).all())}, scores bit-identical {bool((vtc == v32).all())}", flush=True)
    hs, hv = ix.predict_host(q, k)
    print(f"host call: ids identical {bool((hs.cuda() == ids32).all())}, scores bit-identical {bool((hv.cuda() == v32).all())}, "
          f"fallbacks so far {ix.n_fallback}", flush=True)
    same = (idstc == ids32).all(1).float().mean().item()
    sets = np.mean([len(set(a) & set(b)) / k for a, b in zip(idstc.cpu().numpy(), ids32.cpu().numpy())])
    rec32 = float(np.mean([t in g for t, g in zip(targets, ids32.cpu().numpy())]))
    rectc = float(np.mean([t in g for t, g in zip(targets, idstc.cpu().numpy())]))
    print(f"top-{k}: identical ordered lists {same:.4f}, set agreement {sets:.4f}, recall fp32 {rec32:.4f} tf32x3 {rectc:.4f}; "
          f"max |score diff| {(vtc - v32).abs().max().item():.3e}", flush=True)

    # timing
    for mode in ("fp32", "tf32x3"):
        ix.set_mode(mode)
        ws = ix.workspace(nq, k)
        for _ in range(2):
            ix._node_scores_call(Q, nq, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ix._node_scores_call(Q, nq, ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 4.0 * nq * ix.nn * d
        print(f"{mode}: node scores {ms:.3f} ms  ({fl / ms / 1e9:.1f} algorithmic TFLOP/s, 4*Q*Nn*D)", flush=True)
        e0.record()
        for _ in range(reps):
            ix.predict(Q, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{mode}: predict (scores + paths + top-k) {ms:.3f} ms = {nq / ms * 1e3:.0f} q/s", flush=True)


if __name__ == "__main__":
    main()
