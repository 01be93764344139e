#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box); the .npz files it writes are committed and are what the tests read.  Nothing from the
reference is copied: it is imported, driven on seeded synthetic inputs
(rag-cobweb_b200/synth.py) and its outputs are recorded.

Recorded per case (all trees are described in BFS order, children in list order, exactly the
numbering CobwebWrapper.build_prediction_index uses, src/cobweb/CobwebWrapper.py:107-132):
  ops, ops_off      per-insert decision trace of CobwebTorchTree.cobweb (CobwebTorchTree.py:182-232)
                    codes: 0 best, 1 new, 2 merge, 3 split, 4 leaf-increment, 5 fringe-split
  op_gap            per internal decision: relative gap between the two best partition utilities
  dec_b1, dec_b2    per internal decision: position of best1 / best2 in the node's child list (-1: none)
  dec_pus           per internal decision: partition utility of best/new/merge/split (NaN: not a candidate)
  bfs_parent/count/nchild, leaf_of_sentence, mean_sum/m2_sum (float64 row checksums),
  mean_row0/m2_row0 (root statistics in full)
  queries idx -> rank_scores (cobweb_rank_scores, CobwebWrapper.py:267), node_scores,
  bf_leaves (cobweb_predict pop order, CobwebWrapper.py:435 / CobwebTorchTree.py:235),
  bf_visited, cat_best (categorize(x) over the whole tree)

usage: python tests/golden/make_golden.py [case ...]
"""
import importlib.util
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

CASES = {
    # name: (n, d, kind, n_queries, k)
    "tiny_unit_64": (120, 64, "unit", 8, 5),
    "unit_300x128": (300, 128, "unit", 12, 10),
    "cfg1_unit_1000x384": (1000, 384, "unit", 16, 10),
    "cfg2_unit_1500x1024": (1500, 1024, "unit", 16, 10),
    "whitened_600x256": (600, 256, "whitened", 16, 10),
    "dups_unit_200x32": (200, 32, "unit", 8, 5),
}


def load_synth():
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "rag-cobweb_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def import_reference():
    sys.path.insert(0, REF)
    g = types.ModuleType("graphviz")
    g.Digraph = object
    sys.modules["graphviz"] = g
    from src.cobweb.CobwebTorchNode import CobwebTorchNode
    from src.cobweb.CobwebTorchTree import CobwebTorchTree
    from src.cobweb.CobwebWrapper import CobwebWrapper
    return CobwebTorchNode, CobwebTorchTree, CobwebWrapper


def bfs(root):
    order, parent = [root], [-1]
    i = 0
    while i < len(order):
        for c in order[i].children:
            order.append(c)
            parent.append(i)
        i += 1
    return order, np.asarray(parent, dtype=np.int32)


def run_case(name, n, d, kind, nq, k):
    synth = load_synth()
    Node, Tree, Wrapper = import_reference()
    x = synth.corpus(n, d, kind, seed=0)
    if name.startswith("dups"):
        # exact duplicates exercise the leaf-increment branch (CobwebTorchNode.is_exact_match)
        x[50:60] = x[10:20]
        x[150:155] = x[10:15]
    q, targets = synth.queries(x, nq, kind, seed=1)

    trace, gaps = [], []
    dec_b1, dec_b2, dec_pus = [], [], []
    cur = []
    orig_gbo = Node.get_best_operation

    def gbo(self, instance, best1, best2, best1_pu):
        # same candidate list as the reference builds, recorded before it picks
        pus = [float(best1_pu), float(self.pu_for_new_child(instance))]
        row = [pus[0], pus[1], float("nan"), float("nan")]
        if len(self.children) > 2 and best2:
            pus.append(float(self.pu_for_merge(best1, best2, instance)))
            row[2] = pus[-1]
        if len(best1.children) > 0:
            pus.append(float(self.pu_for_split(best1)))
            row[3] = pus[-1]
        dec_pus.append(row)
        dec_b1.append(self.children.index(best1))
        dec_b2.append(self.children.index(best2) if best2 else -1)
        s = sorted(pus, reverse=True)
        gaps.append(abs(s[0] - s[1]) / max(abs(s[0]), 1e-30))
        res = orig_gbo(self, instance, best1, best2, best1_pu)
        cur.append({"best": 0, "new": 1, "merge": 2, "split": 3}[res[1]])
        return res

    Node.get_best_operation = gbo
    torch.set_num_threads(1)

    # Build through the wrapper exactly as benchmark_utils.load_cobweb_model does
    # (CobwebWrapper(corpus=..., corpus_embeddings=...)); hook ifit to delimit inserts.
    orig_ifit = Tree.ifit

    def ifit(self, instance):
        cur.clear()
        res = orig_ifit(self, instance)
        codes = list(cur)
        # classify the terminal event: a trace not ending in 'new' ended at a leaf, which was
        # either incremented (empty tree / exact match) or fringe-split (new leaf of count 1)
        if not codes or codes[-1] != 1:
            codes.append(4 if (len(trace) == 0 or res.count.item() > 1) else 5)
        trace.append(codes)
        return res

    Tree.ifit = ifit
    t0 = time.time()
    w = Wrapper(corpus=[None] * n, corpus_embeddings=torch.from_numpy(x))
    build_s = time.time() - t0
    Tree.ifit = orig_ifit
    Node.get_best_operation = orig_gbo

    order, parent = bfs(w.tree.root)
    idx_of = {id(nd): i for i, nd in enumerate(order)}
    count = np.asarray([nd.count.item() for nd in order], dtype=np.float32)
    nchild = np.asarray([len(nd.children) for nd in order], dtype=np.int32)
    mean_sum = np.asarray([nd.mean.double().sum().item() for nd in order])
    m2_sum = np.asarray([nd.meanSq.double().sum().item() for nd in order])
    leaf_of_sentence = np.asarray([idx_of[id(w.sentence_to_node[i])] for i in range(n)], dtype=np.int32)

    ops = np.asarray([c for t in trace for c in t], dtype=np.int8)
    ops_off = np.zeros(n + 1, dtype=np.int64)
    ops_off[1:] = np.cumsum([len(t) for t in trace])

    # ---- queries
    w.build_prediction_index()
    rank = np.stack([w.cobweb_rank_scores(torch.from_numpy(qq), is_embedding=True).numpy() for qq in q])
    xq = torch.from_numpy(q)
    node_scores = np.stack([
        (-0.5 * (torch.log(w._node_vars).sum(dim=1)
                 + (((xx.unsqueeze(0) - w._node_means) ** 2) / w._node_vars).sum(dim=1))).numpy()
        for xx in xq])
    bf_leaves, bf_visited, cat_best = [], [], []
    lp_calls = [0]
    orig_lp = Node.log_prob

    def lp(self, instance):
        lp_calls[0] += 1
        return orig_lp(self, instance)

    Node.log_prob = lp
    for qq in xq:
        lp_calls[0] = 0
        leaves = w.tree.categorize(qq, use_best=True, max_nodes=w.max_init_search, retrieve_k=k)
        bf_leaves.append([idx_of[id(l)] for l in leaves])
        bf_visited.append(lp_calls[0])
        best = w.tree.categorize(qq)
        cat_best.append(idx_of[id(best)])
    Node.log_prob = orig_lp

    # root log_prob (includes the 2*pi term, CobwebTorchNode.py:100-104) for a scalar check
    root_lp = np.asarray([w.tree.root.log_prob(qq).item() for qq in xq], dtype=np.float32)
    # partition utility of the root and a few compute_score values as known-answer scalars
    root = w.tree.root
    pu_root = float(root.partition_utility())
    cs = [float(w.tree.compute_score(c.mean, c.var, root.mean, root.var)) for c in root.children]

    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(
        out, n=n, d=d, kind=kind, k=k, build_seconds=build_s,
        ops=ops, ops_off=ops_off, op_gap=np.asarray(gaps, dtype=np.float64),
        dec_b1=np.asarray(dec_b1, dtype=np.int32), dec_b2=np.asarray(dec_b2, dtype=np.int32),
        dec_pus=np.asarray(dec_pus, dtype=np.float32),
        bfs_parent=parent, bfs_count=count, bfs_nchild=nchild, leaf_of_sentence=leaf_of_sentence,
        mean_sum=mean_sum, m2_sum=m2_sum,
        mean_row0=order[0].mean.numpy(), m2_row0=order[0].meanSq.numpy(),
        q_targets=targets, rank_scores=rank.astype(np.float32), node_scores=node_scores.astype(np.float32),
        bf_leaves=np.asarray(bf_leaves, dtype=np.int32), bf_lp_calls=np.asarray(bf_visited, dtype=np.int32),
        cat_best=np.asarray(cat_best, dtype=np.int32), root_lp=root_lp,
        pu_root=pu_root, root_child_scores=np.asarray(cs, dtype=np.float32),
        prior_var=float(w.tree.prior_var),
    )
    hist = np.bincount(ops, minlength=6)
    print(f"{name}: n={n} d={d} nodes={len(order)} build={build_s:.1f}s ({n / build_s:.1f} ins/s) "
          f"ops best/new/merge/split/leaf/fringe={hist.tolist()} min_gap={min(gaps):.2e} "
          f"gaps<1e-6: {int((np.asarray(gaps) < 1e-6).sum())} -> {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


if __name__ == "__main__":
    extra = ("whitening", "greedy", "json")
    names = [a for a in sys.argv[1:] if a not in extra] or ([] if any(a in extra for a in sys.argv[1:]) else list(CASES))
    for nm in names:
        run_case(nm, *CASES[nm])


def run_whitening_case():
    """PCAICAWhiteningModel.fit / transform of the reference (src/whitening/pca_ica.py) on seeded
    clustered data: records the fitted parameters and the reference's transform of held-out rows."""
    sys.path.insert(0, REF)
    from src.whitening.pca_ica import PCAICAWhiteningModel
    rng = np.random.default_rng(7)
    centres = rng.standard_normal((12, 96)).astype(np.float32)
    mix = rng.standard_normal((96, 96)).astype(np.float32) * 0.3 + np.eye(96, dtype=np.float32)
    X = (centres[rng.integers(0, 12, 1500)] + 0.5 * rng.standard_normal((1500, 96)).astype(np.float32)) @ mix
    X = X.astype(np.float32)
    model = PCAICAWhiteningModel.fit(X[:1200], pca_dim=32)
    held = X[1200:1264]
    out = os.path.join(HERE, "whitening_pcaica.npz")
    np.savez_compressed(out, mean=model.mean, pca_components=model.pca_components,
                        pca_explained_var=model.pca_explained_var, ica_unmixing=model.ica_unmixing, eps=model.eps,
                        x=held, y_ica=model.transform(held), y_pca=model.transform(held, is_ica=False),
                        y_single=model.transform(held[0]))
    print("whitening:", {k: (v.dtype, v.shape) for k, v in np.load(out).items()})


if "whitening" in sys.argv[1:]:
    run_whitening_case()


def run_greedy_case():
    """COBWEB_GREEDY_MODE = True (src/utils/constants.py): the reference's ifit takes "new" at every internal node
    (CobwebTorchTree.py:209-213).  Records the resulting tree for 150 x 24 unit rows with 10 duplicated rows."""
    synth = load_synth()
    Node, Tree, Wrapper = import_reference()
    import src.cobweb.CobwebTorchNode as node_mod
    import src.cobweb.CobwebTorchTree as tree_mod
    node_mod.COBWEB_GREEDY_MODE = tree_mod.COBWEB_GREEDY_MODE = True
    try:
        x = synth.corpus(150, 24, "unit", seed=0)
        x[40:50] = x[5:15]
        t = Tree((24,), device="cpu")
        for row in x:
            t.ifit(torch.tensor(row))
        order, parent = bfs(t.root)
        out = os.path.join(HERE, "greedy_unit_150x24.npz")
        np.savez_compressed(out, bfs_parent=parent, bfs_count=np.asarray([float(n.count) for n in order], np.float32),
                            bfs_nchild=np.asarray([len(n.children) for n in order], np.int32),
                            mean=np.stack([n.mean.numpy() for n in order]).astype(np.float32),
                            m2=np.stack([n.meanSq.numpy() for n in order]).astype(np.float32))
        print(f"greedy: nodes={len(order)} root children={len(t.root.children)} -> {out}")
    finally:
        node_mod.COBWEB_GREEDY_MODE = tree_mod.COBWEB_GREEDY_MODE = False


def run_json_case():
    """The reference's own wire format: CobwebTorchTree.dump_json (CobwebTorchTree.py:67-81) of an 80 x 12 tree whose
    leaves carry sentence ids, written verbatim to reference_tree_80x12.json, plus what the reference answers on it
    (rank scores of 6 queries) so that load -> predict can be checked against the reference."""
    synth = load_synth()
    Node, Tree, Wrapper = import_reference()
    x = synth.corpus(80, 12, "unit", seed=0)
    x[30:33] = x[4]
    w = Wrapper(corpus=[f"s{i}" for i in range(80)], corpus_embeddings=x, encode_func=lambda s: s)
    doc = w.tree.dump_json()
    with open(os.path.join(HERE, "reference_tree_80x12.json"), "w") as f:
        f.write(doc)
    q, _ = synth.queries(x, 6, "unit", seed=1)
    rank = np.stack([w.cobweb_rank_scores(torch.tensor(qq), is_embedding=True).detach().numpy() for qq in q]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_tree_80x12_answers.npz"), rank_scores=rank,
                        leaf_count=np.asarray([float(w.sentence_to_node[i].count) for i in range(80)], np.float32))
    print(f"json: {len(doc)} bytes, rank scores {rank.shape}")


if "greedy" in sys.argv[1:]:
    run_greedy_case()
if "json" in sys.argv[1:]:
    run_json_case()
