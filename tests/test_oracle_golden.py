"""Pins the CPU oracle (oracle/cobweb_oracle.c) against fixtures recorded from the reference
itself (tests/golden/make_golden.py).  CPU only.

What "pinned" means here ("Oracle"):
  * free-running: on the four smaller cases the oracle takes exactly the reference's decisions
    and produces the identical tree, identical leaves, identical best-first retrieval order;
  * guided: on every case (including BASELINE configs[0] and [1]) the reference's recorded
    decisions are replayed; the resulting node statistics are bit-identical, every partition
    utility agrees with the reference's to 1e-4 relative plus the reference's own fp32
    cancellation floor eps*D, and wherever the oracle's own arg-max differs from the recorded
    one the two candidates are closer than that floor (the reference's choice was rounding
    noise of torch's machine-dependent fp32 sum order).
"""
import os

import numpy as np
import pytest

from oracle.cobweb_oracle import OracleTree, default_prior_var, leaf_scores, lib
from rag_cobweb_b200 import synth

CASES = ["tiny_unit_64", "dups_unit_200x32", "unit_300x128", "whitened_600x256", "cfg1_unit_1000x384",
         "cfg2_unit_1500x1024"]
FREE_RUNNING_EXACT = ["tiny_unit_64", "dups_unit_200x32", "unit_300x128", "whitened_600x256"]
EPS32 = float(np.finfo(np.float32).eps)


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, d, kind = int(g["n"]), int(g["d"]), str(g["kind"])
    x = synth.corpus(n, d, kind, seed=0)
    if name.startswith("dups"):
        x[50:60] = x[10:20]
        x[150:155] = x[10:15]
    q, _ = synth.queries(x, g["rank_scores"].shape[0], kind, seed=1)
    return g, x, q


def bfs_pos(b):
    pos = np.full(int(b["order"].max()) + 1, -1, np.int64)
    pos[b["order"]] = np.arange(len(b["order"]))
    return pos


def check_tree(t, leaves, g):
    b = t.bfs()
    assert np.array_equal(b["parent"], g["bfs_parent"])
    assert np.array_equal(b["count"], g["bfs_count"])
    assert np.array_equal(b["nchild"], g["bfs_nchild"])
    pos = bfs_pos(b)
    assert np.array_equal(pos[leaves], g["leaf_of_sentence"])
    mean, m2 = t.rows(b["order"])
    assert np.array_equal(mean[0], g["mean_row0"]) and np.array_equal(m2[0], g["m2_row0"])
    np.testing.assert_allclose(mean.astype(np.float64).sum(1), g["mean_sum"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(m2.astype(np.float64).sum(1), g["m2_sum"], rtol=0, atol=1e-12)
    return b, pos


def check_queries(t, pos, g, q):
    k = int(g["k"])
    t.build_index()
    ns, ls = t.dense_scores(q)
    np.testing.assert_allclose(ns, g["node_scores"], rtol=1e-5)  # whitened: torch fp32 sum of ~1e4-sized terms
    np.testing.assert_allclose(ls, g["rank_scores"], rtol=2e-6)
    # the path product alone, fed the reference's own node scores, is bit-exact
    assert np.array_equal(leaf_scores(g["node_scores"], t.index["path_idx"], t.index["path_w"]), g["rank_scores"])
    for i in range(len(q)):  # top-k ids of cobweb_predict_fast (noise-free)
        want = np.argsort(-g["rank_scores"][i], kind="stable")[:k]
        got = np.argsort(-ls[i], kind="stable")[:k]
        if not np.array_equal(want, got):  # only allowed to differ inside an fp32 tie
            s = g["rank_scores"][i]
            assert set(want) ^ set(got) <= set(np.nonzero(np.abs(s - s[want[-1]]) <= 4e-6 * abs(s[want[-1]]))[0])
    lv, nf, _, calls = t.categorize(q, k=k, max_nodes=100000)
    assert np.array_equal(pos[lv], g["bf_leaves"])
    assert np.array_equal(calls, g["bf_lp_calls"])
    assert (nf == k).all()
    _, _, best, _ = t.categorize(q, k=0)
    assert np.array_equal(pos[best], g["cat_best"])
    root = t.bfs()["order"][0]
    rl = np.array([t.log_prob(root, qq) for qq in q], np.float32)
    np.testing.assert_allclose(rl, g["root_lp"], rtol=1e-6)


def test_prior_var_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_unit_64.npz"))
    assert float(g["prior_var"]) == default_prior_var()


def test_logf_accuracy():
    rng = np.random.default_rng(3)
    xs = np.concatenate([np.exp(rng.uniform(-20, 20, 20000)), [1.0, 0.058549832, 1e-38, 3e38]]).astype(np.float32)
    got = np.array([lib().co_logf(float(v)) for v in xs], np.float32)
    ref = np.log(xs.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert np.max(np.abs(got - ref) / np.maximum(ulp, 1e-45)) < 1.0
    assert lib().co_logf(0.0) == -np.inf and np.isnan(lib().co_logf(-1.0))


@pytest.mark.parametrize("name", FREE_RUNNING_EXACT)
def test_free_running_matches_reference(golden_dir, name):
    g, x, q = load_case(golden_dir, name)
    t = OracleTree(x.shape[1])
    leaves, tr, off = t.ifit(x, trace=True)
    assert np.array_equal(tr, g["ops"])
    assert np.array_equal(off, g["ops_off"])
    _, pos = check_tree(t, leaves, g)
    check_queries(t, pos, g, q)


@pytest.mark.parametrize("name", CASES)
def test_guided_replay_matches_reference(golden_dir, name):
    g, x, q = load_case(golden_dir, name)
    d = x.shape[1]
    t = OracleTree(d)
    dec = g["ops"][g["ops"] < 4]
    leaves, pus, st = t.ifit_guided(x, dec, g["dec_b1"], g["dec_b2"])
    assert st["used"] == len(dec)
    _, pos = check_tree(t, leaves, g)
    # partition utilities: 1e-4 relative + the reference's own cancellation floor (s2 ~ D in fp32)
    assert np.array_equal(np.isnan(pus), np.isnan(g["dec_pus"]))
    m = ~np.isnan(pus)
    assert np.all(np.abs(pus[m] - g["dec_pus"][m]) <= 1e-4 * np.abs(g["dec_pus"][m]) + EPS32 * d)
    # the oracle's own choice differs only inside that floor, and never on the operation
    assert st["op_disagree"] == 0
    assert st["rank_disagree"] <= 0.05 * len(dec)
    assert st["rank_margin"] <= 0.25 * EPS32 * d
    check_queries(t, pos, g, q)


def test_load_roundtrip(golden_dir):
    g, x, q = load_case(golden_dir, "unit_300x128")
    t = OracleTree(x.shape[1])
    t.ifit(x)
    b = t.bfs()
    mean, m2 = t.rows(b["order"])
    t2 = OracleTree(x.shape[1])
    t2.load(b["parent"], b["count"], b["nsent"], mean, m2)
    lv1, _, _, c1 = t.categorize(q, k=5)
    lv2, _, _, c2 = t2.categorize(q, k=5)
    assert np.array_equal(bfs_pos(b)[lv1], lv2) and np.array_equal(c1, c2)


def test_oracle_greedy_mode_matches_reference(golden_dir):
    """COBWEB_GREEDY_MODE = True (src/utils/constants.py; CobwebTorchTree.py:209-213): "new" at every internal node.
    Fixture recorded by running the reference with the switch on (make_golden.py greedy)."""
    g = np.load(os.path.join(golden_dir, "greedy_unit_150x24.npz"))
    x = synth.corpus(150, 24, "unit", seed=0)
    x[40:50] = x[5:15]
    t = OracleTree(24, greedy=True)
    t.ifit(x)
    b = t.bfs()
    assert np.array_equal(b["parent"], g["bfs_parent"]) and np.array_equal(b["count"], g["bfs_count"])
    assert np.array_equal(b["nchild"], g["bfs_nchild"]) and b["depth"].max() == 1   # a flat tree: no duplicate is matched
    mean, m2 = t.rows(b["order"])
    assert np.array_equal(mean, g["mean"]) and np.array_equal(m2, g["m2"])


def test_reference_json_document_roundtrip(golden_dir):
    """The reference's own dump_json output (fixture written by make_golden.py json): load_tree_json -> dump_tree_json
    reproduces the document byte for byte (fp32 -> shortest decimal -> fp32 is lossless), and the binary snapshot
    carries the same content."""
    from rag_cobweb_b200 import serialize
    doc = open(os.path.join(golden_dir, "reference_tree_80x12.json")).read()
    params, parent, count, mean, m2, sids = serialize.load_tree_json(doc)
    assert serialize.dump_tree_json(params, parent, count, mean, m2, sids) == doc
    assert len(parent) == 112 and sorted(s for l in sids for s in l) == list(range(80))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "t.cwb")
        leaf = np.zeros(80, np.int32)
        for node, l in enumerate(sids):
            for s in l:
                leaf[s] = node
        serialize.write_snapshot(path, params, parent, count, [len(l) for l in sids], lambda lo, hi: (mean[lo:hi], m2[lo:hi]),
                                 leaf, extra={"n_sentences": 80})
        snap = serialize.read_snapshot(path)
        assert snap["params"] == params and np.array_equal(snap["parent"], parent) and np.array_equal(snap["count"], count)
        assert np.array_equal(snap["mean"], mean) and np.array_equal(snap["m2"], m2) and np.array_equal(snap["leaf_of_sentence"], leaf)
        assert os.path.getsize(path) < 0.3 * len(doc)
