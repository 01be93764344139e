"""JSON wire format of the reference (pure Python / numpy, runs anywhere).

Tree document (CobwebTorchTree.dump_json, src/cobweb/CobwebTorchTree.py:67-81 with
CobwebTorchNode.iterative_output_json, src/cobweb/CobwebTorchNode.py:741-772):
  {"use_info":..,"acuity_cutoff":..,"use_kl":..,"shape":[D],"alpha":..,"prior_var":..,
   "root": {"count": c, "mean": [...], "meanSq": [...], "sentence_id": [...], "children": [ ... ]}}
Documents written here load in the reference's CobwebTorchTree.load_json and vice versa
(the reference's loader reverses child order; ours keeps it).
"""
import json

import numpy as np


def dump_tree_json(params, parent, count, mean, m2, sentence_ids):
    """Nodes are given in BFS order (parent[i] < i, siblings in list order)."""
    n = len(parent)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[int(parent[i])].append(i)

    def node_head(i):
        d = {"count": float(count[i]), "mean": np.asarray(mean[i], np.float32).astype(float).tolist(),
             "meanSq": np.asarray(m2[i], np.float32).astype(float).tolist(), "sentence_id": list(sentence_ids[i])}
        return json.dumps(d)[:-1] + ', "children": ['

    out = [json.dumps(params)[:-1], ', "root": ']
    # iterative pre-order with explicit close markers
    stack = [(0, False)]
    first_child = {0: True}
    while stack:
        i, closing = stack.pop()
        if closing:
            out.append("]}")
            continue
        p = int(parent[i])
        if p >= 0:
            if not first_child[p]:
                out.append(", ")
            first_child[p] = False
        out.append(node_head(i))
        first_child[i] = True
        stack.append((i, True))
        for c in reversed(kids[i]):
            stack.append((c, False))
    out.append("}")
    return "".join(out)


def load_tree_json(json_string):
    """Returns (params, parent[n] (BFS index), count[n], mean[n,D], m2[n,D], sentence_ids[n])."""
    data = json.loads(json_string) if isinstance(json_string, str) else json_string
    params = {k: data[k] for k in ("use_info", "acuity_cutoff", "use_kl", "shape", "alpha", "prior_var")}
    nodes, parent = [data["root"]], [-1]
    i = 0
    while i < len(nodes):
        for c in nodes[i].get("children", []):
            nodes.append(c)
            parent.append(i)
        i += 1
    count = np.asarray([nd["count"] for nd in nodes], np.float32)
    mean = np.asarray([nd["mean"] for nd in nodes], np.float32)
    m2 = np.asarray([nd["meanSq"] for nd in nodes], np.float32)
    sids = [list(nd.get("sentence_id") or []) for nd in nodes]
    return params, np.asarray(parent, np.int32), count, mean, m2, sids
