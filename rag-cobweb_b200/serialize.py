"""JSON wire format of the reference (pure Python / numpy, runs anywhere).

Tree document (CobwebTorchTree.dump_json, src/cobweb/CobwebTorchTree.py:67-81 with
CobwebTorchNode.iterative_output_json, src/cobweb/CobwebTorchNode.py:741-772):
  {"use_info":..,"acuity_cutoff":..,"use_kl":..,"shape":[D],"alpha":..,"prior_var":..,
   "root": {"count": c, "mean": [...], "meanSq": [...], "sentence_id": [...], "children": [ ... ]}}
Documents written here load in the reference's CobwebTorchTree.load_json and vice versa
(the reference's loader reverses child order; ours keeps it).
"""
import json
import struct

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


# ------------------------------------------------------------------ binary snapshot (additive)
# The JSON document costs ~20 bytes of decimal text per float and a Python object per node: a 1M-node tree does not
# leave the device that way.  The snapshot is the same content -- parameters, BFS topology, node statistics, sentence
# ids per node, the wrapper's sentence -> node map -- as raw little-endian arrays:
#   magic "CWB200S1" | header json length (u64) | header json | parent i32[n] | count f32[n] | n_sent i32[n] |
#   mean f32[n, D] | m2 f32[n, D] | leaf_of_sentence i32[L]   (nodes in BFS order, parent[i] < i, node i = row i)
SNAP_MAGIC = b"CWB200S1"
SNAP_CHUNK_ROWS = 1 << 16


def write_snapshot(path, params, parent, count, n_sent, rows_fn, leaf_of_sentence=None, extra=None):
    """rows_fn(lo, hi) -> (mean[lo:hi], m2[lo:hi]) float32 arrays of the BFS rows: called in chunks, so the node matrices
    are streamed from the device without a second full copy on the host."""
    n = len(parent)
    leaf = np.zeros(0, np.int32) if leaf_of_sentence is None else np.ascontiguousarray(leaf_of_sentence, np.int32)
    head = dict(params=params, n_nodes=int(n), n_sentences=int(len(leaf)), extra=extra or {})
    hb = json.dumps(head).encode()
    with open(path, "wb") as f:
        f.write(SNAP_MAGIC)
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        np.ascontiguousarray(parent, "<i4").tofile(f)
        np.ascontiguousarray(count, "<f4").tofile(f)
        np.ascontiguousarray(n_sent, "<i4").tofile(f)
        for which in (0, 1):
            for lo in range(0, n, SNAP_CHUNK_ROWS):
                np.ascontiguousarray(rows_fn(lo, min(n, lo + SNAP_CHUNK_ROWS))[which], "<f4").tofile(f)
        leaf.astype("<i4").tofile(f)


def read_snapshot(path):
    """Returns dict(params, parent, count, n_sent, mean, m2, leaf_of_sentence, extra); mean / m2 are memory-mapped."""
    with open(path, "rb") as f:
        if f.read(len(SNAP_MAGIC)) != SNAP_MAGIC:
            raise ValueError(f"{path}: not a cobweb-b200 snapshot")
        (hl,) = struct.unpack("<Q", f.read(8))
        head = json.loads(f.read(hl).decode())
        off = len(SNAP_MAGIC) + 8 + hl
    n, L, d = head["n_nodes"], head["n_sentences"], int(head["params"]["shape"][0])
    mm = lambda dtype, shape, o: np.memmap(path, dtype=dtype, mode="r", offset=o, shape=shape)
    parent = np.array(mm("<i4", (n,), off)); off += 4 * n
    count = np.array(mm("<f4", (n,), off)); off += 4 * n
    n_sent = np.array(mm("<i4", (n,), off)); off += 4 * n
    mean = mm("<f4", (n, d), off); off += 4 * n * d
    m2 = mm("<f4", (n, d), off); off += 4 * n * d
    leaf = np.array(mm("<i4", (L,), off)) if L else np.zeros(0, np.int32)
    return dict(params=head["params"], parent=parent, count=count, n_sent=n_sent, mean=mean, m2=m2,
                leaf_of_sentence=leaf, extra=head.get("extra", {}))
