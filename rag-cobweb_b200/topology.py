"""Host-side tree topology processing (pure numpy, runs anywhere).

Inputs are the flat topology arrays of the node store (parent / child lists); outputs are the
orderings and per-sentence root->leaf paths that CobwebWrapper.build_prediction_index
(src/cobweb/CobwebWrapper.py:107-182) derives with a Python BFS over node objects.
"""
import numpy as np


def bfs_order(root, child_off, child_cnt, child_pool):
    """Nodes in BFS order, children in list order (the reference's index numbering,
    CobwebWrapper.py:110-132).  Returns (order[nn] node ids, parent_b[nn] BFS index of the
    parent or -1, depth[nn])."""
    level = np.asarray([root], dtype=np.int64)
    orders, parents, depths = [level], [np.asarray([-1], dtype=np.int64)], [np.zeros(1, np.int64)]
    base, d = 0, 0
    while True:
        cnt = child_cnt[level].astype(np.int64)
        tot = int(cnt.sum())
        if tot == 0:
            break
        # gather the children of every node of this level, in order
        par_local = np.repeat(np.arange(len(level), dtype=np.int64), cnt)
        start = np.repeat(child_off[level].astype(np.int64), cnt)
        within = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        nxt = child_pool[start + within].astype(np.int64)
        d += 1
        orders.append(nxt)
        parents.append(base + par_local)
        depths.append(np.full(tot, d, np.int64))
        base += len(level)
        level = nxt
    return np.concatenate(orders), np.concatenate(parents), np.concatenate(depths)


def restrict_to_paths(order, parent_b, depth, leaves, n_slots):
    """Keep only the index rows that lie on the root->leaf path of one of `leaves` (node ids).
    Returns (order', parent_b', depth') in the same relative (BFS) order."""
    order = np.asarray(order, np.int64)
    row_of = np.full(n_slots, -1, np.int64)
    row_of[order] = np.arange(len(order))
    keep = np.zeros(len(order), bool)
    cur = np.unique(row_of[np.asarray(leaves, np.int64)])
    while len(cur):
        cur = cur[~keep[cur]]
        keep[cur] = True
        cur = np.unique(parent_b[cur])
        cur = cur[cur >= 0]
    new_row = np.cumsum(keep) - 1
    par = parent_b[keep]
    par = np.where(par >= 0, new_row[np.maximum(par, 0)], -1)
    return order[keep], par, depth[keep]


def sentence_paths(order, parent_b, depth, leaf_of_sentence, level_weights=None, n_slots=None):
    """Per-sentence root->leaf paths over index rows.

    leaf_of_sentence[sid] = node id of the leaf holding sentence sid.  Sentences are laid out
    in "positions" sorted by (index row of the leaf, sid) so neighbouring positions share
    ancestors.  Returns dict(pos_sid[L], path_idx[max_len, L] (-1 padded), path_w[max_len, L]
    with level_weights[j] / path_len in fp32, exactly the sparse values of
    CobwebWrapper.py:160-169), max_len)."""
    order = np.asarray(order, np.int64)
    n_slots = int(order.max()) + 1 if n_slots is None else n_slots
    row_of = np.full(n_slots, -1, np.int64)
    row_of[order] = np.arange(len(order))
    leaf_row = row_of[np.asarray(leaf_of_sentence, np.int64)]
    if (leaf_row < 0).any():
        raise ValueError("a sentence points at a node that is not in the tree")
    L = len(leaf_row)
    pos_sid = np.lexsort((np.arange(L), leaf_row)).astype(np.int32)
    lr = leaf_row[pos_sid]
    ldepth = depth[lr]
    max_len = int(ldepth.max()) + 1 if L else 1
    path_idx = np.full((max_len, L), -1, np.int32)
    cur = lr.copy()
    cols = np.arange(L)
    for t in range(max_len):
        j = ldepth - t
        ok = j >= 0
        path_idx[j[ok], cols[ok]] = cur[ok]
        cur = np.where(ok, parent_b[np.maximum(cur, 0)], -1)
    lw = [1.0] * 6 if level_weights is None else list(level_weights)
    wrow = np.ones(max_len, np.float64)
    wrow[: min(len(lw), max_len)] = lw[:max_len]
    plen = (ldepth + 1).astype(np.float64)
    path_w = (wrow[:, None] / plen[None, :]).astype(np.float32)
    path_w[path_idx < 0] = 0.0
    # the same values as a (len, depth) table: the kernel reads path_len per position instead of a
    # weight per (depth, position)
    w_table = np.zeros((max_len + 1, max_len), np.float32)
    for ln in range(1, max_len + 1):
        w_table[ln] = (wrow / float(ln)).astype(np.float32)
    # per-position record for the path kernel: {len, common prefix with the previous position, leaf row, sid}
    plen_i = (ldepth + 1).astype(np.int64)
    pfx = np.zeros(L, np.int64)
    if L > 1:
        same = (path_idx[:, 1:] == path_idx[:, :-1]) & (path_idx[:, 1:] >= 0)   # [max_len, L-1]
        lead = np.cumprod(same, axis=0).sum(axis=0)                               # matching leading levels
        pfx[1:] = np.where(plen_i[1:] == plen_i[:-1], np.minimum(lead, plen_i[1:]), 0)
    pos_rec = np.stack([plen_i, pfx, lr, pos_sid.astype(np.int64)], axis=1).astype(np.int32)
    return dict(pos_sid=pos_sid, path_idx=path_idx, path_w=path_w, path_len=plen_i.astype(np.int32),
                w_table=w_table, level_w=wrow.copy(), pos_rec=np.ascontiguousarray(pos_rec), max_len=max_len)


def generate_weight_schedule(schedule_type, max_depth, **kwargs):
    """CobwebWrapper._generate_weight_schedule (CobwebWrapper.py:368-408)."""
    if schedule_type == "constant":
        return [kwargs.get("value", 1.0)] * max_depth
    if schedule_type == "linear":
        start, end = kwargs.get("start", 1.0), kwargs.get("end", 1.0)
        if kwargs.get("direction", "increase") == "decrease":
            start, end = end, start
        if max_depth == 1:
            return [start]
        step = (end - start) / (max_depth - 1)
        return [start + i * step for i in range(max_depth)]
    if schedule_type == "quadratic":
        start_n = kwargs.get("start_n", 1)
        out = []
        for i in range(max_depth):
            n = start_n + i
            if n == 0:
                n = 1
            out.append(1 / (n ** 2))
        return out
    if schedule_type == "exponential":
        base = kwargs.get("base", 0.5)
        return [base ** i for i in range(max_depth)]
    raise ValueError(f"Unknown schedule type: {schedule_type}")


def fused_layout(order, parent_b, depth, leaf_of_sentence, level_weights=None, n_slots=None, sentence_ids=None,
                 tile=256, sample_every=8):
    """Layout of the fused tensor-core predict (DenseIndex mode "tf32x3f").

    The leaf score sum_j (w_j/len) s_j is rewritten as (C[parent] + w_leaf * s_leaf) / len with the cumulative
    ancestor sums C[n] = C[parent(n)] + w_depth(n) * s_n, which do not depend on the leaf.  Index rows are split
    into internal rows (BFS order, levels contiguous) and sentence-leaf rows; every `sample_every`-th leaf is moved to
    the front (whole tiles of `tile` rows): those rows are scored first and give every query a lower bound of its
    k-th best score, against which the remaining tiles are filtered in the scoring kernel's epilogue.  Returns a dict of numpy arrays (rows are BFS index rows of `order`):
      int_rows, int_parent (internal index or -1), int_w (float32), level_off (internal rows per depth, prefix sums),
      leaf_rows (new leaf order), leaf_parent (internal index or -1), leaf_w, leaf_inv_len (float32), leaf_len, n_sample_tiles,
      sent_off [n_leaf + 1], sent_ids (sentence ids, ascending per leaf),
      flat_pos_rec [n_pos_s, 4], flat_path [n_pos_s, 1] (flat index over the sentences of the sampled leaves)."""
    order = np.asarray(order, np.int64)
    nn = len(order)
    n_slots = int(order.max()) + 1 if n_slots is None else n_slots
    row_of = np.full(n_slots, -1, np.int64)
    row_of[order] = np.arange(nn)
    leaf_of_sentence = np.asarray(leaf_of_sentence, np.int64)
    sids = np.arange(len(leaf_of_sentence), dtype=np.int64) if sentence_ids is None else np.asarray(sentence_ids, np.int64)
    leaf_row_of_sent = row_of[leaf_of_sentence]
    if (leaf_row_of_sent < 0).any():
        raise ValueError("a sentence points at a node that is not in the index")
    is_leaf = np.zeros(nn, bool)
    is_leaf[leaf_row_of_sent] = True
    if is_leaf[np.maximum(parent_b, 0)][parent_b >= 0].any():
        raise ValueError("a node holding sentences has children")
    max_len = int(depth[is_leaf].max()) + 1
    lw = [1.0] * 6 if level_weights is None else list(level_weights)
    wrow = np.ones(max_len, np.float64)
    wrow[: min(len(lw), max_len)] = lw[:max_len]
    # internal rows: BFS order, so parents come first and every depth is one contiguous range
    int_rows = np.nonzero(~is_leaf)[0]
    int_of_row = np.full(nn, -1, np.int64)
    int_of_row[int_rows] = np.arange(len(int_rows))
    int_parent = np.where(parent_b[int_rows] >= 0, int_of_row[np.maximum(parent_b[int_rows], 0)], -1)
    int_depth = depth[int_rows]
    int_w = wrow[np.minimum(int_depth, max_len - 1)].astype(np.float32)
    n_levels = int(int_depth.max()) + 1 if len(int_rows) else 0
    level_off = np.concatenate([[0], np.cumsum(np.bincount(int_depth, minlength=n_levels))]).astype(np.int64)
    # leaf rows: every `sample_every`-th leaf (in BFS order) forms the sample -- a strided sample, because BFS order
    # keeps the leaves of a cluster together and whole tiles of neighbours would miss most clusters -- cut to whole
    # tiles and moved to the front; the other leaves keep their BFS order (siblings adjacent: their parents' sums are
    # read together), the partial last tile is theirs
    leaves = np.nonzero(is_leaf)[0]
    n_leaf = len(leaves)
    if n_leaf < sample_every * tile and n_leaf >= 2 * tile:
        sample_every = n_leaf // tile  # small index: one sampled tile with a coarser stride
    samp = np.arange(0, n_leaf, sample_every, dtype=np.int64)
    n_s_tiles = len(samp) // tile if n_leaf >= sample_every * tile else 0
    samp = samp[: n_s_tiles * tile]
    rest_mask = np.ones(n_leaf, bool)
    rest_mask[samp] = False
    idx = np.concatenate([samp, np.nonzero(rest_mask)[0]]).astype(np.int64)
    sampled = np.arange(n_s_tiles)
    leaf_rows = leaves[idx]
    leaf_len = depth[leaf_rows] + 1
    leaf_parent = np.where(parent_b[leaf_rows] >= 0, int_of_row[np.maximum(parent_b[leaf_rows], 0)], -1)
    leaf_w = wrow[leaf_len - 1].astype(np.float32)
    leaf_inv_len = (1.0 / leaf_len).astype(np.float32)
    # sentences per leaf (new leaf order), ascending ids
    new_of_row = np.full(nn, -1, np.int64)
    new_of_row[leaf_rows] = np.arange(n_leaf)
    sent_leaf = new_of_row[leaf_row_of_sent]
    so = np.lexsort((sids, sent_leaf))
    sent_ids = sids[so]
    sent_off = np.concatenate([[0], np.cumsum(np.bincount(sent_leaf, minlength=n_leaf))]).astype(np.int64)
    # flat index over the sentences of the sampled leaves
    n_s_rows = len(sampled) * tile
    n_pos_s = int(sent_off[n_s_rows])
    flat_leaf = sent_leaf[so][:n_pos_s]
    same = np.zeros(n_pos_s, np.int64)
    if n_pos_s > 1:
        same[1:] = flat_leaf[1:] == flat_leaf[:-1]
    flat_pos_rec = np.stack([np.ones(n_pos_s, np.int64), same, flat_leaf, sent_ids[:n_pos_s]], axis=1).astype(np.int32)
    return dict(int_rows=int_rows, int_parent=int_parent.astype(np.int32), int_w=int_w, level_off=level_off,
                leaf_rows=leaf_rows, leaf_parent=leaf_parent.astype(np.int32), leaf_w=leaf_w, leaf_inv_len=leaf_inv_len,
                leaf_len=leaf_len.astype(np.int32),
                n_sample_tiles=int(len(sampled)), sent_off=sent_off.astype(np.int32), sent_ids=sent_ids.astype(np.int32),
                flat_pos_rec=np.ascontiguousarray(flat_pos_rec), flat_path=flat_leaf.astype(np.int32).reshape(-1, 1),
                max_len=max_len)
