"""Tree topology processing ON THE DEVICE: the same orderings, per-sentence paths and fused layout
as topology.py derives with numpy on the host, computed with torch ops on the node store's own tensors, so that a
prediction index is rebuilt after add_sentences without copying the topology to the host (CobwebWrapper.py:80 only
invalidates; :91-208 rebuilds with a Python BFS).

Every function takes torch tensors on any device and mirrors the numpy function of the same name in topology.py, which
stays the CPU-testable statement of the result (tests/test_host_logic.py compares the two on random trees).
"""
import torch


def bfs_order(root, child_off, child_cnt, child_pool):
    """Nodes in BFS order, children in list order (the reference's index numbering, CobwebWrapper.py:110-132).
    Returns (order[nn] node ids, parent_b[nn] BFS index of the parent or -1, depth[nn]) as int64 tensors.
    One round of gathers per tree level."""
    dev = child_off.device
    level = torch.tensor([int(root)], dtype=torch.int64, device=dev)
    orders, parents, depths = [level], [torch.full((1,), -1, dtype=torch.int64, device=dev)], [torch.zeros(1, dtype=torch.int64, device=dev)]
    base, d = 0, 0
    while True:
        cnt = child_cnt[level].to(torch.int64)
        tot = int(cnt.sum())   # one scalar read-back per level: sizes the next level
        if tot == 0:
            break
        par_local = torch.repeat_interleave(torch.arange(level.numel(), device=dev), cnt, output_size=tot)
        start = child_off[level].to(torch.int64)[par_local]
        first = (torch.cumsum(cnt, 0) - cnt)[par_local]
        within = torch.arange(tot, device=dev) - first
        nxt = child_pool[start + within].to(torch.int64)
        d += 1
        orders.append(nxt)
        parents.append(base + par_local)
        depths.append(torch.full((tot,), d, dtype=torch.int64, device=dev))
        base += level.numel()
        level = nxt
    return torch.cat(orders), torch.cat(parents), torch.cat(depths)


def sentence_paths(order, parent_b, depth, leaf_of_sentence, level_weights=None, n_slots=None):
    """Per-sentence root->leaf paths over index rows (topology.sentence_paths).  Returns dict(path_idx [L, max_len]
    int32 (-1 padded; NOTE: position-major, the layout the kernels read), level_w (python list of float64),
    pos_rec [L, 4] int32, pos_leaf_row [L] int64 ascending, path_lens (sorted unique lengths, python list), max_len)."""
    dev = order.device
    nn = order.numel()
    n_slots = int(order.max()) + 1 if n_slots is None else n_slots
    row_of = torch.full((n_slots,), -1, dtype=torch.int64, device=dev)
    row_of[order] = torch.arange(nn, device=dev)
    leaf_row = row_of[leaf_of_sentence.to(torch.int64)]
    if bool((leaf_row < 0).any()):
        raise ValueError("a sentence points at a node that is not in the tree")
    L = leaf_row.numel()
    lr, pos_sid = torch.sort(leaf_row, stable=True)   # positions sorted by (index row of the leaf, sid)
    ldepth = depth[lr]
    max_len = int(ldepth.max()) + 1 if L else 1
    path_idx = torch.full((L, max_len), -1, dtype=torch.int64, device=dev)
    cur = lr.clone()
    cols = torch.arange(L, device=dev)
    for t in range(max_len):
        j = ldepth - t
        ok = j >= 0
        path_idx[cols[ok], j[ok]] = cur[ok]
        cur = torch.where(ok, parent_b[cur.clamp_min(0)], torch.full_like(cur, -1))
    lw = [1.0] * 6 if level_weights is None else list(level_weights)
    wrow = [1.0] * max_len
    for j in range(min(len(lw), max_len)):
        wrow[j] = float(lw[j])
    plen = ldepth + 1
    pfx = torch.zeros(L, dtype=torch.int64, device=dev)
    if L > 1:
        same = (path_idx[1:] == path_idx[:-1]) & (path_idx[1:] >= 0)        # [L-1, max_len]
        lead = torch.cumprod(same.to(torch.int64), dim=1).sum(dim=1)        # matching leading levels
        pfx[1:] = torch.where(plen[1:] == plen[:-1], torch.minimum(lead, plen[1:]), torch.zeros_like(lead))
    pos_rec = torch.stack([plen, pfx, lr, pos_sid], dim=1).to(torch.int32).contiguous()
    return dict(path_idx=path_idx.to(torch.int32).contiguous(), level_w=wrow, pos_rec=pos_rec, pos_leaf_row=lr,
                path_lens=torch.unique(plen).tolist() if L else [], max_len=max_len)


def fused_layout(order, parent_b, depth, leaf_of_sentence, level_weights=None, n_slots=None, sentence_ids=None, tile=256,
                 sample_every=8):
    """Layout of the fused tensor-core predict (topology.fused_layout) from device tensors.  Returns a dict of device
    tensors (int32 / float32) plus python ints: int_rows, int_parent, int_w, level_off, leaf_rows, leaf_parent, leaf_w,
    leaf_inv_len, leaf_len, n_sample_tiles, sent_off, sent_ids, max_len."""
    dev = order.device
    nn = order.numel()
    n_slots = int(order.max()) + 1 if n_slots is None else n_slots
    row_of = torch.full((n_slots,), -1, dtype=torch.int64, device=dev)
    row_of[order] = torch.arange(nn, device=dev)
    leaf_of_sentence = leaf_of_sentence.to(torch.int64)
    sids = torch.arange(leaf_of_sentence.numel(), device=dev) if sentence_ids is None else sentence_ids.to(torch.int64)
    leaf_row_of_sent = row_of[leaf_of_sentence]
    if bool((leaf_row_of_sent < 0).any()):
        raise ValueError("a sentence points at a node that is not in the index")
    is_leaf = torch.zeros(nn, dtype=torch.bool, device=dev)
    is_leaf[leaf_row_of_sent] = True
    has_par = parent_b >= 0
    if bool(is_leaf[parent_b[has_par]].any()):
        raise ValueError("a node holding sentences has children")
    max_len = int(depth[is_leaf].max()) + 1
    lw = [1.0] * 6 if level_weights is None else list(level_weights)
    wrow = torch.ones(max_len, dtype=torch.float64, device=dev)
    m = min(len(lw), max_len)
    if m:
        wrow[:m] = torch.tensor(lw[:m], dtype=torch.float64, device=dev)
    int_rows = torch.nonzero(~is_leaf).view(-1)
    int_of_row = torch.full((nn,), -1, dtype=torch.int64, device=dev)
    int_of_row[int_rows] = torch.arange(int_rows.numel(), device=dev)
    ip = parent_b[int_rows]
    int_parent = torch.where(ip >= 0, int_of_row[ip.clamp_min(0)], torch.full_like(ip, -1))
    int_depth = depth[int_rows]
    int_w = wrow[int_depth.clamp_max(max_len - 1)].to(torch.float32)
    n_levels = int(int_depth.max()) + 1 if int_rows.numel() else 0
    level_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev),
                           torch.cumsum(torch.bincount(int_depth, minlength=n_levels), 0)])
    leaves = torch.nonzero(is_leaf).view(-1)
    n_leaf = leaves.numel()
    if n_leaf < sample_every * tile and n_leaf >= 2 * tile:
        sample_every = n_leaf // tile
    samp = torch.arange(0, n_leaf, sample_every, device=dev)
    n_s_tiles = samp.numel() // tile if n_leaf >= sample_every * tile else 0
    samp = samp[: n_s_tiles * tile]
    rest = torch.ones(n_leaf, dtype=torch.bool, device=dev)
    rest[samp] = False
    idx = torch.cat([samp, torch.nonzero(rest).view(-1)])
    leaf_rows = leaves[idx]
    leaf_len = depth[leaf_rows] + 1
    lp = parent_b[leaf_rows]
    leaf_parent = torch.where(lp >= 0, int_of_row[lp.clamp_min(0)], torch.full_like(lp, -1))
    leaf_w = wrow[leaf_len - 1].to(torch.float32)
    leaf_inv_len = (1.0 / leaf_len.to(torch.float64)).to(torch.float32)
    new_of_row = torch.full((nn,), -1, dtype=torch.int64, device=dev)
    new_of_row[leaf_rows] = torch.arange(n_leaf, device=dev)
    sent_leaf = new_of_row[leaf_row_of_sent]
    # sentences per leaf (new leaf order), ascending ids: sort by id first, then stably by leaf
    o1 = torch.argsort(sids, stable=True)
    o2 = torch.argsort(sent_leaf[o1], stable=True)
    so = o1[o2]
    sent_ids = sids[so]
    sent_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(torch.bincount(sent_leaf, minlength=n_leaf), 0)])
    i32 = lambda t: t.to(torch.int32).contiguous()
    return dict(int_rows=i32(int_rows), int_parent=i32(int_parent), int_w=int_w.contiguous(), level_off=i32(level_off),
                leaf_rows=i32(leaf_rows), leaf_parent=i32(leaf_parent), leaf_w=leaf_w.contiguous(),
                leaf_inv_len=leaf_inv_len.contiguous(), leaf_len=i32(leaf_len), n_sample_tiles=int(n_s_tiles),
                sent_off=i32(sent_off), sent_ids=i32(sent_ids), max_len=max_len)
