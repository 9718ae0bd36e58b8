"""Host-side model of the multi-rank table merge, in numpy.

The product path merges on the devices inside libnpt (csrc/comm.inl: npt_comm_init + npt_end_chrom_all, or npt_mg_*): keys
are exchanged over NCCL, inserted into every rank's tables, numbered identically everywhere and the counters all-reduced.
This module restates the same merge on plain result dicts — unique keys, id = rank of the minimum first-read index, summed
counters, local->global id maps — so that the N>1 logic is covered on CPU (tests/test_multirank_gloo.py: world_size 2,
gloo, oracle output per rank) and serves as an independent check of the device protocol.
"""
import numpy as np

_SENT = np.int64(-(1 << 62))


def _padded(flat, off, width):
    """Ragged rows (flat, offsets) → dense [n, width] int64 matrix padded with a sentinel."""
    off = off.astype(np.int64)
    n = len(off) - 1
    ln = off[1:] - off[:-1]
    out = np.full((n, width), _SENT, np.int64)
    if n and width and len(flat):
        j = np.arange(width)[None, :]
        m = j < ln[:, None]
        idx = np.minimum(off[:-1, None] + j, len(flat) - 1)
        out[m] = flat.astype(np.int64)[idx][m]
    return out, ln


def _unique_rows(K):
    """np.unique(K, axis=0, return_inverse=True) for int64 matrices, an order of magnitude faster: unique on a 64-bit row
    hash, then an exact verification that every row equals its group's representative (falls back if a hash collides)."""
    if len(K) == 0:
        return K, np.zeros(0, np.int64)
    h = np.full(len(K), 0x9E3779B97F4A7C15, np.uint64)
    with np.errstate(over="ignore"):
        for j in range(K.shape[1]):
            h = (h ^ K[:, j].astype(np.uint64)) * np.uint64(0x100000001B3)
            h ^= h >> np.uint64(29)
    _, rep, inv = np.unique(h, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    if not np.array_equal(K, K[rep[inv]]):
        uq, inv = np.unique(K, axis=0, return_inverse=True)
        return uq, inv.reshape(-1)
    return K[rep], inv


def local_tables(res, idx_base=0):
    """The part of a rank's results that takes part in the merge (small: O(clusters + sub-clusters)), as flat arrays."""
    keys = ("cl_key_off", "cl_key_tok", "cl_first_read", "cl_sub_off", "sub_key_off", "sub_key", "sub_first_read", "sub_count", "sub_break_sum",
            "sub_sum_off", "sub_sums", "sub_flags", "pair_src", "pair_level", "pair_start", "pair_end", "pair_count", "total_counts", "spliced",
            "unspliced")
    t = {k: np.asarray(res[k]) for k in keys}
    t["cl_first_read"] = t["cl_first_read"].astype(np.int64) + idx_base
    t["sub_first_read"] = t["sub_first_read"].astype(np.int64) + idx_base
    return t


def merge_tables(tables):
    """tables: list over ranks of local_tables().  Returns (global tables dict, per-rank maps).
    maps[r] = (cluster_map: int32[local clusters] → global id, sub_map: int32[local subs, cluster-major] → global k inside its cluster)."""
    R = len(tables)
    S = len(tables[0]["total_counts"])
    # ---------------- clusters: unique token lists; id = rank of the minimum first read index
    Lc = max([int(np.diff(t["cl_key_off"].astype(np.int64)).max()) if len(t["cl_key_off"]) > 1 else 0 for t in tables] + [0])
    rows, firsts, owner = [], [], []
    for r, t in enumerate(tables):
        M, ln = _padded(t["cl_key_tok"], t["cl_key_off"], Lc)
        rows.append(np.concatenate([ln[:, None], M], axis=1))
        firsts.append(t["cl_first_read"])
        owner.append(np.full(len(ln), r))
    K = np.concatenate(rows) if rows else np.zeros((0, Lc + 1), np.int64)
    first = np.concatenate(firsts) if firsts else np.zeros(0, np.int64)
    uq, inv = _unique_rows(K)
    nU = len(uq)
    ufirst = np.full(nU, np.iinfo(np.int64).max)
    np.minimum.at(ufirst, inv, first)
    order = np.argsort(ufirst, kind="stable")
    gid_of_u = np.empty(nU, np.int64)
    gid_of_u[order] = np.arange(nU)
    cl_gid = gid_of_u[inv]                       # per (rank, local cluster) → global id
    splits = np.cumsum([len(x) for x in firsts])[:-1]
    cl_gid_by_rank = np.split(cl_gid, splits) if R else []
    uq_sorted = uq[order]
    cl_len = uq_sorted[:, 0] if nU else np.zeros(0, np.int64)
    toks = uq_sorted[:, 1:]
    mask = np.arange(Lc)[None, :] < cl_len[:, None] if nU else np.zeros((0, Lc), bool)
    out = dict(n_clusters=nU, cl_first_read=ufirst[order].astype(np.int64), cl_key_tok=toks[mask].astype(np.int32),
               cl_key_off=np.concatenate([[0], np.cumsum(cl_len)]).astype(np.uint32))
    # ---------------- sub-clusters: unique (global cluster id, binned break list)
    Ls = max([int(np.diff(t["sub_key_off"].astype(np.int64)).max()) if len(t["sub_key_off"]) > 1 else 0 for t in tables] + [0])
    Lm = max([int(np.diff(t["sub_sum_off"].astype(np.int64)).max()) if len(t["sub_sum_off"]) > 1 else 0 for t in tables] + [0])
    srows, sfirst, scount, sbs, ssum, sflag, sn = [], [], [], [], [], [], []
    for r, t in enumerate(tables):
        ns = len(t["sub_first_read"])
        M, ln = _padded(t["sub_key"], t["sub_key_off"], Ls)
        per_cl = np.diff(t["cl_sub_off"].astype(np.int64))
        gcl = np.repeat(cl_gid_by_rank[r], per_cl) if ns else np.zeros(0, np.int64)
        srows.append(np.concatenate([gcl[:, None], ln[:, None], M], axis=1))
        sfirst.append(t["sub_first_read"]); scount.append(t["sub_count"].reshape(ns, S).astype(np.int64)); sbs.append(t["sub_break_sum"].astype(np.int64))
        Ms, lns = _padded(t["sub_sums"], t["sub_sum_off"], Lm)
        Ms[Ms == _SENT] = 0
        ssum.append(Ms); sn.append(lns); sflag.append(t["sub_flags"].astype(np.int64))
    SK = np.concatenate(srows) if srows else np.zeros((0, Ls + 2), np.int64)
    sf = np.concatenate(sfirst); sc = np.concatenate(scount); sb = np.concatenate(sbs); sm = np.concatenate(ssum); sl = np.concatenate(sn); sg = np.concatenate(sflag)
    suq, sinv = _unique_rows(SK)
    nS = len(suq)
    usf = np.full(nS, np.iinfo(np.int64).max)
    np.minimum.at(usf, sinv, sf)
    # flags of the read that came first: the gathered row whose first index equals the minimum
    is_min = sf == usf[sinv]
    uflag = np.zeros(nS, np.int64)
    uflag[sinv[is_min]] = sg[is_min]
    ucount = np.zeros((nS, S), np.int64); np.add.at(ucount, sinv, sc)
    ubs = np.zeros(nS, np.int64); np.add.at(ubs, sinv, sb)
    usum = np.zeros((nS, Lm), np.int64); np.add.at(usum, sinv, sm)
    ulen = np.zeros(nS, np.int64); np.maximum.at(ulen, sinv, sl)
    sorder = np.lexsort((usf, suq[:, 0])) if nS else np.zeros(0, np.int64)   # by cluster, then first index
    s_cl = suq[sorder, 0] if nS else np.zeros(0, np.int64)
    cl_sub_off = np.concatenate([[0], np.cumsum(np.bincount(s_cl, minlength=nU))]) if nU else np.zeros(1, np.int64)
    k_in_cl = np.arange(nS) - cl_sub_off[s_cl] if nS else np.zeros(0, np.int64)
    k_of_u = np.empty(nS, np.int64)
    k_of_u[sorder] = k_in_cl
    sub_k = k_of_u[sinv]                          # per (rank, local sub) → global k inside its cluster
    ssplits = np.cumsum([len(x) for x in sfirst])[:-1]
    sub_k_by_rank = np.split(sub_k, ssplits) if R else []
    klen = suq[sorder, 1] if nS else np.zeros(0, np.int64)
    kmask = np.arange(Ls)[None, :] < klen[:, None] if nS else np.zeros((0, Ls), bool)
    smask = np.arange(Lm)[None, :] < ulen[sorder][:, None] if nS else np.zeros((0, Lm), bool)
    counts = ucount[sorder]
    out.update(n_subs=nS, cl_sub_off=cl_sub_off.astype(np.uint32), sub_first_read=usf[sorder], sub_key=suq[sorder, 2:][kmask].astype(np.int32),
               sub_key_off=np.concatenate([[0], np.cumsum(klen)]).astype(np.uint32), sub_count=counts.astype(np.int32),
               sub_break_sum=ubs[sorder].astype(np.int32), sub_sums=usum[sorder][smask].astype(np.int64),
               sub_sum_off=np.concatenate([[0], np.cumsum(ulen[sorder])]).astype(np.uint32), sub_flags=uflag[sorder].astype(np.uint8))
    rc = np.zeros((nU, S), np.int64)
    if nS:
        np.add.at(rc, s_cl, counts)
    out["cl_read_count"] = rc.astype(np.int32)
    # ---------------- break-point pairs, counters
    P = np.concatenate([np.stack([t["pair_src"], t["pair_level"], t["pair_start"], t["pair_end"]], axis=1).astype(np.int64) for t in tables])
    pc = np.concatenate([t["pair_count"].astype(np.int64) for t in tables])
    if len(P):
        pu, pinv = _unique_rows(P)
        srt = np.lexsort((pu[:, 3], pu[:, 2], pu[:, 1], pu[:, 0]))
        rank_of = np.empty(len(pu), np.int64); rank_of[srt] = np.arange(len(pu))
        pu, pinv = pu[srt], rank_of[pinv]
        cnt = np.zeros(len(pu), np.int64); np.add.at(cnt, pinv, pc)
    else:
        pu, cnt = np.zeros((0, 4), np.int64), np.zeros(0, np.int64)
    for i, name in enumerate(("pair_src", "pair_level", "pair_start", "pair_end")):
        out[name] = pu[:, i].astype(np.int32)
    out["pair_count"] = cnt.astype(np.int32)
    out["n_pairs"] = len(pu)
    for k in ("total_counts", "spliced", "unspliced"):
        out[k] = sum(t[k].astype(np.int64) for t in tables).astype(np.int32)
    maps = [(cl_gid_by_rank[r].astype(np.int32), sub_k_by_rank[r].astype(np.int32)) for r in range(R)]
    return out, maps


def remap_reads(res, cmap, smap, cl_sub_off=None):
    """Rewrite a rank's per-read local ids to global ids.  smap is cluster-major over the rank's local sub-clusters."""
    cl, sb = res["cluster_id"].copy(), res["sub_id"].copy()
    ok = cl >= 0
    if len(cmap):
        off = (res["cl_sub_off"] if cl_sub_off is None else cl_sub_off).astype(np.int64)
        sb[ok] = smap[off[cl[ok]] + sb[ok]]
        cl[ok] = cmap[cl[ok]]
    return cl, sb


def permute_rows(rows, cmap, n_global):
    """Host version of npt_remap_clusters: rows [n_local, ...] → [n_global, ...] with zeros for absent clusters."""
    out = np.zeros((n_global,) + rows.shape[1:], rows.dtype)
    if len(cmap):
        out[cmap] = rows
    return out
