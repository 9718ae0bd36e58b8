"""Multi-GPU: reads shard across ranks (contiguous ranges of the global submit order); there is no data-path collective
while reads are walked.  At end of chromosome the per-shard tables merge by key.

Device path (exchange_keys / reduce_aggregates, NCCL): every rank exports the KEYS of its cluster, sub-cluster and
break-point-pair tables with their first-appearance indices as one packed device buffer (npt_export_keys); one all-gather;
every rank inserts the other ranks' keys into its own tables (npt_import_keys).  All ranks then hold the same key set with
the same minima, so npt_end_chrom numbers clusters (CigarClusters.java:83) and sub-clusters (CigarCluster.java:660)
identically everywhere, per-read ids come out global, and the counters — coverage rows, ORF counters, sub-cluster counts,
break sums, pair counts, now in the same order on every rank — are summed in place with all-reduces.

Host path (merge_tables, also through libnpt's npt_merge_tables): the same merge on plain result dicts with vectorised
numpy; the CPU tests drive it with oracle output under gloo, and it is the independent check of the device path.
"""
import time

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


def merge_tables_native(tables):
    """Same contract as merge_tables, through libnpt's npt_merge_tables (C++ hash maps; ~20x faster than the numpy version,
    which stays as the independent check in the tests)."""
    import ctypes as C
    from . import abi, engine
    L = engine.lib()
    S = len(tables[0]["total_counts"])
    keep, structs = [], []
    spec = dict(cl_first_read=np.int64, cl_key_off=np.uint32, cl_key_tok=np.int32, cl_sub_off=np.uint32, sub_first_read=np.int64,
                sub_key_off=np.uint32, sub_key=np.int32, sub_count=np.int32, sub_break_sum=np.int32, sub_sum_off=np.uint32, sub_sums=np.int64,
                sub_flags=np.uint8, total_counts=np.int32, spliced=np.int32, unspliced=np.int32, pair_src=np.int32, pair_level=np.int32,
                pair_start=np.int32, pair_end=np.int32, pair_count=np.int32)
    ct = {np.int64: C.c_int64, np.uint32: C.c_uint32, np.int32: C.c_int32, np.uint8: C.c_uint8}
    for t in tables:
        r = abi.npt_results()
        r.n_clusters, r.n_subs, r.n_pairs = len(t["cl_first_read"]), len(t["sub_first_read"]), len(t["pair_src"])
        r.n_orf = t["spliced"].shape[-1] if t["spliced"].size else 0
        for k, dt in spec.items():
            a = np.ascontiguousarray(t[k], dtype=dt)
            keep.append(a)
            setattr(r, k, a.ctypes.data_as(C.POINTER(ct[dt])))
        structs.append(r)
    arr = (C.POINTER(abi.npt_results) * len(structs))(*[C.pointer(x) for x in structs])
    h = L.npt_merge_tables(len(structs), arr, S)
    try:
        m = L.npt_merged_results(h).contents
        out = engine.Results.from_struct(m, S, abi.NPT_FETCH_TABLES)
        maps = []
        for r, t in enumerate(tables):
            nc, ns = len(t["cl_first_read"]), len(t["sub_first_read"])
            cm = np.ctypeslib.as_array(L.npt_merged_cluster_map(h, r), shape=(nc,)).copy() if nc else np.zeros(0, np.int32)
            sm = np.ctypeslib.as_array(L.npt_merged_sub_map(h, r), shape=(ns,)).copy() if ns else np.zeros(0, np.int32)
            maps.append((cm, sm))
    finally:
        L.npt_merge_free(h)
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


class _DevPtr:
    """Expose a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def exchange_keys(eng, rank, world):
    """Step 1 of the device-side protocol (after the last submit, before end_chrom): every rank's table keys reach every
    other rank over NCCL and are inserted into its tables.  Returns ms."""
    import torch
    import torch.distributed as dist
    t0 = time.perf_counter()
    eng.wait()  # this rank's kernels
    t1 = time.perf_counter()
    ptr, n = eng.export_keys()
    t2 = time.perf_counter()
    sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
    sizes[rank] = n
    dist.all_reduce(sizes)
    sizes = sizes.cpu().tolist()
    t3 = time.perf_counter()  # includes waiting for the slowest rank
    m = int(max(sizes))
    mine = torch.zeros(m, dtype=torch.int32, device="cuda")
    mine[:n] = torch.as_tensor(_DevPtr(ptr, n, "<i4"), device="cuda")
    allbuf = torch.empty(world * m, dtype=torch.int32, device="cuda")
    dist.all_gather_into_tensor(allbuf, mine)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    for r in range(world):
        if r != rank:
            eng.import_keys(allbuf[r * m:].data_ptr(), int(sizes[r]))
    t5 = time.perf_counter()
    exchange_keys.last = dict(wait_own_kernels_ms=1e3 * (t1 - t0), export_ms=1e3 * (t2 - t1), sizes_allreduce_ms=1e3 * (t3 - t2),
                              all_gather_ms=1e3 * (t4 - t3), import_ms=1e3 * (t5 - t4), words=int(n))
    return 1e3 * (t5 - t0)


def reduce_aggregates(eng):
    """Step 2 (after end_chrom): sum the counters, which are now in the same global order on every rank.  Returns ms."""
    import torch
    import torch.distributed as dist
    t0 = time.perf_counter()
    v = eng.device_view()
    for ptr, n, dt in ((v.coverage, v.coverage_elems, "<i4"), (v.small_counts, v.small_elems, "<i4"), (v.sub_count, v.sub_count_elems, "<i4"),
                       (v.sub_break_sum, v.sub_break_sum_elems, "<i4"), (v.sub_sums, v.sub_sums_elems, "<i8"), (v.pair_count, v.pair_count_elems, "<i4")):
        if ptr and n:
            dist.all_reduce(torch.as_tensor(_DevPtr(ptr, n, dt), device="cuda"))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0)


def merge_across_ranks_host(eng, rank, world, timing=False):
    """Older GPU path, kept for comparison (call after end_chrom): tables fetched and merged on the host, coverage rows permuted on
    the device and all-reduced.  0.18 s per step at 10 M reads per rank (sorting and hashing ~0.5 M break-point pairs on the host).
    Returns the merge time in ms (timing=True) or (global tables, (cluster_map, sub_map), local cl_sub_off)."""
    import torch
    import torch.distributed as dist
    from . import abi
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.fetch(abi.NPT_FETCH_TABLES)
    t1 = time.perf_counter()
    mine = local_tables(res, 0)  # first-read indices are already global (npt_set_read_index_base)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t2 = time.perf_counter()
    glob, maps = merge_tables_native(gathered)
    t3 = time.perf_counter()
    cmap, smap = maps[rank]
    eng.remap_clusters(cmap, glob["n_clusters"])
    v = eng.device_view()
    if v.coverage and v.coverage_elems:
        cov = torch.as_tensor(_DevPtr(v.coverage, v.coverage_elems), device="cuda")
        dist.all_reduce(cov)
    small = torch.as_tensor(_DevPtr(v.small_counts, v.small_elems), device="cuda")
    dist.all_reduce(small)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    ms = 1e3 * (t4 - t0)
    merge_across_ranks_host.last = dict(fetch_ms=1e3 * (t1 - t0), gather_ms=1e3 * (t2 - t1), merge_ms=1e3 * (t3 - t2), device_ms=1e3 * (t4 - t3))
    return ms if timing else (glob, (cmap, smap), res["cl_sub_off"])
