"""Multi-GPU: reads shard across ranks (contiguous ranges of the global submit order); there is no data-path collective
while reads are walked.  At end of chromosome the per-shard tables merge by key:

  1. all-gather of the small per-rank tables {cluster key tokens, first global read index, counts} and
     {cluster key, binned breaks, first index, counts, break sums}                       (torch.distributed, object gather)
  2. every rank merges them deterministically: global cluster id = rank of min first index over ranks
     (CigarClusters.java:83), global sub id = rank inside the cluster (CigarCluster.java:660), counts and sums add
  3. each rank permutes its dense coverage rows into global cluster order on the device (npt_remap_clusters) and the
     rows are summed with one all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests); ORF counters likewise
  4. per-read ids never leave their rank: local ids are rewritten through the local→global maps.

The merge logic works on plain result dicts (numpy), so the CPU tests drive it with oracle output under gloo.
"""
import time

import numpy as np


def local_tables(res, idx_base=0):
    """The part of a rank's results that takes part in the merge (small: O(clusters + sub-clusters))."""
    nc, ns = res["n_clusters"], res["n_subs"]
    ko, so, sko, sso = res["cl_key_off"], res["cl_sub_off"], res["sub_key_off"], res["sub_sum_off"]
    clusters = []
    for c in range(nc):
        key = tuple(int(t) for t in res["cl_key_tok"][int(ko[c]):int(ko[c + 1])])
        subs = []
        for g in range(int(so[c]), int(so[c + 1])):
            subs.append(dict(key=tuple(int(t) for t in res["sub_key"][int(sko[g]):int(sko[g + 1])]),
                             first=int(res["sub_first_read"][g]) + idx_base, count=res["sub_count"][g].astype(np.int64),
                             break_sum=int(res["sub_break_sum"][g]), sums=res["sub_sums"][int(sso[g]):int(sso[g + 1])].astype(np.int64),
                             flags=int(res["sub_flags"][g])))
        clusters.append(dict(key=key, first=int(res["cl_first_read"][c]) + idx_base, subs=subs))
    pairs = {(int(a), int(b), int(c_), int(d)): int(e) for a, b, c_, d, e in
             zip(res["pair_src"], res["pair_level"], res["pair_start"], res["pair_end"], res["pair_count"])}
    return dict(clusters=clusters, pairs=pairs, total_counts=res["total_counts"].astype(np.int64),
                spliced=res["spliced"].astype(np.int64), unspliced=res["unspliced"].astype(np.int64))


def merge_tables(tables):
    """tables: list over ranks of local_tables().  Returns (global results dict, per-rank maps).
    maps[r] = (cluster_map[int32 local→global], sub_map[list per local cluster of int32 local k→global k])."""
    S = len(tables[0]["total_counts"])
    merged = {}
    for t in tables:
        for cl in t["clusters"]:
            m = merged.setdefault(cl["key"], dict(first=cl["first"], subs={}))
            m["first"] = min(m["first"], cl["first"])
            for sb in cl["subs"]:
                ms = m["subs"].get(sb["key"])
                if ms is None:
                    m["subs"][sb["key"]] = dict(first=sb["first"], count=sb["count"].copy(), break_sum=sb["break_sum"], sums=sb["sums"].copy(), flags=sb["flags"])
                else:
                    if sb["first"] < ms["first"]:
                        ms["first"], ms["flags"] = sb["first"], sb["flags"]
                    ms["count"] = ms["count"] + sb["count"]
                    ms["break_sum"] += sb["break_sum"]
                    if len(ms["sums"]) == len(sb["sums"]):
                        ms["sums"] = ms["sums"] + sb["sums"]
                    elif len(ms["sums"]) == 0:
                        ms["sums"] = sb["sums"].copy()
    order = sorted(merged.items(), key=lambda kv: kv[1]["first"])
    cid = {k: i for i, (k, _) in enumerate(order)}
    g = dict(n_clusters=len(order), cl_first_read=[], cl_key_off=[0], cl_key_tok=[], cl_read_count=[], cl_sub_off=[0], sub_first_read=[],
             sub_key_off=[0], sub_key=[], sub_count=[], sub_break_sum=[], sub_sum_off=[0], sub_sums=[], sub_flags=[])
    sid = {}
    for k, m in order:
        g["cl_first_read"].append(m["first"])
        g["cl_key_tok"] += list(k)
        g["cl_key_off"].append(len(g["cl_key_tok"]))
        subs = sorted(m["subs"].items(), key=lambda kv: kv[1]["first"])
        rc = np.zeros(S, np.int64)
        for j, (sk, sb) in enumerate(subs):
            sid[(k, sk)] = j
            g["sub_first_read"].append(sb["first"]); g["sub_key"] += list(sk); g["sub_key_off"].append(len(g["sub_key"]))
            g["sub_count"].append(sb["count"]); g["sub_break_sum"].append(sb["break_sum"]); g["sub_sums"] += list(sb["sums"])
            g["sub_sum_off"].append(len(g["sub_sums"])); g["sub_flags"].append(sb["flags"])
            rc += sb["count"]
        g["cl_read_count"].append(rc)
        g["cl_sub_off"].append(len(g["sub_first_read"]))
    g["n_subs"] = len(g["sub_first_read"])
    out = dict(n_clusters=g["n_clusters"], n_subs=g["n_subs"],
               cl_first_read=np.array(g["cl_first_read"], np.int64), cl_key_off=np.array(g["cl_key_off"], np.uint32),
               cl_key_tok=np.array(g["cl_key_tok"], np.int32), cl_read_count=np.array(g["cl_read_count"], np.int32).reshape(-1, S),
               cl_sub_off=np.array(g["cl_sub_off"], np.uint32), sub_first_read=np.array(g["sub_first_read"], np.int64),
               sub_key_off=np.array(g["sub_key_off"], np.uint32), sub_key=np.array(g["sub_key"], np.int32),
               sub_count=np.array(g["sub_count"], np.int32).reshape(-1, S), sub_break_sum=np.array(g["sub_break_sum"], np.int32),
               sub_sum_off=np.array(g["sub_sum_off"], np.uint32), sub_sums=np.array(g["sub_sums"], np.int64), sub_flags=np.array(g["sub_flags"], np.uint8))
    pairs = {}
    for t in tables:
        for k, v in t["pairs"].items():
            pairs[k] = pairs.get(k, 0) + v
    ks = sorted(pairs)
    for i, name in enumerate(("pair_src", "pair_level", "pair_start", "pair_end")):
        out[name] = np.array([k[i] for k in ks], np.int32)
    out["pair_count"] = np.array([pairs[k] for k in ks], np.int32)
    out["n_pairs"] = len(ks)
    out["total_counts"] = sum(t["total_counts"] for t in tables).astype(np.int32)
    out["spliced"] = sum(t["spliced"] for t in tables).astype(np.int32)
    out["unspliced"] = sum(t["unspliced"] for t in tables).astype(np.int32)
    maps = []
    for t in tables:
        cmap = np.array([cid[cl["key"]] for cl in t["clusters"]], np.int32)
        smap = [np.array([sid[(cl["key"], sb["key"])] for sb in cl["subs"]], np.int32) for cl in t["clusters"]]
        maps.append((cmap, smap))
    return out, maps


def remap_reads(res, cmap, smap):
    """Rewrite a rank's per-read local ids to global ids (in place on copies)."""
    cl, sb = res["cluster_id"].copy(), res["sub_id"].copy()
    ok = cl >= 0
    if len(cmap):
        off = np.zeros(len(smap) + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in smap])
        flat = np.concatenate(smap) if smap else np.zeros(0, np.int32)
        sb[ok] = flat[off[cl[ok]] + sb[ok]]
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

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def merge_across_ranks(eng, rank, world, timing=False, idx_base=0):
    """GPU path: merge the tables, permute + all-reduce the coverage rows and small counters in place on the device.
    Returns the merge time in ms (timing=True) or (global tables, (cluster_map, sub_map)) for this rank."""
    import torch
    import torch.distributed as dist
    from . import abi
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.fetch(abi.NPT_FETCH_TABLES)
    mine = local_tables(res, 0)  # first-read indices are already global (npt_set_read_index_base)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    glob, maps = merge_tables(gathered)
    cmap, smap = maps[rank]
    eng.remap_clusters(cmap, glob["n_clusters"])
    v = eng.device_view()
    if v.coverage and v.coverage_elems:
        cov = torch.as_tensor(_DevPtr(v.coverage, v.coverage_elems), device="cuda")
        dist.all_reduce(cov)
    small = torch.as_tensor(_DevPtr(v.small_counts, v.small_elems), device="cuda")
    dist.all_reduce(small)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    return ms if timing else (glob, (cmap, smap))
