"""Host-side rendering: npt_results → the strings, ids and matrices the reference's writers produce.

These are the "host duties that stay in Java" of SURVEY §8(b) restated in Python so that results can be diffed against
the reference's files: cluster key strings (CigarHash.secondKey), ID strings, 0.readToCluster rows, the break-point
matrix layout of Outputs.printMatrix, the depth matrix of Outputs.writeDepthH5, 0.annot rows and the consensus breaks
of Count.getBreaks.  J = /root/reference/java/src/main/java/npTranscript.  Pure numpy/Python; no device work here.
"""
import math

import numpy as np

from . import abi

TYPE_NAMES = ["5_3", "5_no3", "no5_3", "no5_no3"]  # Annotation.nmes (J/cluster/Annotation.java:239)


def java_g53(x: float) -> str:
    """String.format("%5.3g", x).trim() as Java formats it (no trailing-zero stripping, HALF_UP rounding)."""
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0.00"
    from decimal import ROUND_HALF_UP, Decimal
    d = Decimal(repr(float(x)))
    sign = "-" if d < 0 else ""
    d = abs(d)
    e = d.adjusted()
    r = d.quantize(Decimal(1).scaleb(e - 2), rounding=ROUND_HALF_UP)  # 3 significant digits
    if r.adjusted() != e:
        e = r.adjusted()
        r = r.quantize(Decimal(1).scaleb(e - 2), rounding=ROUND_HALF_UP)
    if Decimal("0.0001") <= r < Decimal(1000):
        places = max(0, 2 - e)
        return sign + f"{r:.{places}f}"
    mant = r.scaleb(-e)
    return sign + f"{mant:.2f}e{'+' if e >= 0 else '-'}{abs(e):02d}"


def token_name(tok, annotation, chrom_index):
    """One key token → the name the reference concatenates (Annotation.java:71,96; EmptyAnnotation.java:19,23)."""
    if annotation.kind == abi.NPT_ANNOT_EMPTY and tok >= 0:
        return f"{chrom_index}.{tok}"
    if tok == abi.NPT_TOK_ST:
        return "st" + (f".{chrom_index}" if chrom_index > 0 else "")
    if tok == abi.NPT_TOK_END:
        return "end" + (f".{chrom_index}" if chrom_index > 0 else "")
    return annotation.genes[tok]


def key_string(tokens, annotation, chrom_index, coronavirus=True):
    """Token list → CigarHash.secondKey (grammar of J/cluster/IdentityProfile1.java:266-299, 325-327)."""
    toks = [int(t) for t in tokens]
    suffix = ""
    if toks and toks[-1] in (abi.NPT_TOK_PLUS, abi.NPT_TOK_MINUS):
        suffix = "+" if toks[-1] == abi.NPT_TOK_PLUS else "-"
        toks = toks[:-1]
    nm = lambda t: token_name(t, annotation, chrom_index)
    if annotation.kind == abi.NPT_ANNOT_GFF:  # :300-323: gene names in HashSet order, else the coordinate of one block end
        if toks and toks[0] == abi.NPT_TOK_GENESET:
            return "".join(annotation.genes[t] + "_" for t in toks[1:]) + suffix
        return f"{chrom_index}.{toks[0]}" + suffix
    if not coronavirus:
        return "".join(nm(t) for t in toks) + suffix
    if toks and toks[0] == abi.NPT_TOK_NOENDS:
        body = toks[1:]
        return "".join(f"{nm(body[i])},{nm(body[i + 1])}_" for i in range(0, len(body) - 1, 2)) + suffix
    s = nm(toks[0]) + "_"
    mid = toks[1:-1]
    s += "".join(f"{nm(mid[i])},{nm(mid[i + 1])}_" for i in range(0, len(mid) - 1, 2))
    return s + nm(toks[-1]) + suffix


def cluster_keys(res, annotation, chrom_index, coronavirus=True):
    off = res["cl_key_off"]
    return [key_string(res["cl_key_tok"][off[c]:off[c + 1]], annotation, chrom_index, coronavirus) for c in range(res["n_clusters"])]


def check_key_aliasing(keys):
    """Distinct token lists must render to distinct strings (a gene name containing '_' or ',' could alias)."""
    if len(set(keys)) != len(keys):
        raise ValueError("two clusters render to the same key string: gene names contain key delimiters; merge on the host")


def cluster_id(chrom_index, c):
    return f"ID{chrom_index}.{c}"  # CigarClusters.java:83


def sub_id(chrom_index, c, k):
    return f"ID{chrom_index}.{c}.t{k}"  # CigarCluster.java:660, CigarClusters.java:85


def sub_names(res, src, read_names, chrom_index, first_is_transcriptome=False):
    """Per-read subID column.  Normally "ID<chr>.<c>.t<k>".  With --firstIsTranscriptome a sub-cluster is named after the
    most recent source-0 read seen in it so far, this read included (CigarClusters.java:85, CigarCluster.java:657-669)."""
    n = res["n_reads"]
    cl, sb = res["cluster_id"].astype(np.int64), res["sub_id"].astype(np.int64)
    base = [sub_id(chrom_index, int(c), int(k)) if c >= 0 else "" for c, k in zip(cl, sb)]
    if not first_is_transcriptome:
        return base
    g = np.where(cl >= 0, res["cl_sub_off"].astype(np.int64)[np.maximum(cl, 0)] + sb, -1)
    order = np.argsort(g, kind="stable")
    gs = g[order]
    mark = np.where(np.asarray(src)[order] == 0, order, -1)        # read index where source 0, else -1
    # running maximum inside each group of equal g: offset every group so that earlier groups cannot leak in
    grp = np.cumsum(np.concatenate([[0], gs[1:] != gs[:-1]]))
    run = np.maximum.accumulate(mark + grp * (n + 1)) - grp * (n + 1)
    last0 = np.full(n, -1, np.int64)
    last0[order] = np.where(run >= 0, run, -1)
    return [read_names[int(j)] if (j >= 0 and c >= 0) else b for j, b, c in zip(last0, base, cl)]


def read_to_cluster_rows(res, read_names, annotation, chrom_name, chrom_index, record_depth, qvals=None, coronavirus=True, spans=None):
    """Rows of 0.readToCluster.txt.gz exactly as IdentityProfile1.commit builds them (:133-136); header Outputs.java:313-318.
    span columns are "." / 0 for CSV and empty annotations (Annotation.getSpan :252-255); GFF mode passes spans = per-read
    (span string, number of genes) from writers.gff_span."""
    keys = cluster_keys(res, annotation, chrom_index, coronavirus)
    rows = []
    bo = res["break_off"]
    for i in range(res["n_reads"]):
        fl = int(res["read_flags"][i])
        if fl & abi.NPT_RF_BADCIGAR:
            continue
        c, k = int(res["cluster_id"][i]), int(res["sub_id"][i])
        st_r, en_r = int(res["read_st"][i]), int(res["read_en"][i])
        strand = "-" if fl & abi.NPT_RF_NEG else "+"
        type_nme = TYPE_NAMES[fl & abi.NPT_RF_TYPE_MASK] if coronavirus else "NA"
        if record_depth:
            m, e = int(res["maps_sum"][i]), int(res["err_sum"][i])
            err = "NaN" if m == 0 else java_g53(e / m)  # CigarCluster.getError :538-544 (no trim there, width 5)
            if m != 0 and len(err) < 5:
                err = err.rjust(5)
        else:
            err = "NA"
        breaks = ",".join(str(int(x)) for x in res["breaks"][int(bo[i]):int(bo[i + 1])])
        q = java_g53(qvals[i]) if qvals is not None else java_g53(500.0)
        rows.append("\t".join([read_names[i], cluster_id(chrom_index, c), sub_id(chrom_index, c, k), str(int(res.get("src", np.zeros(res["n_reads"], int))[i])),
                               str(en_r - st_r), str(st_r), str(en_r), type_nme, chrom_name, str(int(res["start_pos"][i])),
                               str(int(res["end_pos"][i])), strand, "0", "1" if fl & abi.NPT_RF_LEADER else "0", err, keys[c], strand, breaks,
                               spans[i][0] if spans is not None else ".", str(spans[i][1]) if spans is not None else "0", q]))
    return rows


def breakpoint_matrices(res, num_sources, levels=2):
    """(src, level) → int matrix laid out as Outputs.printMatrix (J/cluster/Outputs.java:461-489): row 0 = end positions,
    row 1 = end totals, col 0 = start positions, col 1 = start totals, body = counts."""
    out = {}
    ps, pl, pa, pe, pc = (res[k] for k in ("pair_src", "pair_level", "pair_start", "pair_end", "pair_count"))
    for s in range(num_sources):
        for j in range(levels):
            m = (ps == s) & (pl == j)
            st, en, ct = pa[m], pe[m], pc[m]
            rows, cols = np.unique(st), np.unique(en)
            mat = np.zeros((len(rows) + 2, len(cols) + 2), dtype=np.int32)
            mat[0, 2:] = cols
            mat[2:, 0] = rows
            ri, ci = np.searchsorted(rows, st), np.searchsorted(cols, en)
            np.add.at(mat, (ri + 2, ci + 2), ct)
            mat[1, 2:] = mat[2:, 2:].sum(axis=0)   # breakEnd totals
            mat[2:, 1] = mat[2:, 2:].sum(axis=1)   # breakSt totals
            out[(s, j)] = mat
    return out


def depth_matrix(res, cluster):
    """depth/<key> matrix of Outputs.writeDepthH5 (:559-593): rows = positions with total depth > 0 ascending,
    columns [pos, depth_s0, err_s0, depth_s1, err_s1, ...]."""
    d, e = res["depth"][cluster], res["errors"][cluster]  # [S][stride]
    pos = np.flatnonzero(d.sum(axis=0) > 0)
    S = d.shape[0]
    mat = np.zeros((len(pos), 1 + 2 * S), dtype=np.int32)
    mat[:, 0] = pos
    for s in range(S):
        mat[:, 1 + 2 * s] = d[s, pos]
        mat[:, 2 + 2 * s] = e[s, pos]
    return mat


def start_end_matrix(res, cluster, which="map_start"):
    """depthStart/<key> and depthEnd/<key> (:575-590): rows = positions with an entry, columns [pos, count_s0, ...]."""
    a = res[which][cluster]
    pos = np.flatnonzero(a.sum(axis=0) > 0)
    return np.concatenate([pos[:, None].astype(np.int32), a[:, pos].T.astype(np.int32)], axis=1)


def annot_rows(res, annotation, seqlen=None, polya=0, coronavirus=True):
    """0.annot.txt.gz (Annotation.print, J/cluster/Annotation.java:40-58).  In coronavirus mode the first processed read
    moves the last row's end to seqlen - polyA (adjust3UTR, Annotation.java:205-212 via IdentityProfile1.java:179-183):
    pass the reference length and its polyA tail (workloads.polya_len) to get the End column the reference prints."""
    if coronavirus and seqlen is not None and polya > 0 and res["n_reads"] > 0:
        import copy
        annotation = copy.deepcopy(annotation).adjust3utr(seqlen - polya)
    S = res["spliced"].shape[0]
    head = "Gene\tStart\tEnd" + "".join(f"\tSpliced_5_{j}\tUnspliced_5_{j}" for j in range(S))
    rows = [head]
    for i, g in enumerate(annotation.genes):
        rows.append(f"{g}\t{annotation.start[i]}\t{annotation.end[i]}" + "".join(f"\t{res['spliced'][j, i]}\t{res['unspliced'][j, i]}" for j in range(S)))
    return rows


def consensus_breaks(res, sub_index, src=None, first_is_transcriptome=False):
    """Count.getBreaks (J/cluster/CigarCluster.java:99-108): Math.round(true_breaks[i] * (1e3 / break_sum)) where
    true_breaks[i] is the arrival-order double sum of break/1e3.  The device returns exact int64 sums; the double sum
    differs from the exact value by < n·2^-50 relative, so the rounding can only differ when the exact mean is within
    that distance of a .5 boundary — those (rare, mostly tiny sub-clusters) are replayed in arrival order from the
    per-read break lists."""
    so = res["sub_sum_off"]
    sums = res["sub_sums"][int(so[sub_index]):int(so[sub_index + 1])]
    n = int(res["sub_break_sum"][sub_index])
    if n == 0 or len(sums) == 0:
        return None
    out = np.zeros(len(sums), dtype=np.int64)
    need_replay = False
    for i, s in enumerate(sums):
        s = int(s)
        q, rem = divmod(2 * s + n, 2 * n)  # floor(s/n + 1/2)
        out[i] = q
        # distance of the exact mean from the rounding boundary, in units of 1/(2n); FP error bound ~ n * 2^-50 * mean
        dist = min(rem, 2 * n - rem)
        if dist * (2 ** 40) <= 2 * n * n * max(1, s // max(n, 1)):
            need_replay = True
    if not need_replay:
        return out
    # arrival-order replay (needs the per-read arrays)
    cl = int(np.searchsorted(res["cl_sub_off"], sub_index, side="right") - 1)
    k = sub_index - int(res["cl_sub_off"][cl])
    idx = np.flatnonzero((res["cluster_id"] == cl) & (res["sub_id"] == k))
    if first_is_transcriptome:
        # only source-0 reads and the reads that came before the sub-cluster's first source-0 read count (CigarCluster.java:85)
        s_ = np.asarray(src)[idx]
        first0 = idx[s_ == 0].min() if (s_ == 0).any() else np.iinfo(np.int64).max
        idx = idx[(s_ == 0) | (idx < first0)]
    bo = res["break_off"]
    tb = [0.0] * len(sums)
    for r in idx:
        br = res["breaks"][int(bo[r]):int(bo[r + 1])]
        if len(br) == len(tb):
            for i in range(len(tb)):
                tb[i] += int(br[i]) / 1e3
    mult = 1e3 / float(n)
    return np.array([int(math.floor(t * mult + 0.5)) for t in tb], dtype=np.int64)
