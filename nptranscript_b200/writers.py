"""Cluster-level writers: npt_results -> the rows and files of 0.transcripts.txt.gz, 0.genes.txt.gz,
0.transcripts.fc.txt.gz, 0.readToCluster.txt.gz, 0.annot.txt.gz, and .npy mirrors of the matrices the reference puts into
0.clusters.h5 / 0.breakpoints.h5 (no libhdf5 in this image).  Host-side Python over the arrays the device returned; no
device work and nothing of the walk is recomputed here.

What is restated (J = /root/reference/java/src/main/java/npTranscript):
  CigarClusters.getConsensus   J/cluster/CigarClusters.java:230-262   order of the clusters: iteration order of the
                               ConcurrentHashMap<CigarHash, CigarCluster>, in coronavirus mode stably sorted by -readCountSum
  CigarClusters.process1       J/cluster/CigarClusters.java:138-196   one genes row, one transcripts row per sub-cluster in the
                               iteration order of the cluster's HashMap<CigarHash2, Count>, one transcripts.fc row
  CigarCluster.numIsoforms / getBreaks / exonCount                    J/cluster/CigarCluster.java:702-735
  Outputs (headers, file names)                                       J/cluster/Outputs.java:313-347, 435-449
java.util.HashMap and java.util.concurrent.ConcurrentHashMap iteration orders are restated from OpenJDK (8+): table of 16
doubled at 0.75 load, bucket = spread(hash) & (n - 1), chains in insertion order; HashMap.resize splits a chain into two
that keep their order, ConcurrentHashMap.transfer re-uses the trailing run of a chain and PREPENDS the nodes before it.
A bin that would become a red-black tree (8 colliding keys in a table of 64 or more) is not restated: TreeBinNotRestated.
"""
import gzip
import os

import numpy as np

from . import abi, render


class TreeBinNotRestated(NotImplementedError):
    pass


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def java_string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return _i32(h)


def java_list_hash(ints) -> int:
    """ArrayList<Integer>.hashCode() (AbstractList): h = 31 * h + e, from 1."""
    h = 1
    for e in ints:
        h = (31 * h + (int(e) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return _i32(h)


def _spread(h):
    h &= 0xFFFFFFFF
    return h ^ (h >> 16)


def hashmap_order(hashes):
    """Iteration order of a java.util.HashMap after inserting distinct keys with these hashCode()s in this order:
    a list of insertion indices."""
    n_tab, table, size = 16, [[] for _ in range(16)], 0
    for i, h in enumerate(hashes):
        sp = _spread(h)
        b = table[sp & (n_tab - 1)]
        b.append((sp, i))
        if len(b) > 8:   # binCount >= TREEIFY_THRESHOLD - 1 while appending the 9th node
            if n_tab < 64:
                n_tab, table = _hm_resize(table, n_tab)
            else:
                raise TreeBinNotRestated("HashMap bin with 9 colliding keys in a table of %d" % n_tab)
        size += 1
        if size > 0.75 * n_tab:
            n_tab, table = _hm_resize(table, n_tab)
    return [i for b in table for (_, i) in b]


def _hm_resize(table, n):
    new = [[] for _ in range(2 * n)]
    for j, b in enumerate(table):
        for sp, i in b:                      # lo / hi lists keep the chain's order
            new[j + n if sp & n else j].append((sp, i))
    return 2 * n, new


def concurrent_hashmap_order(hashes):
    """Iteration order of a ConcurrentHashMap (default constructor) after putting distinct keys with these hashCode()s in
    this order from one thread."""
    n_tab, table, size, size_ctl = 16, [[] for _ in range(16)], 0, 12
    for i, h in enumerate(hashes):
        sp = _spread(h) & 0x7FFFFFFF
        b = table[sp & (n_tab - 1)]
        b.append((sp, i))
        if len(b) >= 8 + 1:   # binCount counts the nodes walked: the 9th node makes it 8 = TREEIFY_THRESHOLD
            if n_tab < 64:
                # treeifyBin -> tryPresize(n << 1): c = tableSizeFor(3n + 1) = 4n, and the table is transferred until sizeCtl >= c,
                # i.e. three times (16 -> 128)
                c = 4 * n_tab
                while size_ctl < c:
                    n_tab, table = _chm_transfer(table, n_tab)
                    size_ctl = int(0.75 * n_tab)
            else:
                raise TreeBinNotRestated("ConcurrentHashMap bin with 9 colliding keys in a table of %d" % n_tab)
        size += 1
        while size >= size_ctl and n_tab < (1 << 30):
            n_tab, table = _chm_transfer(table, n_tab)
            size_ctl = int(0.75 * n_tab)
    return [i for b in table for (_, i) in b]


def _chm_transfer(table, n):
    new = [[] for _ in range(2 * n)]
    for j, b in enumerate(table):
        if not b:
            continue
        # lastRun: the longest tail whose nodes all go to the same half; it is re-used as it is, the nodes before it are
        # prepended one by one to the list of their half
        run_bit, last = b[0][0] & n, 0
        for k in range(1, len(b)):
            if (b[k][0] & n) != run_bit:
                run_bit, last = b[k][0] & n, k
        lo, hi = ([], list(b[last:])) if run_bit else (list(b[last:]), [])
        for sp, i in b[:last]:
            if sp & n:
                hi.insert(0, (sp, i))
            else:
                lo.insert(0, (sp, i))
        new[j], new[j + n] = lo, hi
    return 2 * n, new


def _java_round(x):
    import math
    return int(math.floor(x + 0.5))


def gff_span(breaks, forward, annotation, enforce_strand=False, tolerance1=5):
    """GFFAnnotation.getSpan (J/cluster/GFFAnnotation.java:219-292) for one read: which genes explain its breaks.
    Returns (span string, exon rows hit = what the call adds to CigarCluster.span, gene names in SortedSet order).
    Rows are the exon lines in file order; the greedy cover walks the HashMap<String, Set<Integer>> entries in iteration order,
    stably sorted by descending set size, and unions them into the FIRST entry's set until it holds every break that lies in
    an exon (the reference mutates that set; the count of genes taken is what matters)."""
    n = len(annotation.genes)
    ok = (lambda i: True) if not enforce_strand else (lambda i: bool(forward) == bool(annotation.strand[i]))
    start_, end_ = int(breaks[0]), int(breaks[-1])
    istart = next((i for i in range(n) if ok(i) and annotation.end[i] > start_ - tolerance1), -1)
    iend = next((i for i in range(n - 1, -1, -1) if ok(i) and annotation.start[i] < end_ + tolerance1), -1)
    if not (istart >= 0 and iend >= 0 and istart <= iend):
        return ".", [], []
    m, order, rows, in_exons = {}, [], [], set()
    for i in range(istart, iend + 1):
        for j, pos in enumerate(breaks):
            if annotation.start[i] <= pos + tolerance1 and annotation.end[i] >= pos - tolerance1:
                g = annotation.genes[i]
                if g not in m:
                    m[g] = set(); order.append(g)
                m[g].add(j); rows.append(i); in_exons.add(j)
    names = set()
    if m:
        it = [order[k] for k in hashmap_order([java_string_hash(g) for g in order])]   # entrySet() order
        vals = sorted(it, key=lambda g: -len(m[g]))                                   # Collections.sort is stable
        cover = m[vals[0]]
        names.add(vals[0])
        k = 0
        while len(cover) < len(in_exons):
            k += 1
            cover |= m[vals[k]]
            names.add(vals[k])
    gl = sorted(names)   # TreeSet<String>: String.compareTo = UTF-16 code unit order = Python's order for BMP text
    return ";".join(gl), sorted(set(rows)), gl


class ChromosomeReport:
    """Everything process1 prints for one chromosome, from one npt_results (tables + per-read arrays)."""

    def __init__(self, res, annotation, chrom_name, chrom_index, cfg, seqlen, src=None, read_names=None, source_names=None):
        self.res, self.annotation, self.chrom, self.ci, self.cfg, self.seqlen = res, annotation, chrom_name, chrom_index, cfg, seqlen
        self.S = cfg.num_sources
        self.src, self.read_names = src, read_names
        self.source_names = source_names or [f"bam{i}" for i in range(self.S)]
        self.keys = render.cluster_keys(res, annotation, chrom_index, bool(cfg.coronavirus))
        nc = res["n_clusters"]
        cl = res["cluster_id"]
        ok = cl >= 0
        self.start_pos = np.full(nc, np.iinfo(np.int32).max, np.int64)
        self.end_pos = np.full(nc, np.iinfo(np.int32).min, np.int64)
        np.minimum.at(self.start_pos, cl[ok], res["start_pos"][ok])   # CigarCluster.merge :633-634
        np.maximum.at(self.end_pos, cl[ok], res["end_pos"][ok])
        self._sub_name = None
        # GFF mode: the span of every read (IdentityProfile1.java:333) and of every cluster (CigarCluster.span, merged :640)
        self.read_span = None
        self.cluster_span = [set() for _ in range(nc)]
        if annotation.kind == abi.NPT_ANNOT_GFF:
            bo = res["break_off"].astype(np.int64)
            self.read_span = []
            for i in range(res["n_reads"]):
                if cl[i] < 0:
                    self.read_span.append((".", 0)); continue
                fwd = not (int(res["read_flags"][i]) & abi.NPT_RF_NEG)
                st, rows, names = gff_span([int(x) for x in res["breaks"][bo[i]:bo[i + 1]]], fwd, annotation, bool(cfg.enforce_strand))
                self.read_span.append((st, len(names)))
                self.cluster_span[int(cl[i])].update(rows)

    def _gene_names(self, c):
        """Annotation.getString(cc.span, geneNames) :257-268: genes of the span's exon rows in row order, ';'-joined (a gene
        appears once per row), and the number of distinct names."""
        rows = sorted(self.cluster_span[c])
        if not rows:
            return "-", 0
        g = [self.annotation.genes[i] for i in rows]
        return ";".join(g), len(set(g))

    # ---- ids
    def sub_names_final(self):
        """Count.id at end of chromosome: ".t<k>", or with --firstIsTranscriptome the name of the LAST source-0 read that hit the
        sub-cluster (CigarCluster.java:652-664)."""
        if self._sub_name is None:
            r = self.res
            so = r["cl_sub_off"].astype(np.int64)
            names = []
            for c in range(r["n_clusters"]):
                names += [render.sub_id(self.ci, c, k) for k in range(int(so[c + 1] - so[c]))]
            if self.cfg.first_is_transcriptome:
                if self.src is None or self.read_names is None:
                    raise ValueError("firstIsTranscriptome: per-read source indices and read names are needed for the sub-cluster ids")
                cl, sb = r["cluster_id"].astype(np.int64), r["sub_id"].astype(np.int64)
                m = (cl >= 0) & (np.asarray(self.src) == 0)
                g = so[cl[m]] + sb[m]
                last = {}
                for gi, ri in zip(g.tolist(), np.flatnonzero(m).tolist()):
                    last[gi] = ri
                for gi, ri in last.items():
                    names[gi] = self.read_names[ri]
            self._sub_name = names
        return self._sub_name

    def cluster_order(self):
        """Order in which getConsensus visits the clusters (ids)."""
        base = concurrent_hashmap_order([java_string_hash(k) for k in self.keys])
        if self.cfg.coronavirus:   # Collections.sort is stable
            tot = self.res["cl_read_count"].sum(axis=1)
            base = sorted(base, key=lambda c: -int(tot[c]))
        return base

    def sub_order(self, c):
        """Iteration order of cluster c's HashMap<CigarHash2, Count>: global sub indices."""
        r = self.res
        so, ko = r["cl_sub_off"].astype(np.int64), r["sub_key_off"].astype(np.int64)
        subs = range(int(so[c]), int(so[c + 1]))   # insertion order = .t order
        hs = [java_list_hash(r["sub_key"][ko[s]:ko[s + 1]]) for s in subs]
        return [int(so[c]) + i for i in hashmap_order(hs)]

    # ---- rows
    def rows(self):
        """(genes rows, transcripts rows, feature-count rows) in file order."""
        r, cfg = self.res, self.cfg
        so, ko = r["cl_sub_off"].astype(np.int64), r["sub_key_off"].astype(np.int64)
        names = self.sub_names_final()
        genes, trans, fc = [], [], []
        for c in self.cluster_order():
            order = self.sub_order(c)
            first_key = r["sub_key"][ko[so[c]]:ko[so[c] + 1]]          # cc.breaks = cloneBreaks of the read that created the cluster
            forward = bool(r["sub_flags"][so[c]] & 1)
            leader = bool(cfg.coronavirus) and len(first_key) > 1 and self._is_leader(int(first_key[1]) * cfg.bin)
            st, en = int(self.start_pos[c]), int(self.end_pos[c])
            type_ind = (0 if en >= self.seqlen - cfg.end_thresh else 1) if st <= cfg.start_thresh else (2 if en >= self.seqlen - cfg.end_thresh else 3)
            type_nme = render.TYPE_NAMES[type_ind]
            gene_nme, n_gene = self._gene_names(c)   # "-" / 0 on the empty span of CSV and empty annotations
            read_count = "\t".join(str(int(x)) for x in r["cl_read_count"][c])
            best, best_sum = None, 0
            for s in order:
                cnt = r["sub_count"][s]
                ssum = int(cnt.sum())
                br = render.consensus_breaks(r, s, self.src, bool(cfg.first_is_transcriptome))
                if br is None:
                    raise ValueError("transcripts rows need the break sums (track_break_sums = writeGFF || writeIsoforms): Count.getBreaks() is null otherwise")
                if ssum > best_sum:
                    best, best_sum = br, ssum
                key = ",".join(str(int(x)) for x in r["sub_key"][ko[s]:ko[s + 1]])
                trans.append("\t".join([names[s], self.chrom, str(int(br[0])), str(int(br[-1])), gene_nme, str(_java_round(len(br) / 2.0)), "1",
                                        "1" if leader else "0", key, names[s], str(n_gene), "NA", str(ssum), "\t".join(str(int(x)) for x in cnt)]))
            exon_counts = sorted({_java_round((int(ko[s + 1]) - int(ko[s])) / 2.0) for s in order})
            n_iso = _java_round((len(order) - 2.0) / 2.0)
            genes.append("\t".join([render.cluster_id(self.ci, c), self.chrom, str(st), str(en), type_nme, ",".join(map(str, exon_counts)), str(n_iso),
                                    "1" if leader else "0", self.keys[c], gene_nme, str(n_gene), "-1", str(int(r["cl_read_count"][c].sum())), read_count]))
            if best is not None:
                pairs = [(int(best[i]), int(best[i + 1])) for i in range(0, len(best) - 1, 2)]
                ln = sum(b - a for a, b in pairs)
                fc.append("\t".join([self.keys[c], ";".join([self.chrom] * len(pairs)), ";".join(str(a) for a, _ in pairs), ";".join(str(b) for _, b in pairs),
                                     "+;+", str(ln), read_count]))
            else:
                fc.append("\t".join([self.keys[c], str(en), str(st), str(en), "+;+", "0", read_count]))
        return genes, trans, fc

    def gff_bed(self, gff_thresh=(1,), max_transcripts=1):
        """CigarCluster.writeGFF / writeGFF1 / printBed (J/cluster/CigarCluster.java:143-330) for every cluster in getConsensus
        order: (lines of each GFF writer — one, or two with --firstIsTranscriptome —, lines of each source's BED file).
        gff_thresh = --gffThresh a[:b], max_transcripts = --maxTranscriptsPerGeneInGFF.  Literal quirks kept: the threshold
        loop looks at the first `number of writers` sources only; a transcript written by both writers reaches the BED files
        twice; BED colour is always that of source 0.  The <type>.ref.fa sequences written alongside are not produced."""
        r, cfg = self.res, self.cfg
        nw = 2 if cfg.first_is_transcriptome else 1
        gff = [[] for _ in range(nw)]
        bed = [[] for _ in range(self.S)]
        so = r["cl_sub_off"].astype(np.int64)
        names = self.sub_names_final()

        def line(k, typ, parent, ID, start, end, type_nme, key2, gene_id, strand, counts, breaks):
            cs = ",".join(str(int(x)) for x in counts)
            gn = key2.replace(";", "_")
            gff[k].append(f"{self.chrom}\tnp\t{typ}\t{start}\t{end}\t.\t{strand}\t.\tID={ID}" + (f";Parent={parent}" if parent is not None else "") +
                          f";gene_id={ID if parent is None else parent};gene_type={type_nme};type=ORF;gene_name={gn};count={cs}")
            if breaks is None:
                return
            for i in range(0, len(breaks) - 1, 2):
                gff[k].append(f"{self.chrom}\tnp\texon\t{int(breaks[i])}\t{int(breaks[i + 1])}\t.\t{strand}\t.\tID={ID}.e.{i};Parent={ID};gene_id={gene_id};"
                              f"type=ORF;gene_name={gn};count={cs}")
            sp, ep = int(breaks[0]) - 1, int(breaks[-1]) - 1
            sizes = ",".join(str(int(breaks[i + 1]) - int(breaks[i])) for i in range(0, len(breaks) - 1, 2))
            starts = ",".join(str(int(breaks[i]) - 1 - sp) for i in range(0, len(breaks) - 1, 2))
            for i in range(self.S):
                if counts[i] > 0:
                    bed[i].append(f"{self.chrom}\t{sp}\t{ep}\t{ID}.{gene_id}\t{int(counts[i])}\t{strand}\t{sp}\t{ep}\t0,0,0\t{len(breaks) // 2}\t{sizes}\t{starts}")

        for c in self.cluster_order():
            st, en = int(self.start_pos[c]), int(self.end_pos[c])
            type_ind = (0 if en >= self.seqlen - cfg.end_thresh else 1) if st <= cfg.start_thresh else (2 if en >= self.seqlen - cfg.end_thresh else 3)
            type_nme = render.TYPE_NAMES[type_ind]
            strand = "-" if r["sub_flags"][so[c]] & 2 else "+"
            cid = render.cluster_id(self.ci, c)
            subs = sorted(self.sub_order(c), key=lambda s_: int(r["sub_count"][s_].sum()))   # Collections.sort by Count.sum(), stable
            if int(r["sub_count"][subs[-1]].sum()) < 1:
                continue
            write_any, starts, ends = [False] * nw, [2 ** 31 - 1] * nw, [0] * nw
            for i in range(len(subs) - 1, len(subs) - 1 - max_transcripts, -1):
                if i < 0:
                    break
                s_ = subs[i]
                cnt = r["sub_count"][s_]
                wg = [any(int(cnt[j]) >= gff_thresh[0] for j in range(min(nw, self.S)))] * nw
                if cfg.first_is_transcriptome:
                    t1 = gff_thresh[1]
                    wg[0] = int(cnt[0]) >= 1 and any(int(cnt[j]) >= t1 for j in range(1, min(nw, self.S)))
                br = render.consensus_breaks(r, s_, self.src, bool(cfg.first_is_transcriptome))
                if br is None:
                    raise ValueError("GFF output needs the break sums (track_break_sums)")
                for k in range(nw):
                    if wg[k]:
                        write_any[k] = True
                        starts[k], ends[k] = min(starts[k], int(br[0])), max(ends[k], int(br[-1]))
                        line(k, "transcript", cid, names[s_], int(br[0]), int(br[-1]), type_nme, self.keys[c], names[s_], strand, cnt, br)
            for k in range(nw):
                if write_any[k]:
                    line(k, "gene", None, cid, starts[k], ends[k], type_nme, self.keys[c], cid, strand, r["cl_read_count"][c], None)
        return gff, bed

    def _is_leader(self, pos):
        if self.annotation.kind == abi.NPT_ANNOT_EMPTY:
            return pos < 100   # EmptyAnnotation.isLeader
        return pos < self.annotation.start[1] - self.cfg.bin   # Annotation.isLeader :214-216 (tolerance = --bin)

    # ---- files
    def headers(self):
        S = self.S
        nme = "".join(n + "\t" for n in self.source_names)
        counts = "\t".join(f"count{i}" for i in range(S))
        th = "ID\tchrom\tstart\tend\tgene_nme\tnum_exons\tisoforms\tleader_break\tORFs\tspan\tspan_length\ttotLen\tcountTotal\t" + counts
        gh = "ID\tchrom\tstart\tend\ttype_nme\tnum_exons\tisoforms\tleader_break\tORFs\tspan\tspan_length\ttotLen\tcountTotal\t" + counts
        return dict(transcripts=["#" + nme, th], genes=["#" + nme, gh], fc=["#npTranscript output", "Geneid\tChr\tStart\tEnd\tStrand\tLength\t" + nme])

    def write(self, resdir, genome_index=0, polya=0, npy=True, gff_thresh=None, max_transcripts=1):
        """Writes <genome_index>transcripts.txt.gz, genes.txt.gz, transcripts.fc.txt.gz, readToCluster.txt.gz (when read names
        were given), annot.txt.gz, with gff_thresh (--writeGFF) <k>.gff.gz and <genome_index><k>.bed.gz (the "#date:" header line of the
        GFF is left out), and .npy mirrors of the break-point and depth matrices.  Returns the list of paths."""
        os.makedirs(resdir, exist_ok=True)
        genes, trans, fc = self.rows()
        h = self.headers()
        out = []

        def gz(name, lines):
            p = os.path.join(resdir, f"{genome_index}{name}.txt.gz")
            with gzip.open(p, "wt") as f:
                for ln in lines:
                    f.write(ln + "\n")
            out.append(p)

        gz("transcripts", h["transcripts"] + trans)
        gz("genes", h["genes"] + genes)
        gz("transcripts.fc", h["fc"] + fc)
        gz("annot", render.annot_rows(self.res, self.annotation, seqlen=self.seqlen, polya=polya, coronavirus=bool(self.cfg.coronavirus)))
        if self.read_names is not None:
            r = dict(self.res)
            if self.src is not None:
                r["src"] = np.asarray(self.src)
            head = ("readID\tclusterId\tsubID\tsource\tlength\tstart_read\tend_read\ttype_nme\tchrom\tstartPos\tendPos\tstrand\tnum_breaks\tleader_break\t"
                    "errorRatio\tORFs\tstrand\tbreaks\tspan\tspan_count\tqval")
            gz("readToCluster", [head] + render.read_to_cluster_rows(r, self.read_names, self.annotation, self.chrom, self.ci, bool(self.cfg.record_depth),
                                                                     coronavirus=bool(self.cfg.coronavirus), spans=self.read_span))
        if gff_thresh is not None:   # --writeGFF
            gff, bed = self.gff_bed(gff_thresh, max_transcripts)
            for k, lines in enumerate(gff):
                p = os.path.join(resdir, f"{k}.gff.gz")
                with gzip.open(p, "wt") as f:
                    f.write("##gff-version 3\n#description: \n#provider: \n#contact: \n#format: gff3\n")
                    f.write(f"##sequence-region {self.chrom} 0 {self.seqlen}\n")
                    for ln in lines:
                        f.write(ln + "\n")
                out.append(p)
            for k, lines in enumerate(bed):
                p = os.path.join(resdir, f"{genome_index}{k}.bed.gz")
                with gzip.open(p, "wt") as f:
                    f.write(f'track name="{self.source_names[k]}" description="{self.source_names[k]}" itemRgb="On" \n')
                    for ln in lines:
                        f.write(ln + "\n")
                out.append(p)
        if npy:
            for (s, j), m in render.breakpoint_matrices(self.res, self.S).items():
                p = os.path.join(resdir, f"{genome_index}.breakpoints.{self.source_names[s]}.{j}.npy")
                np.save(p, m); out.append(p)
            if self.cfg.record_depth and "depth" in self.res:
                for c in range(self.res["n_clusters"]):
                    safe = self.keys[c].replace("/", "_")
                    for kind, m in (("depth", render.depth_matrix(self.res, c)), ("depthStart", render.start_end_matrix(self.res, c, "map_start")),
                                    ("depthEnd", render.start_end_matrix(self.res, c, "map_end"))):
                        p = os.path.join(resdir, f"{genome_index}.clusters.{kind}.{safe}.npy")
                        np.save(p, m); out.append(p)
        return out
