"""Second, independent restatement of the reference path in pure Python (small inputs only).

TEST INFRASTRUCTURE ONLY — imported by tests/ to cross-check oracle.cpp; PARITY UNPINNED like oracle.cpp (no runnable
reference, no golden vectors in the reference).  Written from the Java sources without looking at oracle.cpp's structure:
it keeps Java's own shapes — string cluster keys built from gene names, lists of Integers, dicts keyed by tuples — so
that a token/ordering mistake in the C++ oracle or the device path shows up as a difference here.
J = /root/reference/java/src/main/java/npTranscript.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

OPS = "MIDNSHP=X"


@dataclass
class Flags:
    """The static configuration the path reads (J/run/ViralTranscriptAnalysisCmd2.java:221-386, 496-511)."""
    bin: int = 1
    break_thresh: int = 1000
    include_start: bool = True
    coronavirus: bool = True
    start_thresh: int = 100
    end_thresh: int = 100
    enforce_strand: bool = False
    record_depth: bool = False
    calc_breaks: bool = True
    write_gff_or_isoforms: bool = True
    first_is_transcriptome: bool = False
    num_sources: int = 1
    rna: Tuple[bool, ...] = (True,) * 16

    @property
    def min_first_last_exon_length(self):
        return 0 if self.coronavirus else 10  # :339, :360


def jround(pos, rnd):
    """TranscriptUtils.round (J/cluster/TranscriptUtils.java:20-23)."""
    return int(math.floor(float(pos) / float(rnd)))


def java_string_hash(s):
    """String.hashCode() as a signed 32-bit value computed mod 2^32."""
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def java_hashset_order(items):
    """Iteration order of a java.util.HashSet<String> after add()ing `items` in order (OpenJDK 8+ HashMap):
    16 buckets, load factor 0.75, bucket = (h ^ h>>>16) & (cap-1), chains in insertion order, a chain that already holds
    8 nodes at an append doubles the table while it is smaller than 64 (treeifyBin -> resize).  Returns None if a real
    tree bin would be needed (not restated)."""
    cap, chains, size = 16, {}, 0
    spread = lambda x: (java_string_hash(x) ^ (java_string_hash(x) >> 16)) & 0xFFFFFFFF

    def rehash(newcap, old):
        new = {}
        for b in sorted(old):
            for x in old[b]:
                new.setdefault(spread(x) & (newcap - 1), []).append(x)
        return new

    for x in items:
        b = spread(x) & (cap - 1)
        ch = chains.setdefault(b, [])
        if x in ch:
            continue
        before = len(ch)
        ch.append(x)
        if before >= 8:
            if cap >= 64:
                return None
            cap *= 2
            chains = rehash(cap, chains)
        size += 1
        if size > cap * 3 // 4:
            cap *= 2
            chains = rehash(cap, chains)
    return [x for b in sorted(chains) for x in chains[b]]


class Annot:
    """J/cluster/Annotation.java (CSV rows) / EmptyAnnotation.java (no rows)."""

    def __init__(self, genes, start, end, strand, seqlen, flags: Flags, empty=False, chrom_index=0, gff=False):
        self.genes, self.start, self.end, self.strand = list(genes), list(start), list(end), list(strand)
        self.seqlen, self.f, self.empty, self.chrom_index, self.gff = seqlen, flags, empty, chrom_index, gff
        n_orf = 0 if gff else len(self.start)
        self.spliced = [[0] * n_orf for _ in range(flags.num_sources)]
        self.unspliced = [[0] * n_orf for _ in range(flags.num_sources)]

    # ---- GFFAnnotation.java (host mode)
    def gff_contains(self, l2, r2, i):  # :72-76
        t = self.f.bin
        return r2 - t <= self.end[i] and r2 + t >= self.start[i] and l2 - t <= self.end[i] and l2 + t >= self.start[i]

    def gff_next_upstream(self, l1, r1, forward):  # :80-100; exon row or None
        if l1 < 0:
            return None
        best = {}  # TreeMap<score,row>: later put (lower row) wins on equal score
        for i in range(len(self.start) - 1, -1, -1):
            if forward is None or forward == self.strand[i]:
                if self.gff_contains(l1, r1, i):
                    best[abs(l1 - self.start[i]) + abs(r1 - self.end[i])] = i
        if best:
            k = min(best)
            if k <= 2 * self.f.bin:
                return best[k]
        return None

    def next_downstream(self, right_break, forward):  # :62-72 / Empty :18-20
        if self.empty:
            return f"{self.chrom_index}.{jround(right_break, self.f.bin)}"
        for i in range(len(self.start)):
            if (not self.f.enforce_strand) or forward == self.strand[i]:
                if right_break - self.f.bin <= self.start[i]:
                    return self.genes[i]
        return "end" + (f".{self.chrom_index}" if self.chrom_index > 0 else "")

    def next_upstream(self, left_break, forward):  # :87-97 / Empty :22-24
        if self.empty:
            return f"{self.chrom_index}.{jround(left_break, self.f.bin)}"
        for i in range(len(self.start) - 1, -1, -1):
            if (not self.f.enforce_strand) or forward == self.strand[i]:
                if left_break + self.f.bin >= self.start[i]:
                    return self.genes[i]
        return "st" + (f".{self.chrom_index}" if self.chrom_index > 0 else "")

    def is_leader(self, prev_pos):  # :214-216 / Empty :28-30
        return prev_pos < 100 if self.empty else prev_pos < self.start[1] - self.f.bin

    def type_ind(self, st, en):  # :233-236
        if st <= self.f.start_thresh:
            return 0 if en >= self.seqlen - self.f.end_thresh else 1
        return 2 if en >= self.seqlen - self.f.end_thresh else 3


@dataclass
class Cluster:
    id: str
    key: str
    breaks: List[int]
    read_count: List[int]
    subs: Dict[Tuple[int, ...], dict] = field(default_factory=dict)  # insertion order = .t order
    depth: Dict[Tuple[int, int], int] = field(default_factory=dict)   # (src, pos) -> count
    errors: Dict[Tuple[int, int], int] = field(default_factory=dict)
    map_start: Dict[Tuple[int, int], int] = field(default_factory=dict)
    map_end: Dict[Tuple[int, int], int] = field(default_factory=dict)


def read_position_at_reference_position(cig, aln_start, pos):
    """htsjdk SAMRecord.getReadPositionAtReferencePosition over getAlignmentBlocks()."""
    if pos <= 0:
        return 0
    read_base, ref_base = 1, aln_start
    for ln, op in cig:
        c = OPS[op]
        if c in "SI":
            read_base += ln
        elif c in "ND":
            ref_base += ln
        elif c in "M=X":
            if ref_base + ln - 1 >= pos:
                return 0 if pos < ref_base else pos - ref_base + read_base
            read_base += ln
            ref_base += ln
    return 0


class Run:
    """One chromosome processed sequentially (IdentityProfileHolder + CigarClusters + BreakPoints)."""

    def __init__(self, flags: Flags, ref_codes, annot: Annot, chrom_index=0):
        self.f, self.ref, self.annot, self.chrom_index = flags, ref_codes, annot, chrom_index
        self.clusters: Dict[str, Cluster] = {}
        self.total_counts = [0] * flags.num_sources
        self.pairs: Dict[Tuple[int, int, int, int], int] = {}
        self.calc_breaks1 = flags.calc_breaks and flags.coronavirus and flags.break_thresh < len(ref_codes)  # :882
        self.rows = []  # per read: dict

    def process(self, pos, flag, src, cig, bases, read_name=None):
        f, T, G = self.f, self.f.break_thresh, len(self.ref)
        if pos <= 0 or any(op > 8 for _, op in cig):
            self.rows.append(dict(ok=False))
            return
        # ---- processCigar (J/cluster/IdentityProfile1.java:482-585) with CigarCluster.add (:470-498)
        breaks, prev = [pos], pos
        events = []  # (pos, match)
        ms, me = [], []

        def add(p, match):
            nonlocal prev
            break_p = prev < 0 or p - prev > T
            if f.record_depth:
                events.append((p, match))
                if break_p and prev > 0:
                    ms.append(p)
                    me.append(prev)
            if match:
                if break_p:
                    if prev > 0:
                        breaks.append(prev)
                    breaks.append(p)
                prev = p

        read_pos, ref_pos, span = 0, pos - 1, 0
        for ln, op in cig:
            c = OPS[op]
            if c in "HP":
                pass
            elif c == "S":
                read_pos += ln
            elif c == "N":
                ref_pos += ln
                span += ln
            elif c == "D":
                ref_pos += ln
                span += ln
                i = 0
                while i < ln and ref_pos + i < G:
                    add(ref_pos + i + 1, False)
                    i += 1
            elif c == "I":
                read_pos += ln
            elif c == "M":
                n_emit = max(0, min(ln, G - ref_pos))
                if n_emit > 0 and read_pos + n_emit > len(bases):  # getBase past the end of the read throws in the reference
                    self.rows.append(dict(ok=False))
                    return
                i = 0
                while i < ln and ref_pos + i < G:
                    add(ref_pos + i + 1, self.ref[ref_pos + i] == bases[read_pos + i])
                    i += 1
                read_pos += ln
                ref_pos += ln
                span += ln
            elif c == "=":
                read_pos += ln
                ref_pos += ln
                span += ln
                i = 0
                while i < ln and ref_pos + i < G:
                    add(ref_pos + i + 1, True)
                    i += 1
            elif c == "X":
                read_pos += ln
                ref_pos += ln
                span += ln
                i = 0
                while i < ln and ref_pos + i < G:
                    add(ref_pos + i + 1, False)
                    i += 1
        aln_end = pos + span - 1
        breaks.append(aln_end)
        st_r = read_position_at_reference_position(cig, pos, pos)
        end_r = read_position_at_reference_position(cig, pos, aln_end)
        # ---- processRefPositions (:149-356)
        neg = bool(flag & 0x10)
        forward = (not neg) if f.rna[src] else None
        start_pos, end_pos = breaks[0], breaks[-1]
        m = f.min_first_last_exon_length
        if m > 0 and len(breaks) > 2:  # :205-228
            sze = len(breaks)
            diff0, diff1 = breaks[1] - start_pos, end_pos - breaks[-2]
            if sze == 4 and diff0 < m and diff1 < m:
                if diff1 < diff0:
                    del breaks[sze - 2:]
                else:
                    del breaks[:2]
            else:
                if diff1 < m:
                    del breaks[sze - 2:]
                if diff0 < m:
                    del breaks[:2]
            end_pos, start_pos = breaks[-1], breaks[0]
        nb = len(breaks)
        a = self.annot
        leader = (nb > 1 and a.is_leader(breaks[1])) if f.coronavirus else False
        key = ""
        if f.coronavirus:  # :266-292
            if f.include_start or nb == 2:
                key += a.next_upstream(start_pos, forward) + "_"
            first = True
            for i in range(1, nb - 1, 2):
                gap = breaks[i + 1] - breaks[i]
                key += a.next_upstream(breaks[i], forward) + ","
                key += a.next_downstream(breaks[i + 1], forward) + "_"
                if gap > T:
                    if self.calc_breaks1:
                        k = (src, 0 if first else 1, breaks[i], breaks[i + 1])
                        self.pairs[k] = self.pairs.get(k, 0) + 1
                    if first:
                        leader = True
                    first = False
            if f.include_start or nb == 2:
                key += a.next_upstream(breaks[-1], forward)
        elif a.gff:  # :300-323
            genes_in, coords = [], []
            fwd_or_null = forward is None or forward
            for i in range(0, nb - 1, 2):
                upst = a.gff_next_upstream(breaks[i], breaks[i + 1], forward)
                if upst is not None:
                    genes_in.append(a.genes[upst])
                else:
                    coords.append(f"{a.chrom_index}.{jround(breaks[i + 1] if fwd_or_null else breaks[i], f.bin)}")
            if genes_in:
                order = java_hashset_order(genes_in)
                if order is None:  # a real tree bin: not restated, treated like a record the reference throws on
                    self.rows.append(dict(ok=False))
                    return
                key += "".join(g + "_" for g in order)
            else:
                key += coords[-1] if fwd_or_null else coords[0]
        else:  # EmptyAnnotation branch :293-299
            if forward is not None and not forward:
                key += a.next_upstream(breaks[0], forward)
            else:
                key += a.next_upstream(breaks[-1], forward)
        if f.enforce_strand:
            key += "+" if forward else "-"
        type_ind = a.type_ind(start_pos, end_pos) if f.coronavirus else 0
        if f.record_depth:  # setStartEnd :332
            ms.append(start_pos)
            me.append(end_pos)
        if f.coronavirus:  # :349-356
            st1 = breaks[-2]
            for i in range(len(a.start) - 1, -1, -1):
                if a.start[i] < st1 - 10:
                    break
                if end_pos > a.start[i] + 100:
                    (a.spliced if start_pos <= f.start_thresh else a.unspliced)[src][i] += 1
        # ---- commit → matchCluster (J/cluster/CigarClusters.java:75-102), copy-ctor / merge (CigarCluster.java:366-405, 633-698)
        self.total_counts[src] += 1
        lo, hi = 0, nb
        if not f.include_start and forward is not None:  # cloneBreaks :620-631
            if forward:
                lo = 1
            else:
                hi = nb - 1
        br = tuple(jround(x, f.bin) for x in breaks[lo:hi])
        cl = self.clusters.get(key)
        if cl is None:
            cl = Cluster(id=f"ID{self.chrom_index}.{len(self.clusters)}", key=key, breaks=list(br), read_count=[0] * f.num_sources)
            self.clusters[key] = cl
        fit = f.first_is_transcriptome
        sub = cl.subs.get(br)
        if sub is None:
            sid = read_name if (fit and src == 0) else f"{cl.id}.t{len(cl.subs)}"   # CigarClusters.java:85 / CigarCluster.java:657-661
            sub = dict(id=sid, count=[0] * f.num_sources, true_breaks=None, break_sum=0, forward=forward, neg=neg)
            cl.subs[br] = sub
        elif fit and src == 0:
            sub["id"] = read_name                                                      # CigarCluster.java:667-669
        sub["count"][src] += 1
        if f.write_gff_or_isoforms and (not fit or src == 0 or sub["count"][0] == 0):  # Count.addBreaks :84-96
            sub["break_sum"] += 1
            if sub["true_breaks"] is None:
                sub["true_breaks"] = [0.0] * nb
            if len(sub["true_breaks"]) == nb:
                for i in range(nb):
                    sub["true_breaks"][i] += breaks[i] / 1e3
        cl.read_count[src] += 1
        if f.record_depth:
            for p, mt in events:
                cl.depth[(src, p)] = cl.depth.get((src, p), 0) + 1
                if not mt:
                    cl.errors[(src, p)] = cl.errors.get((src, p), 0) + 1
            for p in ms:
                cl.map_start[(src, p)] = cl.map_start.get((src, p), 0) + 1
            for p in me:
                cl.map_end[(src, p)] = cl.map_end.get((src, p), 0) + 1
        self.rows.append(dict(ok=True, cluster=cl.id, sub=sub["id"], start_pos=start_pos, end_pos=end_pos, read_st=st_r, read_en=end_r,
                              type_ind=type_ind, leader=leader, neg=neg, breaks=list(breaks), key=key,
                              maps_sum=len(events), err_sum=sum(1 for _, mt in events if not mt)))


def unpack_batch(b):
    """synth.Batch → list of (pos, flag, src, [(len, op)], [base codes])."""
    out = []
    for i in range(b.n):
        cg = b.cigar[int(b.cigar_off[i]):int(b.cigar_off[i + 1])]
        cig = [(int(w) >> 4, int(w) & 15) for w in cg]
        sq = b.seq4[int(b.seq_off[i]):int(b.seq_off[i + 1])]
        bases = []
        for byte in sq:
            bases.append(int(byte) >> 4)
            bases.append(int(byte) & 15)
        out.append((int(b.pos[i]), int(b.flag[i]), int(b.src[i]), cig, bases))
    return out
