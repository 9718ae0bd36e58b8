"""Named workloads: genome shape, ORF table and read-model parameters of the BASELINE.json configs.

The ORF tables restate /root/reference/data/SARS-Cov2/VIC01/Coordinates.csv:1-13 and
/root/reference/data/229E_CoV/Coordinates.csv:1-10 (gene, Minimum, Maximum; all rows Direction=forward).
The genome sequence itself is synthetic (uniform ACGT + polyA tail, seeded): reads are generated against it, so
only its length and tail matter to the path."""
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Annotation:
    """Host-side mirror of J/cluster/Annotation.java:138-179 (CSV ORF table) / EmptyAnnotation."""
    kind: int  # abi.NPT_ANNOT_*
    genes: List[str] = field(default_factory=list)
    start: List[int] = field(default_factory=list)
    end: List[int] = field(default_factory=list)
    strand: List[bool] = field(default_factory=list)

    @property
    def name_id(self):
        first = {}
        return [first.setdefault(g, i) for i, g in enumerate(self.genes)]

    @classmethod
    def from_csv(cls, path, enforce_strand=False):
        """Parse a Coordinates.csv exactly as Annotation.java:138-179 does (header-located columns, 'both' rows duplicated)."""
        a = cls(kind=0)
        with open(path) as f:
            header = f.readline().rstrip("\n").split(",")
            gi, si, ei, di = (header.index(k) for k in ("gene", "Minimum", "Maximum", "Direction"))
            for line in f:
                line = line.rstrip("\n")
                if not line:
                    continue
                t = line.split(",")
                both, fwd = t[di] == "both", t[di] == "forward"
                a.genes.append(t[gi]); a.start.append(int(t[si])); a.end.append(int(t[ei]))
                if both and enforce_strand:
                    a.strand += [True, False]
                    a.genes.append(t[gi]); a.start.append(int(t[si])); a.end.append(int(t[ei]))
                else:
                    a.strand.append(fwd)
        return a

    @classmethod
    def empty(cls):
        return cls(kind=1)

    def adjust3utr(self, seqlen2):
        """Annotation.adjust3UTR (J/cluster/Annotation.java:205-212): the last row's end is pulled back to seqlen2 =
        reference length - polyA tail when it lies beyond it.  The reference does this on the first read it processes in
        coronavirus mode (IdentityProfile1.java:179-183 with polyAlen from :649); only the End column of 0.annot changes."""
        if self.kind == 0 and self.end and self.end[-1] > seqlen2 and self.start[-1] < seqlen2:
            self.end[-1] = seqlen2
        return self

    @classmethod
    def gff(cls, genes, start, end, strand):
        """Exon rows of one chromosome in file order (GFFAnnotation.java:409-428): gene = the exon's parent attribute."""
        return cls(kind=2, genes=list(genes), start=list(start), end=list(end), strand=list(strand))

    @classmethod
    def from_gff_lines(cls, lines, chrom, gtf=False, features=("Name", "description", "ID", "biotype", "Parent")):
        """Exon rows of one chromosome from GFF3 / GTF text, as GFFAnnotation's constructor reads them
        (J/cluster/GFFAnnotation.java:368-459): feature types ending in "exon", gene = the value of the attribute named by
        the 5th --GFF_features entry ("Parent" by default) after GFFAnnotation.process (:337-362: GFF values keep the part
        after the last ':', GTF values lose their quotes).  The reference takes one zip entry per chromosome and tries the
        name with and without a "chr" prefix (:370-372); here rows are selected by their first column under the same rule."""
        name, description, id_, biotype, parent = features
        headers = [id_, name, description, biotype, parent]  # GFFAnnotation.java:387
        split = " " if gtf else "="
        seqids = {chrom}
        seqids.add(chrom[3:] if chrom.startswith("chr") else "chr" + chrom)
        a = cls(kind=2)
        for st in lines:
            st = st.rstrip("\n")
            if not st or st.startswith("#"):
                continue
            t = st.split("\t")
            typ = t[2]
            if typ == "chromosome" or typ.endswith("biological_region") or typ.startswith("unknown") or "scaffold" in typ or "contig" in typ:
                continue
            if t[0] not in seqids or not typ.endswith("exon"):
                continue
            target = [""] * 5
            for d in t[8].split(";"):
                dj = d.strip().split(split)
                if len(dj) < 2:
                    continue
                key = dj[0].strip()
                val = dj[1][1:-1] if gtf else dj[1].split(":")[-1]
                for k, h in enumerate(headers):
                    if key == h:
                        target[k] = dj[1].split("[")[0] if k == 2 else val
            a.genes.append(target[4]); a.start.append(int(t[3])); a.end.append(int(t[4])); a.strand.append(t[6][0] == "+")
        return a

    @property
    def name_hash(self):
        """h ^ (h >>> 16) of Java's String.hashCode() per row (HashMap.hash): decides HashSet<String> iteration order."""
        out = []
        for g in self.genes:
            h = 0
            for ch in g:
                h = (31 * h + ord(ch)) & 0xFFFFFFFF
            out.append((h ^ (h >> 16)) & 0xFFFFFFFF)
        return out


def _csv(rows):
    return Annotation(kind=0, genes=[r[0] for r in rows], start=[r[1] for r in rows], end=[r[2] for r in rows],
                      strand=[True] * len(rows))


SARS2_VIC01 = _csv([
    ("leader", 1, 75), ("ORF1ab", 266, 21555), ("S", 21563, 25384), ("ORF3a", 25393, 26220), ("E", 26245, 26472),
    ("M", 26523, 27191), ("ORF6", 27202, 27387), ("ORF7a", 27394, 27759), ("ORF8", 27894, 28259), ("N", 28274, 29533),
    ("ORF10", 29558, 29674), ("3UTR", 29675, 29893)])

COV_229E = _csv([
    ("leader", 1, 100), ("ORF1ab", 293, 20568), ("S", 20570, 24091), ("ORF4a", 24091, 24492), ("ORF4b", 24482, 24748),
    ("E", 24750, 24983), ("M", 24995, 25672), ("N", 25686, 26855), ("3UTR", 26856, 27317)])


def read_fasta(path):
    """First record of a (gzipped) FASTA file: (name, description, sequence).  The reference reads its --reference with
    japsa SequenceReader.readAll (ViralTranscriptAnalysisCmd2.java:594); name = header up to the first blank."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    name, desc, seq = None, "", []
    with op(path, "rt") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    break
                head = line[1:].split(None, 1)
                name, desc = head[0], (head[1] if len(head) > 1 else "")
            elif name is not None:
                seq.append(line)
    return name, desc, "".join(seq)


def polya_len(seq: str) -> int:
    """TranscriptUtils.polyAlen (J/cluster/TranscriptUtils.java:116-132): 0 unless the last 10 bases are all 'A', else the
    length of the trailing run of 'A'."""
    if len(seq) < 10 or any(ch != "A" for ch in seq[-10:]):
        return 0
    i = 1
    while i <= len(seq) and seq[len(seq) - i] == "A":
        i += 1
    return i - 1


@dataclass
class Workload:
    name: str
    genome_len: int
    polya_len: int
    annotation: Annotation
    donor: int
    acceptors: List[int]
    acceptor_w: List[float]
    genome_file: str = ""          # FASTA of the real reference; empty = seeded random genome + polyA tail
    genome_seed: int = 20200101
    w_mix: tuple = (0.60, 0.25, 0.10, 0.05)  # canonical, leaderless, genomic, double-junction
    p_mismatch: float = 0.04
    p_ins: float = 0.03
    p_del: float = 0.03
    indel_mean: float = 1.5
    p_neg: float = 0.02
    leaderless_len: float = 600.0
    genomic_len: float = 1000.0


# TRS-B acceptor sites: N 28260 (the KAT of SURVEY §8c), the others ~14 nt upstream of each ORF start
# (fimo TRS hits /root/reference/data/SARS-Cov2/VIC01/fimo.tsv: 21556, 25385, 26237 ...).
SARS2 = Workload(
    name="sars2_vic01_drna", genome_len=29893, polya_len=33, annotation=SARS2_VIC01, donor=69,
    acceptors=[28260, 26473, 25385, 27388, 27888, 21556, 26237, 27041, 29530],
    acceptor_w=[0.65, 0.08, 0.03, 0.06, 0.07, 0.01, 0.03, 0.03, 0.04])

HCOV229E = Workload(
    name="hcov229e_drna", genome_len=27317, polya_len=20, annotation=COV_229E, donor=65,
    acceptors=[25672, 24983, 24740, 24480, 24080, 20560],
    acceptor_w=[0.40, 0.20, 0.10, 0.10, 0.10, 0.10])


def from_files(name, fasta_path, csv_path, donor, acceptors, acceptor_w, enforce_strand=False, **kw) -> Workload:
    """A workload on a real reference + Coordinates.csv (BASELINE.json configs[0]: run_example.sh's VIC01 files): genome
    length and polyA tail from the FASTA, ORF table through the reference's own parser rules (Annotation.from_csv)."""
    _n, _d, seq = read_fasta(fasta_path)
    return Workload(name=name, genome_len=len(seq), polya_len=polya_len(seq), annotation=Annotation.from_csv(csv_path, enforce_strand),
                    donor=donor, acceptors=list(acceptors), acceptor_w=list(acceptor_w), genome_file=str(fasta_path), **kw)


def golden_path(*parts):
    """tests/golden/<...>: copies of the reference's own data files (tests/golden/make_reference_fixtures.py)."""
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", *parts)


def sars2_vic01_real() -> Workload:
    """configs[0] as specified: MT007544.1 (29 893 bp, 33-A tail) + data/SARS-Cov2/VIC01/Coordinates.csv."""
    return from_files("sars2_vic01_real", golden_path("vic01", "wuhan_coronavirus_australia.fasta.gz"), golden_path("vic01", "Coordinates.csv"),
                      donor=SARS2.donor, acceptors=SARS2.acceptors, acceptor_w=SARS2.acceptor_w)


def cov229e_real() -> Workload:
    """configs[2]: data/229E_CoV/WT_229_reference.fasta.gz (27 317 bp) + its Coordinates.csv."""
    return from_files("hcov229e_real", golden_path("cov229e", "WT_229_reference.fasta.gz"), golden_path("cov229e", "Coordinates.csv"),
                      donor=HCOV229E.donor, acceptors=HCOV229E.acceptors, acceptor_w=HCOV229E.acceptor_w)


def nc045512_real() -> Workload:
    """data/SARS-Cov2/NC: NC_045512.2 (29 903 bp, 33-A tail) + its Coordinates.csv, where ORF7b (27756-27887) overlaps ORF7a."""
    return from_files("sars2_nc045512_real", golden_path("nc045512", "NC_045512.fasta.gz"), golden_path("nc045512", "Coordinates.csv"),
                      donor=SARS2.donor, acceptors=[28270, 26483, 25395, 27398, 27766, 27898, 21566, 26247, 27051], acceptor_w=SARS2.acceptor_w)


def u13369_real() -> Workload:
    """data/U13369: the human ribosomal DNA repeat (42 999 bp, IUPAC codes in the sequence, no polyA tail) + its 8-row table
    (5'ETS, 18S, ITS1, 5.8S, ITS2, 28S, 3'ETS, repeat region); reads modelled as precursor fragments spliced between the genes."""
    return from_files("rdna_u13369_real", golden_path("u13369", "Human_ribosomal_DNA_complete_repeating_unit.fasta.gz"),
                      golden_path("u13369", "Coordinates.csv"), donor=3650, acceptors=[6623, 7935, 12970, 5528], acceptor_w=[0.4, 0.3, 0.2, 0.1])


def ref_char_to_code(seq: str) -> np.ndarray:
    """Reference letters → BAM 4-bit codes (=ACMGRSVTWYHKDBN), case-insensitive; anything else → N(15).
    This is the comparison domain of refSeq.getBase(..) == readSeq.getBase(..) (IdentityProfile1.java:534)."""
    table = np.full(256, 15, dtype=np.uint8)
    for i, ch in enumerate("=ACMGRSVTWYHKDBN"):
        table[ord(ch)] = i
        table[ord(ch.lower())] = i
    return table[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


# ---------------------------------------------------------------------------------------------- host mode (configs[4])
@dataclass
class HostChromosome:
    """One chromosome of the synthetic human-like gene model (no human reference / GTF ships with the
    reference): genes laid out left to right, GENCODE-like exon counts, exon and intron lengths; the exon table is what
    GFFAnnotation's constructor would hold for this chromosome (one row per exon line, gene = the exon's Parent)."""
    name: str
    index: int
    length: int
    seed: int
    gene_exon_off: np.ndarray = None   # int32 [n_genes + 1]
    exon_start: np.ndarray = None      # int32, 1-based inclusive
    exon_end: np.ndarray = None
    gene_strand: np.ndarray = None     # bool [n_genes]
    # read model (spliced long reads, 8-12 % error)
    p_skip: float = 0.12
    p_trunc5: float = 0.5
    p_intergenic: float = 0.03
    gene_skew: float = 3.0
    p_mismatch: float = 0.04
    p_ins: float = 0.03
    p_del: float = 0.03
    indel_mean: float = 1.5
    p_neg: float = 0.5

    @property
    def n_genes(self):
        return len(self.gene_exon_off) - 1

    @property
    def annotation(self) -> Annotation:
        per = np.diff(self.gene_exon_off)
        gid = np.repeat(np.arange(self.n_genes), per)
        genes = [f"G{self.index}x{g}" for g in gid]
        return Annotation.gff(genes, self.exon_start.tolist(), self.exon_end.tolist(), np.repeat(self.gene_strand, per).tolist())


def host_chromosome(index, length, seed=20200500, name=None) -> HostChromosome:
    rng = np.random.default_rng(seed + 7919 * index)
    starts, ends, off, strand = [], [], [0], []
    margin = 5000 if length > 200000 else 300
    p = margin
    while True:
        n_ex = int(min(40, 1 + rng.geometric(0.13)))
        ex_len = np.maximum(30, np.exp(np.log(140.0) + 0.7 * rng.standard_normal(n_ex))).astype(np.int64)
        in_len = np.maximum(80, np.exp(np.log(1500.0) + 1.1 * rng.standard_normal(n_ex))).astype(np.int64)
        span = int(ex_len.sum() + in_len[:-1].sum())
        if p + span + margin > length:
            if starts or span + 2 * margin > length:
                break
            continue   # a chromosome this short (chrM) gets at least one gene that fits
        q = p
        for k in range(n_ex):
            starts.append(q); ends.append(q + int(ex_len[k]) - 1)
            q += int(ex_len[k]) + int(in_len[k])
        off.append(len(starts)); strand.append(bool(rng.random() < 0.5))
        p = q + int(np.exp(np.log(8000.0) + 1.0 * rng.standard_normal()))
    c = HostChromosome(name=name or f"chr{index + 1}", index=index, length=length, seed=seed)
    c.gene_exon_off = np.asarray(off, np.int32); c.exon_start = np.asarray(starts, np.int32); c.exon_end = np.asarray(ends, np.int32)
    c.gene_strand = np.asarray(strand, bool)
    return c


# GRCh38-like relative lengths of chr1..22, X, Y, M (Mb), scaled by `scale` (1.0 = the real lengths)
_GRCH38_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57, 0.0166]


def human_like(scale=0.04, seed=20200500):
    """25 chromosomes; scale 1.0 gives ~2e5 genes / ~1.6e6 exons, the default 0.04 keeps genome generation and HBM use of a
    bench run small (chr1 = 9.9 Mb) while every per-chromosome code path (GFF exon search, reference out of shared memory,
    a new cluster table per chromosome) is the same."""
    return [host_chromosome(i, max(20000, int(mb * 1e6 * scale)), seed) for i, mb in enumerate(_GRCH38_MB)]
