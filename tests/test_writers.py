"""Cluster-level writers (nptranscript_b200/writers.py): Java map iteration orders against hand-derived cases and the oracle's
independent HashSet restatement; rows of 0.genes / 0.transcripts / 0.transcripts.fc built from oracle results (the same
dict layout the device returns) against the invariants CigarClusters.process1 implies; files written and read back."""
import gzip
import os

import numpy as np
import pytest

from nptranscript_b200 import engine, synth, workloads, writers
from oracle import oracle, py_oracle


def test_hashmap_order_matches_the_oracles_hashset_restatement():
    rng = np.random.default_rng(3)
    for trial in range(30):
        n = int(rng.integers(1, 200))
        names = list(dict.fromkeys("".join(chr(int(c)) for c in rng.integers(65, 91, int(rng.integers(1, 4)))) for _ in range(n)))
        want = py_oracle.java_hashset_order(names)
        if want is None:
            continue
        got = [names[i] for i in writers.hashmap_order([writers.java_string_hash(x) for x in names])]
        assert got == want


def test_map_orders_hand_derived():
    # three keys in bucket 1 of a 16-table: A (hash 1), B (33), C (17); nine fillers make 12 entries
    hashes = [1, 33, 17] + list(range(2, 11))
    # HashMap: no resize before the 13th entry -> bucket order, chains in insertion order
    assert writers.hashmap_order(hashes) == [0, 1, 2] + list(range(3, 12))
    # ConcurrentHashMap: the 12th put reaches sizeCtl = 12 -> transfer to 32.  Bucket 1 = A, B (bit 16 clear), C (set): lastRun = C,
    # A and B are PREPENDED to the low list -> B, A; C moves to bucket 17
    assert writers.concurrent_hashmap_order(hashes) == [1, 0] + list(range(3, 12)) + [2]
    # a 13th entry doubles the HashMap too; its split keeps the chain order: A, B stay, C moves
    assert writers.hashmap_order(hashes + [11]) == [0, 1] + list(range(3, 13)) + [2]
    # nine keys in one bucket of a small table: HashMap doubles once per append from the 9th on, ConcurrentHashMap presizes 16 -> 128
    same = [5 + 1024 * k for k in range(9)]
    assert sorted(writers.hashmap_order(same)) == list(range(9))
    assert sorted(writers.concurrent_hashmap_order(same)) == list(range(9))
    with pytest.raises(writers.TreeBinNotRestated):
        writers.hashmap_order([7] * 11)   # equal hashCodes ("Aa" / "BB" ...): the 9th and 10th double the table to 64, the 11th treeifies
    assert writers.java_string_hash("leader_N") == (sum(ord(c) * 31 ** (7 - i) for i, c in enumerate("leader_N")) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert writers.java_list_hash([0, 282, 298]) == ((((31 + 0) * 31 + 282) * 31 + 298) + 2 ** 31) % 2 ** 32 - 2 ** 31


def _report(n=4000, S=2, depth=0, fit=0):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 11, n, num_sources=S)
    kw = dict(bin=100, break_thresh=1000, record_depth=depth, num_sources=S, track_break_sums=1, first_is_transcriptome=fit)
    res = oracle.run(oracle.make_config(**kw), g, w.annotation, [b])
    names = [f"read{i}" for i in range(n)]
    rep = writers.ChromosomeReport(res, w.annotation, "MT007544.1", 0, engine.make_config(**kw), w.genome_len, src=b.src, read_names=names,
                                   source_names=[f"s{i}.bam" for i in range(S)])
    return rep, res, b


def test_rows_follow_process1():
    rep, res, b = _report()
    genes, trans, fc = rep.rows()
    nc, ns = res["n_clusters"], res["n_subs"]
    assert len(genes) == nc == len(fc) and len(trans) == ns
    gcols = [x.split("\t") for x in genes]
    assert all(len(x) == 13 + rep.S for x in gcols)
    # coronavirus mode: clusters by descending read count (stable), every id once
    tot = [int(x[12]) for x in gcols]
    assert tot == sorted(tot, reverse=True) and sum(tot) == int((res["cluster_id"] >= 0).sum())
    assert sorted(x[0] for x in gcols) == sorted(f"ID0.{c}" for c in range(nc))
    assert all(x[11] == "-1" and x[9] == "-" and x[10] == "0" for x in gcols)      # totLen before writeDepthH5, empty span
    keys = rep.keys
    assert [x[8] for x in gcols] == [keys[int(x[0].split(".")[1])] for x in gcols]
    # the most frequent cluster is the canonical leader -> N junction and has a leader break
    top = gcols[0]
    assert top[8].startswith("leader_leader,N_") and top[7] == "1" and top[4] == "5_3"
    tcols = [x.split("\t") for x in trans]
    assert all(len(x) == 13 + rep.S and x[0] == x[9] and x[6] == "1" and x[11] == "NA" for x in tcols)
    assert sum(int(x[12]) for x in tcols) == sum(tot)
    # transcripts of a cluster are contiguous and in the order of the cluster rows
    owner = [x[0].rsplit(".t", 1)[0] for x in tcols]
    assert [o for i, o in enumerate(owner) if i == 0 or owner[i - 1] != o] == [x[0] for x in gcols]
    # consensus start <= end, exon count = round(breaks / 2)
    assert all(int(x[2]) <= int(x[3]) for x in tcols)
    # feature counts: Start/End lists pair up, Length = sum of block lengths of the best-supported transcript
    for x in (y.split("\t") for y in fc):
        st, en = [int(v) for v in x[2].split(";")], [int(v) for v in x[3].split(";")]
        assert len(st) == len(en) == len(x[1].split(";")) and int(x[5]) == sum(e - s for s, e in zip(st, en))


def test_first_is_transcriptome_names_and_files(tmp_path):
    rep, res, b = _report(n=1500, depth=1, fit=1)
    names = rep.sub_names_final()
    so = res["cl_sub_off"].astype(np.int64)
    g = so[np.maximum(res["cluster_id"], 0)] + res["sub_id"]
    for s in range(res["n_subs"]):
        hits = np.flatnonzero((res["cluster_id"] >= 0) & (g == s) & (b.src == 0))
        assert names[s] == (f"read{hits[-1]}" if len(hits) else names[s]) and (len(hits) > 0 or ".t" in names[s])
    paths = rep.write(str(tmp_path), polya=33)
    base = {os.path.basename(p) for p in paths}
    assert {"0transcripts.txt.gz", "0genes.txt.gz", "0transcripts.fc.txt.gz", "0annot.txt.gz", "0readToCluster.txt.gz"} <= base
    assert any(x.startswith("0.breakpoints.s0.bam.0") for x in base) and any(x.startswith("0.clusters.depth.") for x in base)
    with gzip.open(tmp_path / "0genes.txt.gz", "rt") as f:
        lines = f.read().splitlines()
    assert lines[0] == "#s0.bam\ts1.bam\t" and lines[1].startswith("ID\tchrom\tstart\tend\ttype_nme") and len(lines) == 2 + res["n_clusters"]
    with gzip.open(tmp_path / "0annot.txt.gz", "rt") as f:
        assert f.read().splitlines()[-1].split("\t")[:3] == ["3UTR", "29675", "29860"]     # the reference-held 0.annot.tsv row
    with gzip.open(tmp_path / "0readToCluster.txt.gz", "rt") as f:
        assert len(f.read().splitlines()) == 1 + int((res["read_flags"] & 0x10 == 0).sum())


def test_gff_span_hand_cases():
    A = workloads.Annotation.gff(["g1", "g1", "g2", "g3"], [100, 300, 290, 1000], [200, 400, 420, 1100], [True, True, True, False])
    # both breaks of the first block lie in g1's first exon, the second block's in g1's second exon and in g2: g1 alone explains all
    assert writers.gff_span([110, 190, 305, 395], True, A) == ("g1", [0, 1, 2], ["g1"])
    # the last break only lies in g2's exon (401..420 is outside g1's 300..400 + 5): g1 (3 breaks) first, then g2 to cover the 4th
    assert writers.gff_span([110, 190, 305, 415], True, A) == ("g1;g2", [0, 1, 2], ["g1", "g2"])
    # no exon row between the first one that ends after the read's start and the last one that starts before its end -> "."
    assert writers.gff_span([500, 600], True, A) == (".", [], [])
    assert writers.gff_span([2000, 2100], True, A) == (".", [], [])
    # rows out of position order (file order is what counts): the window exists, nothing overlaps -> empty string, not "."
    B = workloads.Annotation.gff(["g9", "g1"], [900, 100], [1000, 200], [True, True])
    assert writers.gff_span([500, 600], True, B) == ("", [], [])
    # enforceStrand: only rows of the read's strand take part
    assert writers.gff_span([1005, 1095], True, A, enforce_strand=True) == (".", [], [])
    assert writers.gff_span([1005, 1095], False, A, enforce_strand=True) == ("g3", [3], ["g3"])


def test_report_in_gff_mode():
    c = workloads.human_like()[22]
    g = synth.host_genome_codes(c)
    b = synth.generate_host(c, g, 5, 1500, num_sources=2, transcriptome=True)
    kw = dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, first_is_transcriptome=1, track_break_sums=1, num_sources=2, rna=[1, 1])
    res = oracle.run(oracle.make_config(**kw), g, c.annotation, [b], c.index)
    rep = writers.ChromosomeReport(res, c.annotation, c.name, c.index, engine.make_config(**kw), c.length, src=b.src, read_names=[f"r{i}" for i in range(b.n)])
    genes, trans, fc = rep.rows()
    assert len(genes) == res["n_clusters"] and len(trans) == res["n_subs"]
    # host mode: clusters in ConcurrentHashMap order (no sort by count); spliced reads name their gene
    named = [x.split("\t") for x in genes if x.split("\t")[9] != "-"]
    assert len(named) > 0.5 * len(genes) and all(int(x[10]) >= 1 for x in named)
    assert sum(1 for s, n in rep.read_span if n > 0) > 0.8 * b.n
    # source-0 reads are the transcripts themselves: their sub-clusters carry a read name as id
    assert any(not x.split("\t")[0].startswith("ID") for x in trans)


def test_gff_and_bed_lines(tmp_path):
    rep, res, b = _report(n=3000, S=2, depth=0, fit=1)
    gff, bed = rep.gff_bed(gff_thresh=(2, 1), max_transcripts=3)
    assert len(gff) == 2 and len(bed) == 2
    for lines in gff:
        cols = [x.split("\t") for x in lines]
        assert all(len(c) == 9 and c[1] == "np" and c[2] in ("transcript", "exon", "gene") and int(c[3]) <= int(c[4]) for c in cols)
        # every transcript is followed by its exons, every cluster's block ends with its gene line spanning its transcripts
        genes = [c for c in cols if c[2] == "gene"]
        assert genes and all("Parent=" not in c[8] and c[8].startswith("ID=ID0.") for c in genes)
        tr = [c for c in cols if c[2] == "transcript"]
        assert tr and all(";Parent=ID0." in c[8] for c in tr)
        per_gene = {}
        for c in tr:
            per_gene.setdefault(c[8].split(";Parent=")[1].split(";")[0], []).append((int(c[3]), int(c[4])))
        for g in genes:
            span = per_gene[g[8].split(";")[0][3:]]
            assert int(g[3]) == min(a for a, _ in span) and int(g[4]) == max(e for _, e in span) and len(span) <= 3
    # writer 0 (the transcriptome's) only lists transcripts seen in source 0 and in some other source
    for c in (x.split("\t") for x in gff[0]):
        if c[2] == "transcript":
            cnt = [int(v) for v in c[8].split("count=")[1].split(",")]
            assert cnt[0] >= 1 and cnt[1] >= 1
    for k, lines in enumerate(bed):
        for c in (x.split("\t") for x in lines):
            assert len(c) == 12 and int(c[4]) > 0 and int(c[9]) == len(c[10].split(",")) == len(c[11].split(",")) and c[11].split(",")[0] == "0"
            assert int(c[2]) - int(c[1]) == int(c[11].split(",")[-1]) + int(c[10].split(",")[-1])   # last block ends at the transcript's end
    paths = rep.write(str(tmp_path), gff_thresh=(2, 1), max_transcripts=3, npy=False)
    assert {"0.gff.gz", "1.gff.gz", "00.bed.gz", "01.bed.gz"} <= {os.path.basename(p) for p in paths}


def test_rows_known_answer_derived_by_hand_from_the_java_lines():
    """KAT1 of SURVEY §8c (one read 10S65M28190N1600M30S at position 5 on the all-A reference, --bin 10): every field of the
    three rows worked out by hand from CigarClusters.process1 (:138-196), CigarCluster.exonCount / numIsoforms (:702-735),
    Annotation.getTypeInd (:233-236) and isLeader (:214-216)."""
    w = workloads.SARS2
    g = np.full(w.genome_len, 1, np.uint8)
    b = synth.from_reads([(5, 0, 0, "10S65M28190N1600M30S", "A" * 1705)])
    kw = dict(bin=10, break_thresh=1000, record_depth=0, track_break_sums=1)
    res = oracle.run(oracle.make_config(**kw), g, w.annotation, [b])
    rep = writers.ChromosomeReport(res, w.annotation, "MT007544.1", 0, engine.make_config(**kw), w.genome_len, src=b.src, read_names=["r0"],
                                   source_names=["a.bam"])
    genes, trans, fc = rep.rows()
    key = "leader_leader,N_3UTR"           # up(5) _ up(69),down(28260) _ up(29859)   (IdentityProfile1.java:266-292)
    # cnt.id, chrom, breaks[0], breaks[last], geneNme ("-": empty span), round(4/2), "1", leader break (6*10 < 266-10), h2 = breaks/10,
    # cnt.id, geneNames.size(), "NA", cnt.sum(), count0
    assert trans == ["\t".join(["ID0.0.t0", "MT007544.1", "5", "29859", "-", "2", "1", "1", "0,6,2826,2985", "ID0.0.t0", "0", "NA", "1", "1"])]
    # cc.id, chrom, startPos, endPos, type (5 <= 100 and 29859 >= 29893-100), exonCount "2", numIsoforms round((1-2)/2) = 0, leader break,
    # secondKey, geneNme, 0, totLen -1 (printed before writeDepthH5 sets it), readCountSum, count0
    assert genes == ["\t".join(["ID0.0", "MT007544.1", "5", "29859", "5_3", "2", "0", "1", key, "-", "0", "-1", "1", "1"])]
    # secondKey, chrom per block, starts, ends, "+;+", (69-5)+(29859-28260), count0
    assert fc == ["\t".join([key, "MT007544.1;MT007544.1", "5;28260", "69;29859", "+;+", "1663", "1"])]
    gff, bed = rep.gff_bed(gff_thresh=(1,), max_transcripts=1)
    assert gff[0][0] == "MT007544.1\tnp\ttranscript\t5\t29859\t.\t+\t.\tID=ID0.0.t0;Parent=ID0.0;gene_id=ID0.0;gene_type=5_3;type=ORF;gene_name=" + key + ";count=1"
    assert gff[0][1].startswith("MT007544.1\tnp\texon\t5\t69\t.\t+\t.\tID=ID0.0.t0.e.0;Parent=ID0.0.t0;gene_id=ID0.0.t0;")
    assert gff[0][3] == "MT007544.1\tnp\tgene\t5\t29859\t.\t+\t.\tID=ID0.0;gene_id=ID0.0;gene_type=5_3;type=ORF;gene_name=" + key + ";count=1"
    # chrom, start-1, end-1, transcript.gene, depth, strand, thick start/end, colour of source 0, floor(4/2), block sizes, block starts
    assert bed[0] == ["MT007544.1\t4\t29858\tID0.0.t0.ID0.0.t0\t1\t+\t4\t29858\t0,0,0\t2\t64,1599\t0,28255"]
