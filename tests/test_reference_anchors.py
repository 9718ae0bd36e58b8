"""The few artefacts of this path that the reference itself holds, pinned:

  data/SARS-Cov2/VIC01/0.annot.tsv       an 0.annot output of a real run: Gene / Start / End columns (the counts come from
                                         an unknown BAM and cannot be reproduced), incl. the adjust3UTR end 29860
  data/SARS-Cov2/VIC01/Coordinates.csv   parsed by the rules of J/cluster/Annotation.java:138-179
  wuhan_coronavirus_australia.fasta.gz   MT007544.1: length 29 893, polyA tail 33 (TranscriptUtils.polyAlen :116-132)
  J/run/MergeCmd.java:46-47              one transcripts row of an old run: start 8, end 29858, type 5_3, startBreak 69,
                                         endBreak 26237, leftGene LEADER, rightGene E
  R/www/shiny-instructions.html          key grammar example leader_ORF1ab,S_3UTR style
The copies under tests/golden/ are made by tests/golden/make_reference_fixtures.py."""
import numpy as np

from nptranscript_b200 import render, synth, workloads
from oracle import oracle


def _ref_str(w):
    return workloads.read_fasta(w.genome_file)[2]


def test_fasta_and_polya():
    w = workloads.sars2_vic01_real()
    name, desc, seq = workloads.read_fasta(w.genome_file)
    assert name == "MT007544.1" and len(seq) == 29893 == w.genome_len
    assert workloads.polya_len(seq) == 33 == w.polya_len == workloads.SARS2.polya_len
    assert workloads.polya_len("ACGT" * 5 + "A" * 9) == 0 and workloads.polya_len("C" + "A" * 10) == 10
    w9 = workloads.cov229e_real()
    assert w9.genome_len == 27317 == workloads.HCOV229E.genome_len


def test_coordinates_csv_parser():
    w = workloads.sars2_vic01_real()
    a, h = w.annotation, workloads.SARS2_VIC01
    assert (a.genes, a.start, a.end, a.strand) == (h.genes, h.start, h.end, h.strand)
    a9, h9 = workloads.cov229e_real().annotation, workloads.COV_229E
    assert (a9.genes, a9.start, a9.end) == (h9.genes, h9.start, h9.end)


def test_direction_both_rows_are_duplicated(tmp_path):
    """Annotation.java:160-164: a Direction=both row appears twice (forward, reverse) when enforceStrand"""
    p = tmp_path / "c.csv"
    p.write_text("Name,Type,Minimum,Maximum,Length,Direction,gene\nx,gene,1,75,75,forward,leader\ny,gene,100,900,801,both,G2\nz,gene,950,990,41,reverse,G3\n")
    a = workloads.Annotation.from_csv(str(p), enforce_strand=True)
    assert (a.genes, a.start, a.strand) == (["leader", "G2", "G2", "G3"], [1, 100, 100, 950], [True, True, False, False])
    b = workloads.Annotation.from_csv(str(p), enforce_strand=False)
    assert (b.genes, b.strand) == (["leader", "G2", "G3"], [True, False, False])


def test_annot_rows_reproduce_the_reference_sample():
    w = workloads.sars2_vic01_real()
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200301, 2000, num_sources=2)
    r = oracle.run(oracle.make_config(bin=10, break_thresh=1000, num_sources=2), g, w.annotation, [b])
    rows = render.annot_rows(r, w.annotation, seqlen=w.genome_len, polya=w.polya_len)
    want = open(workloads.golden_path("vic01", "0.annot.tsv")).read().splitlines()
    assert rows[0] == want[0]                                   # header incl. Spliced_5_0 Unspliced_5_0 Spliced_5_1 Unspliced_5_1
    assert [x.split("\t")[:3] for x in rows[1:]] == [x.split("\t")[:3] for x in want[1:]]
    assert rows[-1].split("\t")[:3] == ["3UTR", "29675", "29860"]  # = 29893 - 33 (adjust3UTR)
    # without a processed read the reference never calls adjust3UTR
    empty = oracle.run(oracle.make_config(bin=10, num_sources=2), g, w.annotation, [b.slice(0, 0)])
    assert render.annot_rows(empty, w.annotation, seqlen=w.genome_len, polya=w.polya_len)[-1].split("\t")[2] == "29893"


def test_mergecmd_transcript_row():
    """start 8, startBreak 69, endBreak 26237, end 29858 on the real reference: LEADER -> E, type 5_3"""
    w = workloads.sars2_vic01_real()
    seq = _ref_str(w)
    bases = seq[8 - 1:69] + seq[26237 - 1:29858]
    b = synth.from_reads([(8, 0, 0, f"{69 - 8 + 1}M{26237 - 69 - 1}N{29858 - 26237 + 1}M", bases)])
    g = synth.genome_codes(w)
    r = oracle.run(oracle.make_config(bin=10, break_thresh=1000), g, w.annotation, [b])
    assert r["breaks"].tolist() == [8, 69, 26237, 29858]
    assert render.cluster_keys(r, w.annotation, 0)[0] == "leader_leader,E_3UTR"
    row = render.read_to_cluster_rows(r, ["r0"], w.annotation, "MT007544.1", 0, True)[0].split("\t")
    assert row[7] == "5_3" and row[9:11] == ["8", "29858"] and row[13] == "1"   # type_nme, startPos/endPos, leader_break
    assert r["pair_start"].tolist() == [69] and r["pair_end"].tolist() == [26237] and r["pair_level"].tolist() == [0]


def test_config1_on_the_real_files_two_restatements_agree():
    """configs[0] shape on the real reference + Coordinates.csv (run_example.sh flags + --RNA=true --writeGFF=true
    --maxThreads=1): oracle.cpp and py_oracle.py agree on the first 1 500 of the seed-20200301 reads, strings included"""
    from tests.test_oracle_crosscheck import _compare
    w = workloads.sars2_vic01_real()
    b = synth.generate(w, 20200301, 1500)
    _compare(dict(bin=10, break_thresh=1000, record_depth=1), w.annotation, [b], w=w)
    _compare(dict(bin=10, break_thresh=1000, record_depth=0), w.annotation, [b.slice(0, 300)], w=w)


def test_other_reference_pairs_the_reference_ships():
    """data/SARS-Cov2/NC (ORF7b overlapping ORF7a, 3'UTR row ending at 29 903 = 33 bases past seqlen - polyA) and data/U13369
    (43 kb rDNA repeat, IUPAC codes in the reference, no leader row): parsed by the reference's rules, the two restatements
    agree on reads generated against them, and adjust3UTR moves NC's last row like VIC01's"""
    from tests.test_oracle_crosscheck import _compare
    nc = workloads.nc045512_real()
    assert nc.genome_len == 29903 and nc.polya_len == 33
    a = nc.annotation
    assert a.genes[7:9] == ["ORF7a", "ORF7b"] and a.start[8] < a.end[7] and a.end[-1] == 29903
    b = synth.generate(nc, 20200301, 1200)
    _compare(dict(bin=10, break_thresh=1000, record_depth=1), a, [b], w=nc)
    res = oracle.run(oracle.make_config(bin=10, break_thresh=1000), synth.genome_codes(nc), a, [b])
    rows = render.annot_rows(res, a, seqlen=nc.genome_len, polya=nc.polya_len)
    assert rows[-1].split("\t")[:3] == ["3UTR", "29675", "29870"] and rows[9].split("\t")[0] == "ORF7b"
    u = workloads.u13369_real()
    assert u.genome_len == 42999 and u.polya_len == 0 and u.annotation.genes[0] == "5ETS" and len(u.annotation.genes) == 8
    g = synth.genome_codes(u)
    assert set(np.unique(g).tolist()) - {1, 2, 4, 8}          # R / Y / N in the reference sequence
    bu = synth.generate(u, 5, 1200, num_sources=2)
    _compare(dict(bin=100, break_thresh=500, record_depth=1, num_sources=2), u.annotation, [bu], w=u)
