"""Host-side rendering helpers (nptranscript_b200/render.py) on oracle output: Java %5.3g, matrix layouts."""
import numpy as np

from nptranscript_b200 import render, synth, workloads
from oracle import oracle


def test_java_g53():
    cases = {0.0: "0.00", 0.1: "0.100", 0.123456: "0.123", 1.0: "1.00", 12.345: "12.3", 999.4: "999", 999.6: "1.00e+03", 0.00012345: "0.000123",
             0.000012345: "1.23e-05", 500.0: "500", 12345.0: "1.23e+04", 0.0625: "0.0625", 0.04: "0.0400"}
    for x, s in cases.items():
        assert render.java_g53(x) == s, (x, render.java_g53(x), s)


def test_matrices_and_rows():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 21, 4000)
    r = oracle.run(oracle.make_config(bin=10, break_thresh=1000, record_depth=1), g, w.annotation, [b])
    mats = render.breakpoint_matrices(r, 1)
    m0 = mats[(0, 0)]
    assert m0[2:, 2:].sum() == r["pair_count"][r["pair_level"] == 0].sum()
    assert np.array_equal(m0[1, 2:], m0[2:, 2:].sum(axis=0)) and np.array_equal(m0[2:, 1], m0[2:, 2:].sum(axis=1))
    assert np.all(np.diff(m0[0, 2:]) > 0) and np.all(np.diff(m0[2:, 0]) > 0)
    big = int(np.argmax(r["cl_read_count"].sum(axis=1)))
    dm = render.depth_matrix(r, big)
    assert np.all(dm[:, 1] > 0) and np.all(dm[:, 2] <= dm[:, 1]) and np.all(np.diff(dm[:, 0]) > 0)
    assert dm[:, 1].sum() == r["maps_sum"][r["cluster_id"] == big].sum()
    rows = render.read_to_cluster_rows(r, [f"read{i}" for i in range(r["n_reads"])], w.annotation, "MT007544.1", 0, True)
    assert len(rows) == r["n_reads"] and all(len(x.split("\t")) == 21 for x in rows)
    keys = render.cluster_keys(r, w.annotation, 0)
    assert rows[0].split("\t")[15] == keys[int(r["cluster_id"][0])]
    assert render.annot_rows(r, w.annotation)[1].split("\t")[0] == "leader"


def test_gff_and_gtf_exon_tables():
    """GFFAnnotation's constructor in small (J/cluster/GFFAnnotation.java:368-459, process :337-362)."""
    from nptranscript_b200 import workloads
    gff = """##gff-version 3
1\tensembl\tchromosome\t1\t5000\t.\t.\t.\tID=chromosome:1
1\tensembl\tgene\t100\t900\t.\t+\t.\tID=gene:G1;Name=abc;biotype=protein_coding
1\tensembl\tmRNA\t100\t900\t.\t+\t.\tID=transcript:T1;Parent=gene:G1
1\tensembl\texon\t100\t200\t.\t+\t.\tParent=transcript:T1;Name=E1;rank=1
1\tensembl\tCDS\t120\t200\t.\t+\t.\tParent=transcript:T1
1\tensembl\texon\t400\t900\t.\t+\t.\tParent=transcript:T1;Name=E2;rank=2
2\tensembl\texon\t10\t20\t.\t-\t.\tParent=transcript:T9
1\tensembl\texon\t1500\t1600\t.\t-\t.\tParent=transcript:T2;Name=E3
""".splitlines()
    a = workloads.Annotation.from_gff_lines(gff, "chr1")
    assert (a.kind, a.genes, a.start, a.end, a.strand) == (2, ["T1", "T1", "T2"], [100, 400, 1500], [200, 900, 1600], [True, True, False])
    assert a.name_id == [0, 0, 2]
    gtf = """chr1\tHAVANA\texon\t13453\t13670\t.\t+\t.\tgene_id "ENSG1.5"; transcript_id "ENST1.2"; gene_name "DDX11L1";
chr1\tHAVANA\tUTR\t13453\t13500\t.\t+\t.\tgene_id "ENSG1.5";
""".splitlines()
    b = workloads.Annotation.from_gff_lines(gtf, "1", gtf=True, features=("gene_name", "description", "gene_id", "gene_type", "gene_id"))
    assert (b.genes, b.start, b.end, b.strand) == (["ENSG1.5"], [13453], [13670], [True])
    h = b.name_hash[0]
    jh = 0
    for ch in "ENSG1.5":
        jh = (31 * jh + ord(ch)) & 0xFFFFFFFF
    assert h == (jh ^ (jh >> 16))
