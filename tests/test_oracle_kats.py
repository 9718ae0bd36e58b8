"""Known-answer tests for the oracle, hand-derived from the reference's lines.  The reference ships no
tests or golden vectors of its own (parity unpinned), so these pin the restatement to the cited Java lines.
J = /root/reference/java/src/main/java/npTranscript."""
import numpy as np

from nptranscript_b200 import abi, render, synth, workloads
from oracle import oracle

G = 29893
ANN = workloads.SARS2_VIC01


def run(reads, ref=None, **kw):
    ref = np.full(G, 1, np.uint8) if ref is None else ref  # all 'A'
    b = synth.from_reads(reads)
    cfg = oracle.make_config(**dict(dict(bin=10, break_thresh=1000), **kw))
    return oracle.run(cfg, ref, ANN, [b])


def breaks_of(r, i=0):
    bo = r["break_off"]
    return r["breaks"][int(bo[i]):int(bo[i + 1])].tolist()


def nz(a):
    idx = np.flatnonzero(a)
    return {int(i): int(a[i]) for i in idx}


def test_kat1_canonical_n_sgrna():
    """pos=5 10S65M28190N1600M30S, all bases match (IdentityProfile1.java:482-585, CigarCluster.java:442-498)."""
    r = run([(5, 0, 0, "10S65M28190N1600M30S", "A" * (10 + 65 + 1600 + 30))])
    assert breaks_of(r) == [5, 69, 28260, 29859]
    assert render.cluster_keys(r, ANN, 0) == ["leader_leader,N_3UTR"]
    assert r["sub_key"].tolist() == [0, 6, 2826, 2985]                       # floor(b/10), CigarCluster.java:620-631
    assert (r["pair_level"].tolist(), r["pair_start"].tolist(), r["pair_end"].tolist(), r["pair_count"].tolist()) == ([0], [69], [28260], [1])
    assert (int(r["read_st"][0]), int(r["read_en"][0])) == (11, 1675)        # htsjdk read offsets incl. the soft clip
    assert int(r["read_flags"][0]) & abi.NPT_RF_LEADER
    assert int(r["read_flags"][0]) & abi.NPT_RF_TYPE_MASK == 0               # 5_3: start<=100, end>=G-100


def test_kat2_breaks_follow_matching_bases():
    """Mismatches next to the junction move the break to the nearest matching base (IdentityProfile1.java:531-551)."""
    bases = ["A"] * (65 + 1600)
    for k in (63, 64, 65, 66, 67):
        bases[k] = "C"
    r = run([(5, 0, 0, "65M28190N1600M", "".join(bases))])
    assert breaks_of(r) == [5, 67, 28263, 29859]


def test_kat3_deletion_quirk_depth():
    """pos=100 5M2D5M: the D emits its positions AFTER advancing (:513-522): depth lands on 107,108 as errors."""
    r = run([(100, 0, 0, "5M2D5M", "A" * 10)], record_depth=1)
    assert breaks_of(r) == [100, 111]
    assert nz(r["depth"][0, 0]) == {100: 1, 101: 1, 102: 1, 103: 1, 104: 1, 107: 2, 108: 2, 109: 1, 110: 1, 111: 1}
    assert nz(r["errors"][0, 0]) == {107: 1, 108: 1}
    assert (int(r["maps_sum"][0]), int(r["err_sum"][0])) == (12, 2)


def test_kat4_eq_x_quirk():
    """pos=100 4=1X3=: = and X advance first, then emit (:552-576)."""
    r = run([(100, 0, 0, "4=1X3=", "A" * 8)], record_depth=1)
    assert breaks_of(r) == [100, 107]
    assert nz(r["depth"][0, 0]) == {104: 1, 105: 2, 106: 1, 107: 1, 108: 1, 109: 1, 110: 1}
    assert nz(r["errors"][0, 0]) == {105: 1}


def test_kat5_map_start_end_fire_on_mismatches():
    """pos=100 10M2000N10M, first two bases after the N mismatch: mapStart/mapEnd fire until a match (CigarCluster.java:479-484)."""
    r = run([(100, 0, 0, "10M2000N10M", "A" * 10 + "CC" + "A" * 8)], record_depth=1)
    assert breaks_of(r) == [100, 109, 2112, 2119]
    ms, me = nz(r["map_start"][0, 0]), nz(r["map_end"][0, 0])
    assert ms == {100: 1, 2110: 1, 2111: 1, 2112: 1}     # + setStartEnd(startPos) (CigarCluster.java:360-365)
    assert me == {109: 3, 2119: 1}                        # + setStartEnd(endPos)


def test_first_appearance_ids_and_sub_ids():
    """Cluster n = rank of first appearance (CigarClusters.java:83); sub k = rank inside the cluster (CigarCluster.java:660)."""
    a = (5, 0, 0, "65M28190N1600M", "A" * 1665)          # leader_leader,N_3UTR
    b = (28300, 0, 0, "1500M", "A" * 1500)               # N_3UTR
    a2 = (25, 0, 0, "45M28190N1600M", "A" * 1645)        # same cluster as a, other sub (start bin 2)
    r = run([a, b, a, a2, b, a2])
    assert r["cluster_id"].tolist() == [0, 1, 0, 0, 1, 0]
    assert r["sub_id"].tolist() == [0, 0, 0, 1, 0, 1]
    assert render.cluster_keys(r, ANN, 0) == ["leader_leader,N_3UTR", "N_3UTR"]
    assert r["cl_read_count"][:, 0].tolist() == [4, 2]
    assert r["sub_count"][:, 0].tolist() == [2, 2, 2]
    assert r["cl_first_read"].tolist() == [0, 1]
    assert r["sub_first_read"].tolist() == [0, 3, 1]
    assert r["total_counts"].tolist() == [6]


def test_break_point_levels_and_orf_counts():
    """Level 0 for the first internal gap > T, 1 for the rest (IdentityProfile1.java:276-286); ORF counters (:349-356)."""
    r = run([(5, 0, 0, "65M10000N500M15000N2000M", "A" * 2565)])
    assert breaks_of(r) == [5, 69, 10070, 10569, 25570, 27569]
    assert sorted(zip(r["pair_level"].tolist(), r["pair_start"].tolist(), r["pair_end"].tolist())) == [(0, 69, 10070), (1, 10569, 25570)]
    # st1 = 25570: the loop runs from the last ORF down to E (26245 >= st1-10; ORF3a 25393 breaks it);
    # endPos(27569) > start+100 holds for E, M, ORF6, ORF7a (not ORF8 27894 and later)
    genes = ANN.genes
    got = {genes[i] for i in np.flatnonzero(r["spliced"][0])}
    assert got == {"E", "M", "ORF6", "ORF7a"}
    assert r["unspliced"].sum() == 0


def test_alignment_past_reference_end_and_bad_ops():
    """Emission stops at the reference end (every loop is bounded by refPos+i < refSeq.length()); op > 8 is not clustered."""
    r = run([(G - 5, 0, 0, "20M", "A" * 20), (10, 0, 0, "5M3?5M", "A" * 13)], record_depth=1)
    assert breaks_of(r, 0) == [G - 5, G + 14]
    assert int(r["maps_sum"][0]) == 6
    assert int(r["read_flags"][1]) & abi.NPT_RF_BADCIGAR and int(r["cluster_id"][1]) == -1


def test_host_mode_trims_short_terminal_exons():
    """min_first_last_exon_length=10 in host mode (ViralTranscriptAnalysisCmd2.java:360; IdentityProfile1.java:205-228)."""
    b = synth.from_reads([(100, 0, 0, "5M300N500M300N4M", "A" * 509)])
    cfg = oracle.make_config(bin=100, break_thresh=50, coronavirus=0, include_start=0)
    r = oracle.run(cfg, np.full(G, 1, np.uint8), workloads.Annotation.empty(), [b], chrom_index=2)
    assert breaks_of(r) == [405, 904]
    assert render.cluster_keys(r, workloads.Annotation.empty(), 2, coronavirus=False) == ["2.9"]
