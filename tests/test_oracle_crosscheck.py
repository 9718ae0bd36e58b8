"""oracle.cpp against the independent pure-Python restatement (oracle/py_oracle.py) on seeded random reads:
two restatements written separately from the Java sources must agree field by field, including the rendered
string keys and ids."""
import numpy as np
import pytest

from nptranscript_b200 import abi, render, synth, workloads
from oracle import oracle, py_oracle
from tests.fuzz import random_batch


def _compare(kw, annotation, batches, chrom_index=0, w=None):
    w = w or workloads.SARS2
    g = synth.genome_codes(w)
    ora = oracle.run(oracle.make_config(**kw), g, annotation, batches, chrom_index)
    fl = py_oracle.Flags(bin=kw.get("bin", 1), break_thresh=kw.get("break_thresh", 1000), include_start=bool(kw.get("include_start", 1)),
                         coronavirus=bool(kw.get("coronavirus", 1)), start_thresh=kw.get("start_thresh", 100), end_thresh=kw.get("end_thresh", 100),
                         enforce_strand=bool(kw.get("enforce_strand", 0)), record_depth=bool(kw.get("record_depth", 0)),
                         num_sources=kw.get("num_sources", 1), rna=tuple(bool(x) for x in kw.get("rna", [1] * 16)),
                         first_is_transcriptome=bool(kw.get("first_is_transcriptome", 0)))
    empty = annotation.kind == abi.NPT_ANNOT_EMPTY
    ann = py_oracle.Annot(annotation.genes, annotation.start, annotation.end, annotation.strand, len(g), fl, empty=empty, chrom_index=chrom_index,
                          gff=annotation.kind == abi.NPT_ANNOT_GFF)
    run = py_oracle.Run(fl, g.tolist(), ann, chrom_index)
    names, srcs = [], []
    for b in batches:
        for rd in py_oracle.unpack_batch(b):
            names.append(f"read{len(names)}")
            srcs.append(rd[2])
            run.process(*rd, read_name=names[-1])
    keys = render.cluster_keys(ora, annotation, chrom_index, coronavirus=fl.coronavirus)
    render.check_key_aliasing(keys)
    assert keys == [c.key for c in run.clusters.values()]
    bo = ora["break_off"]
    subn = render.sub_names(ora, srcs, names, chrom_index, fl.first_is_transcriptome)
    for i, row in enumerate(run.rows):
        if not row["ok"]:
            assert int(ora["read_flags"][i]) & abi.NPT_RF_BADCIGAR
            continue
        c, k = int(ora["cluster_id"][i]), int(ora["sub_id"][i])
        assert render.cluster_id(chrom_index, c) == row["cluster"] and subn[i] == row["sub"], i
        assert ora["breaks"][int(bo[i]):int(bo[i + 1])].tolist() == row["breaks"], i
        assert (int(ora["start_pos"][i]), int(ora["end_pos"][i]), int(ora["read_st"][i]), int(ora["read_en"][i])) == \
               (row["start_pos"], row["end_pos"], row["read_st"], row["read_en"]), i
        fl_i = int(ora["read_flags"][i])
        assert (fl_i & 3, bool(fl_i & 4), bool(fl_i & 8)) == (row["type_ind"], row["leader"], row["neg"]), i
        assert (int(ora["maps_sum"][i]), int(ora["err_sum"][i])) == (row["maps_sum"], row["err_sum"]), i
    assert ora["total_counts"].tolist() == run.total_counts
    assert ora["spliced"].tolist() == ann.spliced and ora["unspliced"].tolist() == ann.unspliced
    got = {(int(a), int(b_), int(c_), int(d)): int(e) for a, b_, c_, d, e in zip(ora["pair_src"], ora["pair_level"], ora["pair_start"], ora["pair_end"], ora["pair_count"])}
    assert got == run.pairs
    so, ko, to = ora["cl_sub_off"], ora["sub_key_off"], ora["sub_sum_off"]
    for ci, cl in enumerate(run.clusters.values()):
        assert ora["cl_read_count"][ci].tolist() == cl.read_count
        subs = list(cl.subs.items())
        assert int(so[ci + 1]) - int(so[ci]) == len(subs)
        for k, (br, sub) in enumerate(subs):
            g_ = int(so[ci]) + k
            assert tuple(ora["sub_key"][int(ko[g_]):int(ko[g_ + 1])].tolist()) == br
            assert ora["sub_count"][g_].tolist() == sub["count"]
            assert int(ora["sub_break_sum"][g_]) == sub["break_sum"]
            tb = ora["sub_true_breaks"][int(to[g_]):int(to[g_ + 1])]
            assert np.array_equal(tb, np.array(sub["true_breaks"])), "arrival-order double sums must be bit-identical"
            cons = render.consensus_breaks(ora, g_, srcs, fl.first_is_transcriptome)
            ref = [int(np.floor(t * (1e3 / sub["break_sum"]) + 0.5)) for t in sub["true_breaks"]]
            assert cons.tolist() == ref
        if fl.record_depth:
            for name, d in (("depth", cl.depth), ("errors", cl.errors), ("map_start", cl.map_start), ("map_end", cl.map_end)):
                dense = np.zeros_like(ora[name][ci])
                for (s, p), v in d.items():
                    if 0 <= p < dense.shape[1]:
                        dense[s, p] = v
                assert np.array_equal(ora[name][ci], dense), (name, ci)


@pytest.mark.parametrize("kw", [
    dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2),
    dict(bin=100, break_thresh=50, record_depth=1, num_sources=2),
    dict(bin=7, break_thresh=3, record_depth=1, num_sources=2),
    dict(bin=10, break_thresh=500, include_start=0, num_sources=2),
    dict(bin=10, break_thresh=1000, enforce_strand=1, record_depth=1, num_sources=2),
])
def test_virus_mode(kw):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(11 + kw["bin"])
    small_t = kw["break_thresh"] < 8
    bs = [random_batch(rng, 250, w.genome_len, g, num_sources=2, max_ops=12 if small_t else 40), synth.generate(w, 3, 150, num_sources=2)]
    _compare(kw, w.annotation, bs)


@pytest.mark.parametrize("kw", [
    dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, num_sources=2, rna=[1, 1], first_is_transcriptome=1),
    dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, num_sources=2, rna=[1, 0]),
    dict(bin=10, break_thresh=50, coronavirus=0, include_start=1, record_depth=1, num_sources=2, rna=[1, 1]),
])
def test_host_mode_empty_annotation(kw):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(5)
    _compare(kw, workloads.Annotation.empty(), [random_batch(rng, 300, w.genome_len, g, num_sources=2, big_n=300)], chrom_index=4)


def test_threads_do_not_change_results():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 9, 3000)
    cfg = oracle.make_config(bin=10, break_thresh=1000, record_depth=1)
    a, c = oracle.run(cfg, g, w.annotation, [b], nthreads=1), oracle.run(cfg, g, w.annotation, [b], nthreads=4)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, c[k]), k


def _genome_str(g):
    lut = np.full(16, ord("N"), np.uint8)
    lut[[1, 2, 4, 8]] = [ord(c) for c in "ACGT"]
    return lut[g].tobytes().decode()


@pytest.mark.parametrize("kw,collide", [
    (dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, num_sources=2, rna=[1, 1], first_is_transcriptome=1), False),
    (dict(bin=10, break_thresh=50, coronavirus=0, include_start=0, num_sources=2, rna=[1, 0]), False),
    (dict(bin=5, break_thresh=50, coronavirus=0, include_start=1, record_depth=1, num_sources=2, rna=[0, 0]), True),
    (dict(bin=10, break_thresh=50, coronavirus=0, include_start=0, enforce_strand=1, num_sources=2, rna=[1, 1]), True),
])
def test_host_mode_gff_annotation(kw, collide):
    from tests.fuzz import gff_case
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(17 + kw["bin"])
    ann, b = gff_case(rng, 400, w.genome_len, _genome_str(g), hash_collide=collide)
    _compare(kw, ann, [b, random_batch(rng, 80, w.genome_len, g, num_sources=2, big_n=300)], chrom_index=3)


def test_java_hashset_order_known_cases():
    # "Aa" and "BB" share String.hashCode() 2112: same bucket, insertion order decides
    assert py_oracle.java_string_hash("Aa") == py_oracle.java_string_hash("BB") == 2112
    assert py_oracle.java_hashset_order(["BB", "Aa"]) == ["BB", "Aa"] and py_oracle.java_hashset_order(["Aa", "BB", "Aa"]) == ["Aa", "BB"]
    # single characters hash to their code point: buckets 1, 2, 3 in a 16-bucket table whatever the insertion order
    assert py_oracle.java_hashset_order(["c", "a", "b"]) == ["a", "b", "c"]
    # 13 entries exceed 0.75 * 16: the table doubles, 'q' (113 & 31 = 17) now sorts before 'a' (97 & 31 = 1)?  no: 1 < 17
    thirteen = [chr(97 + i) for i in range(12)] + ["q"]
    assert py_oracle.java_hashset_order(thirteen) == sorted(thirteen, key=lambda s: ord(s) & 31)


def test_edge_case_reads():
    """Shortest alignment, reads without reference-consuming ops, clips/pads only, alignment past the reference end,
    zero-length and maximum-length elements, deletion longer than the threshold, 32/33-base step boundaries, many junctions."""
    from tests.test_gpu_parity import _edge_batch
    w = workloads.SARS2
    b = _edge_batch(synth.genome_codes(w))
    for kw in (dict(bin=10, break_thresh=1000, record_depth=1), dict(bin=10, break_thresh=2, record_depth=1)):
        _compare(kw, w.annotation, [b, b.slice(3, 4), b])
    _compare(dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=1), workloads.Annotation.empty(), [b])


@pytest.mark.parametrize("seed", range(6))
def test_random_flag_combinations(seed):
    """Flags drawn at random (both modes, all three annotation kinds): the two restatements must agree on everything."""
    from tests.fuzz import gff_case
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(1000 + seed)
    corona = bool(rng.integers(0, 2))
    S = int(rng.integers(1, 4))
    kw = dict(bin=int(rng.choice([1, 5, 10, 100])), break_thresh=int(rng.choice([2, 20, 50, 400, 1000])), coronavirus=int(corona),
              include_start=int(rng.integers(0, 2)), record_depth=int(rng.integers(0, 2)), num_sources=S,
              start_thresh=int(rng.choice([0, 100, 1000])), end_thresh=int(rng.choice([0, 100, 1000])))
    if corona:
        kw["enforce_strand"] = int(rng.integers(0, 2))
        kw["rna"] = [1] * S
        ann = w.annotation
        bs = [random_batch(rng, 150, w.genome_len, g, num_sources=S, max_ops=12 if kw["break_thresh"] < 8 else 40), synth.generate(w, seed, 60, num_sources=S)]
    else:
        kw["rna"] = [int(x) for x in rng.integers(0, 2, S)]
        kw["first_is_transcriptome"] = int(rng.integers(0, 2))
        if rng.integers(0, 2):
            ann = workloads.Annotation.empty()
            bs = [random_batch(rng, 200, w.genome_len, g, num_sources=S, big_n=300)]
        else:
            ann, b = gff_case(rng, 200, w.genome_len, _genome_str(g), num_sources=S, hash_collide=bool(rng.integers(0, 2)))
            bs = [b]
    _compare(kw, ann, bs, chrom_index=int(rng.integers(0, 5)))
