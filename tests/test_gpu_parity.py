"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded reads.  Bit-exact."""
import numpy as np
import pytest

from nptranscript_b200 import abi, engine, synth, workloads
from oracle import oracle
from tests.common import assert_same
from tests.fuzz import gff_case, random_batch

pytestmark = pytest.mark.gpu


def _both(cfg_kw, w, batches, annotation=None, chrom_index=0, genome=None):
    g = synth.genome_codes(w) if genome is None else genome
    ann = annotation or w.annotation
    dev = engine.run(engine.make_config(**cfg_kw), g, ann, batches, chrom_index)
    ora = oracle.run(oracle.make_config(**cfg_kw), g, ann, batches, chrom_index)
    return dev, ora


@pytest.mark.parametrize("depth", [0, 1])
def test_sars2_20k(depth):
    w = workloads.SARS2
    b = synth.generate(w, 20200301, 20000)
    dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=depth), w, [b])
    assert_same(dev, ora)
    assert dev["n_slow_reads"] < 5 * 20000  # exact per-base windows (junction windows): a diagnostic, ~1 per read


def test_config1_as_specified_from_the_reference_files():
    """BASELINE.json configs[0]: VIC01 reference FASTA + Coordinates.csv (copies of the reference's own data files under
    tests/golden/), 100 000 synthetic direct-RNA reads, seed 20200301, run_example.sh flags (--bin 10 --breakThresh 1000)
    + --RNA=true --writeGFF=true --maxThreads=1; then the 0.annot rows against the reference-held sample"""
    from nptranscript_b200 import render
    w = workloads.sars2_vic01_real()
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200301, 100000)
    kw = dict(bin=10, break_thresh=1000, track_break_sums=1)
    dev = engine.run(engine.make_config(**kw), g, w.annotation, [b])
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, [b])
    assert_same(dev, ora)
    rows = render.annot_rows(dev, w.annotation, seqlen=w.genome_len, polya=w.polya_len)
    want = open(workloads.golden_path("vic01", "0.annot.tsv")).read().splitlines()
    assert [x.split("\t")[:3] for x in rows[1:]] == [x.split("\t")[:3] for x in want[1:]]


@pytest.mark.parametrize("which", ["nc045512", "u13369"])
def test_other_reference_pairs_the_reference_ships(which):
    """NC_045512.2 with its overlapping ORF7a / ORF7b rows, and the 43 kb rDNA repeat U13369 with IUPAC codes in the reference
    and no leader row (data files of the reference, copied under tests/golden/)"""
    w = workloads.nc045512_real() if which == "nc045512" else workloads.u13369_real()
    S = 1 if which == "nc045512" else 2
    b = synth.generate(w, 20200301, 30000, num_sources=S)
    dev, ora = _both(dict(bin=10 if S == 1 else 100, break_thresh=1000 if S == 1 else 500, record_depth=1, num_sources=S, max_coverage_clusters=1024), w, [b])
    assert_same(dev, ora)


def test_config2_flags_200k():
    """the benchmark's own flag set (bin=100, breakThresh=1000, depth recording) on 200 k reads"""
    w = workloads.SARS2
    b = synth.generate(w, 20200302, 200000)
    dev, ora = _both(dict(bin=100, break_thresh=1000, record_depth=1), w, [b])
    assert_same(dev, ora)
    assert 0 < dev["n_slow_reads"] < 0.02 * 200000   # the fast walk took (nearly) all of them


def test_first_is_transcriptome_many_host_batches():
    """firstIsTranscriptome + break sums with four host batches of different sizes: the end-of-chromosome fix-up reads the
    source index of every read long after the staging buffers were handed to later batches"""
    w = workloads.SARS2
    sizes = [3000, 7000, 1500, 5000]
    bs, first = [], 0
    for k, n in enumerate(sizes):
        bs.append(synth.generate(w, 77, n, first_index=first, num_sources=2))
        first += n
    dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2, rna=[1, 1], first_is_transcriptome=1), w, bs)
    assert_same(dev, ora)


def test_bad_source_index_is_an_error():
    w = workloads.SARS2
    b = synth.generate(w, 3, 500, num_sources=3)   # src in 0..2, configuration says 2 sources
    with pytest.raises(engine.NptError) as e:
        engine.run(engine.make_config(bin=10, num_sources=2, record_depth=1), synth.genome_codes(w), w.annotation, [b])
    assert e.value.code == abi.NPT_EINVAL


def test_multi_batch_and_sources():
    w = workloads.HCOV229E
    bs = [synth.generate(w, 20200310 + k, 5000, first_index=0, num_sources=4, src_fixed=k) for k in range(4)]
    dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=1, num_sources=4), w, bs)
    assert_same(dev, ora)


def test_eqx_cigars():
    w = workloads.SARS2
    b = synth.generate(w, 5, 8000, use_eqx=True)
    dev, ora = _both(dict(bin=100, break_thresh=500, record_depth=1), w, [b])
    assert_same(dev, ora)


@pytest.mark.parametrize("kw", [
    dict(bin=10, break_thresh=1000, record_depth=1),
    dict(bin=1, break_thresh=50, record_depth=1),
    dict(bin=100, break_thresh=8, record_depth=1),
    dict(bin=7, break_thresh=3, record_depth=1),            # breakThresh < 8: exact serial device walker
    dict(bin=10, break_thresh=1000, record_depth=0, include_start=0),
    dict(bin=10, break_thresh=1000, record_depth=1, enforce_strand=1),
    dict(bin=10, break_thresh=200, record_depth=1, start_thresh=30, end_thresh=500),
    dict(bin=10, break_thresh=40000, record_depth=1),       # breakThresh >= seqlen: no break-point matrix
])
def test_fuzz_virus(kw):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(1234 + kw["bin"] + kw["break_thresh"])
    small_t = kw["break_thresh"] < 8  # keep reads under the 256-breaks-per-read limit
    bs = [random_batch(rng, 1500, w.genome_len, g, num_sources=3, long_reads=(i == 1 and not small_t), max_ops=12 if small_t else 60)
          for i in range(3)]
    dev, ora = _both(dict(kw, num_sources=3, max_coverage_clusters=4096), w, bs)
    assert_same(dev, ora)


def test_fuzz_bad_ops():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(99)
    bs = [random_batch(rng, 2000, w.genome_len, g, bad_op=0.01)]
    dev, ora = _both(dict(bin=10, break_thresh=300, record_depth=1), w, bs)
    assert (ora["read_flags"] & abi.NPT_RF_BADCIGAR).any()
    assert_same(dev, ora)


@pytest.mark.parametrize("kw", [
    dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1], first_is_transcriptome=1),  # opts_host.txt
    dict(bin=10, break_thresh=1000, coronavirus=1, record_depth=1, rna=[1, 1], first_is_transcriptome=1),
    dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1]),
    dict(bin=10, break_thresh=50, coronavirus=0, include_start=1, record_depth=1, rna=[1, 1]),
    dict(bin=10, break_thresh=16, coronavirus=0, include_start=1, record_depth=1, rna=[1, 1], enforce_strand=1),
])
def test_fuzz_host_mode_empty_annotation(kw):
    """Host mode (opts_host.txt flag set minus firstIsTranscriptome) with EmptyAnnotation: chr.bin keys, exon trimming."""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(7)
    bs = [random_batch(rng, 2000, w.genome_len, g, num_sources=2, big_n=300)]
    dev, ora = _both(dict(kw, num_sources=2, max_coverage_clusters=4096), w, bs, annotation=workloads.Annotation.empty(), chrom_index=3)
    assert_same(dev, ora)


@pytest.mark.parametrize("kw,collide", [
    (dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1], first_is_transcriptome=1), False),  # opts_host.txt
    (dict(bin=10, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 0]), False),
    (dict(bin=5, break_thresh=50, coronavirus=0, include_start=1, record_depth=1, rna=[0, 0]), True),
    (dict(bin=10, break_thresh=50, coronavirus=0, include_start=0, enforce_strand=1, record_depth=0, rna=[1, 1]), True),
    (dict(bin=1, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1]), False),
])
def test_fuzz_host_mode_gff_annotation(kw, collide):
    """Host mode with an exon table (GFFAnnotation): per-block exon lookup, gene-set keys in HashSet order, coordinate keys."""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    lut = np.full(16, ord("N"), np.uint8)
    lut[[1, 2, 4, 8]] = [ord(ch) for ch in "ACGT"]
    rng = np.random.default_rng(23 + kw["bin"])
    ann, b = gff_case(rng, 6000, w.genome_len, lut[g].tobytes().decode(), n_genes=60, hash_collide=collide)
    bs = [b, random_batch(rng, 500, w.genome_len, g, num_sources=2, big_n=300)]
    dev, ora = _both(dict(kw, num_sources=2, max_coverage_clusters=8192), w, bs, annotation=ann, chrom_index=3)
    assert_same(dev, ora)
    toks = dev["cl_key_tok"]
    assert (toks == abi.NPT_TOK_GENESET).sum() > 20 and dev["n_clusters"] > (toks == abi.NPT_TOK_GENESET).sum()


def test_empty_annotation_needs_rna():
    """coronavirus=0 + EmptyAnnotation + rna[src]=0: the reference unboxes a null Boolean (IdentityProfile1.java:293-299) and
    throws; the library rejects the combination instead of inventing semantics"""
    w = workloads.SARS2
    for rna in ([0, 1], [1, 0]):
        with pytest.raises(engine.NptError) as e:
            engine.run(engine.make_config(coronavirus=0, num_sources=2, rna=rna), synth.genome_codes(w), workloads.Annotation.empty(),
                       [synth.generate(w, 1, 10)])
        assert e.value.code == abi.NPT_EINVAL


def test_gff_needs_host_mode():
    w = workloads.SARS2
    ann = workloads.Annotation.gff(["g"], [10], [100], [True])
    with pytest.raises(engine.NptError):
        engine.run(engine.make_config(coronavirus=1), synth.genome_codes(w), ann, [synth.generate(w, 1, 10)])


def test_small_tables_report_capacity():
    w = workloads.SARS2
    b = synth.generate(w, 1, 5000)
    with pytest.raises(engine.NptError) as e:
        engine.run(engine.make_config(bin=1, break_thresh=1000, max_subclusters=64), synth.genome_codes(w), w.annotation, [b])
    assert e.value.code == abi.NPT_ECAPACITY


@pytest.mark.parametrize("depth", [0, 1])
def test_many_breaks_overflow_area(depth):
    """Reads with > BRK_INLINE breaks exercise the overflow area and both of its re-run paths: without depth recording the
    fast walk stores the long lists itself and finds the area too small (whole batch redone from the store-all pass); with
    depth recording those reads go to the serial walker, whose chain alone is redone beside the finished bulk."""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(5)
    bs = [random_batch(rng, 6000, w.genome_len, g, big_n=400, max_ops=120)]
    dev, ora = _both(dict(bin=10, break_thresh=100, record_depth=depth), w, bs)
    nb = np.diff(ora["break_off"].astype(np.int64))
    assert (nb > 6).sum() > 1500
    assert_same(dev, ora)


EDGE_READS = [
    (5, 0, 0, "10S65M28190N1600M30S", None),          # KAT1 shape
    (1, 0, 0, "1M", None),                             # shortest possible alignment
    (100, 16, 0, "20S", None),                         # no reference-consuming op at all
    (100, 0, 0, "5I", None),
    (100, 0, 0, "3H2P4S", None),
    (29890, 0, 0, "10M", None),                        # runs past the end of the reference
    (29893, 0, 0, "1M5D3M", None),
    (200, 0, 0, "0M5M0D0N7M", None),                   # zero-length elements
    (300, 0, 0, "4M268435455N4M", None),               # the longest CIGAR element BAM can hold
    (400, 0, 0, "50M1200D50M", None),                  # a deletion longer than the break threshold
    (500, 0, 0, "33M1I31M2D65M1X3=40M", None),         # step boundaries at 32/33 bases
    (600, 16, 0, "5M3N5M3N5M3N5M3N5M3N5M3N5M3N5M", None),
    (700, 0, 0, "30M", "ACGTACGTAC"),                  # the CIGAR wants more bases than the record holds: the reference throws
    (710, 0, 0, "10M5S", "ACGTACGTACGT"),              # ... but a clip past the end is never looked at
    (29890, 0, 0, "10M", "ACGTAC"),                    # the part of the M op inside the reference fits the 6 bases
]


def _edge_batch(g):
    lut = np.full(16, ord("N"), np.uint8)
    lut[[1, 2, 4, 8]] = [ord(ch) for ch in "ACGT"]
    gs = lut[g].tobytes().decode()
    import re
    reads = []
    for pos, flag, src, cig, given in EDGE_READS:
        if given is not None:
            reads.append((pos, flag, src, cig, given))
            continue
        ref, bases = pos - 1, []
        for ln, op in re.findall(r"(\d+)([MIDNSHP=X])", cig):
            ln = int(ln)
            if op in "M=X":
                seg = "".join(gs[i] if 0 <= i < len(gs) else "A" for i in range(ref, min(ref + ln, ref + 4096)))
                bases.append(seg if op != "X" else "".join("ACGT"[("ACGT".index(c) + 1) % 4] if c in "ACGT" else "A" for c in seg))
                ref += ln
            elif op in "IS":
                bases.append("C" * ln)
            elif op in "DN":
                ref += ln
        reads.append((pos, flag, src, cig, "".join(bases)))
    return synth.from_reads(reads)


@pytest.mark.parametrize("kw", [
    dict(bin=10, break_thresh=1000, record_depth=1),
    dict(bin=10, break_thresh=2, record_depth=1),
    dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=1),
])
def test_edge_case_reads(kw):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = _edge_batch(g)
    ann = w.annotation if kw.get("coronavirus", 1) else workloads.Annotation.empty()
    dev, ora = _both(dict(kw, max_coverage_clusters=64), w, [b, b.slice(0, 0), b.slice(3, 4), b], annotation=ann)
    assert_same(dev, ora)


def test_no_reads_at_all():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    empty = synth.generate(w, 1, 4).slice(0, 0)
    for batches in ([], [empty]):
        dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=1), w, batches)
        assert dev["n_reads"] == 0 and dev["n_clusters"] == 0 and dev["n_subs"] == 0 and dev["n_pairs"] == 0
        assert_same(dev, ora)


def test_bam_file_to_results(tmp_path):
    """The caller's side of the seam: BAM file -> libnptbam (filter chain, pinned SoA batches) -> npt_submit, two source
    files one after the other (sequential multi-BAM order, src = file index) -> same results as the oracle on the reads."""
    import os
    from nptranscript_b200 import bam
    w = workloads.SARS2
    g = synth.genome_codes(w)
    parts = [synth.generate(w, 31 + k, 4000, num_sources=2, src_fixed=k) for k in range(2)]
    paths = []
    for k, b in enumerate(parts):
        p = os.path.join(str(tmp_path), f"s{k}.bam")
        bam.write_bam(p, [(w.name, w.genome_len)], bam.records_from_batch(b), block_bytes=20000)
        paths.append(p)
    kw = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2)
    eng = engine.Engine(engine.make_config(**kw))
    try:
        eng.begin_chrom(0, g, w.annotation)
        for k, p in enumerate(paths):
            rd = bam.Reader(p, nthreads=2, pinned_from=engine.lib())
            for ref_id, b, extra in rd.batches(src=k, max_batch_reads=1500, copy=False):
                assert ref_id == 0
                eng.submit(b)
                eng.wait()  # the reader reuses its buffers for the next batch
            rd.close()
        eng.end_chrom()
        dev = eng.fetch()
    finally:
        eng.close()
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, parts)
    assert_same(dev, ora)


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_fuzz_more_seeds(seed):
    """More random CIGARs around the step / tile boundaries of the walk: thresholds near the 32-base step, long reads,
    many mismatches, weird ops, all with coverage on."""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(seed)
    T = int(rng.choice([16, 31, 32, 33, 64, 100, 500]))
    kw = dict(bin=int(rng.choice([1, 10, 100])), break_thresh=T, record_depth=1, include_start=int(rng.integers(0, 2)),
              enforce_strand=int(rng.integers(0, 2)), num_sources=2, max_coverage_clusters=8192)
    bs = [random_batch(rng, 1200, w.genome_len, g, num_sources=2, weird=float(rng.choice([0.0, 0.3, 0.8])),
                       mismatch=float(rng.choice([0.02, 0.1, 0.4])), long_reads=bool(i == 0), max_ops=80) for i in range(2)]
    dev, ora = _both(kw, w, bs)
    assert_same(dev, ora)


def test_229e_eight_sources():
    """BASELINE config 3 in small: 229E genome + ORF table, eight source BAMs processed one after the other."""
    w = workloads.HCOV229E
    bs = [synth.generate(w, 20200310 + k, 5000, num_sources=8, src_fixed=k) for k in range(8)]
    kw = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=8, max_coverage_clusters=512)
    dev, ora = _both(kw, w, bs)
    assert_same(dev, ora)
    assert (dev["total_counts"] == 5000).all()


@pytest.mark.parametrize("kw,empty", [
    (dict(bin=10, break_thresh=1000, record_depth=1, max_coverage_clusters=512), False),
    (dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1]), True),
])
def test_reference_too_large_for_shared_memory(kw, empty):
    """A 200 kb reference does not fit the walk's shared-memory copy (96 KB packed): the kernels read it through L1/L2."""
    w = workloads.SARS2
    rng = np.random.default_rng(8)
    G = 200_003
    g = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, G)]
    bs = [random_batch(rng, 800, G, g, num_sources=2, big_n=3000, long_reads=(i == 0)) for i in range(2)]
    ann = workloads.Annotation.empty() if empty else w.annotation
    dev, ora = _both(dict(kw, num_sources=2), w, bs, annotation=ann, genome=g, chrom_index=1)
    assert_same(dev, ora)


def test_reference_forced_out_of_shared_memory(monkeypatch):
    """Same kernels, template instance without the shared-memory reference, on the benchmark genome."""
    monkeypatch.setenv("NPT_REF_SMEM", "0")
    w = workloads.SARS2
    b = synth.generate(w, 41, 6000)
    dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=1), w, [b])
    assert_same(dev, ora)


def test_chromosomes_one_after_the_other_on_one_context():
    """The reference's chromosome switch (ViralTranscriptAnalysisCmd2.java:830-901): one context, several begin/end cycles
    with different references, annotations and batch sizes; results of each equal a fresh oracle run; fetch twice; a
    cycle without fetch in between."""
    rng = np.random.default_rng(21)
    w1, w2 = workloads.SARS2, workloads.HCOV229E
    G3 = 250_001
    g3 = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, G3)]
    plan = [
        (0, w1, synth.genome_codes(w1), w1.annotation, [synth.generate(w1, 3, 5000), synth.generate(w1, 4, 100)]),
        (1, w2, synth.genome_codes(w2), w2.annotation, [synth.generate(w2, 5, 3000)]),
        (2, w1, g3, w1.annotation, [random_batch(rng, 700, G3, g3, big_n=3000)]),   # larger reference: reallocation + global-memory reference
        (3, w1, synth.genome_codes(w1), w1.annotation, []),                          # a chromosome without reads
        (4, w1, synth.genome_codes(w1), w1.annotation, [synth.generate(w1, 6, 2000)]),
    ]
    kw = dict(bin=10, break_thresh=1000, record_depth=1, max_coverage_clusters=256)
    eng = engine.Engine(engine.make_config(**kw))
    try:
        for ci, w, g, ann, bs in plan:
            eng.begin_chrom(ci, g, ann)
            for b in bs:
                eng.submit(b)
            eng.end_chrom()
            if ci == 1:
                eng.release()  # results never fetched
                continue
            dev = eng.fetch()
            again = eng.fetch(abi.NPT_FETCH_TABLES)
            assert np.array_equal(dev["sub_count"], again["sub_count"]) and np.array_equal(dev["cl_key_tok"], again["cl_key_tok"])
            ora = oracle.run(oracle.make_config(**kw), g, ann, bs, ci)
            assert_same(dev, ora)
            eng.release()
    finally:
        eng.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_flag_combinations(seed):
    """Flags drawn at random (both modes, all three annotation kinds, 1-3 sources): device against oracle."""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(2000 + seed)
    corona = bool(rng.integers(0, 2))
    S = int(rng.integers(1, 4))
    kw = dict(bin=int(rng.choice([1, 5, 10, 100])), break_thresh=int(rng.choice([2, 20, 50, 400, 1000])), coronavirus=int(corona),
              include_start=int(rng.integers(0, 2)), record_depth=int(rng.integers(0, 2)), num_sources=S,
              start_thresh=int(rng.choice([0, 100, 1000])), end_thresh=int(rng.choice([0, 100, 1000])), max_coverage_clusters=1024)
    if corona:
        kw["enforce_strand"] = int(rng.integers(0, 2))
        kw["rna"] = [1] * S
        ann = w.annotation
        bs = [random_batch(rng, 1500, w.genome_len, g, num_sources=S, max_ops=12 if kw["break_thresh"] < 8 else 40), synth.generate(w, seed, 2000, num_sources=S)]
    else:
        kw["rna"] = [int(x) for x in rng.integers(0, 2, S)]
        kw["first_is_transcriptome"] = int(rng.integers(0, 2))
        kw["max_coverage_clusters"] = 4096
        if rng.integers(0, 2):
            ann = workloads.Annotation.empty()
            kw["rna"] = [1] * S   # EmptyAnnotation + rna=0 throws in the reference and is rejected by the library
            bs = [random_batch(rng, 2000, w.genome_len, g, num_sources=S, big_n=300)]
        else:
            lut = np.full(16, ord("N"), np.uint8)
            lut[[1, 2, 4, 8]] = [ord(ch) for ch in "ACGT"]
            ann, b = gff_case(rng, 3000, w.genome_len, lut[g].tobytes().decode(), n_genes=50, num_sources=S, hash_collide=bool(rng.integers(0, 2)))
            bs = [b]
    dev, ora = _both(kw, w, bs, annotation=ann, chrom_index=int(rng.integers(0, 5)))
    assert_same(dev, ora)
