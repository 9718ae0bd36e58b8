"""Small-size parity cases for the BASELINE.json configurations that bench.py --config 3|4|5 measures (config 1 and 2 have
theirs in test_gpu_parity.py).  Device (through the C ABI) against the oracle, bit-exact."""
import numpy as np
import pytest

from nptranscript_b200 import engine, synth, workloads
from oracle import oracle
from tests.common import assert_same

pytestmark = pytest.mark.gpu


def test_config3_229e_real_reference_sources_alternate():
    """configs[2] as bench.py --config 3 runs it below 8 GPUs: the real 229E FASTA + Coordinates.csv, sources alternating"""
    w = workloads.cov229e_real()
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200310, 40000, num_sources=8)
    kw = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=8, max_coverage_clusters=1024)
    dev = engine.run(engine.make_config(**kw), g, w.annotation, [b])
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, [b])
    assert_same(dev, ora)


def test_config4_waves_of_one_resident_batch():
    """configs[3]: the same batch submitted several times on one chromosome (more submits than batches may be in flight, so the
    early retirement of finished batches runs), as two interleaved 'ranks' would number them: wave k starts at 2 * k * n"""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200400, 6000)
    kw = dict(bin=100, break_thresh=1000, record_depth=1, max_coverage_clusters=512)
    e = engine.Engine(engine.make_config(**kw))
    try:
        e.begin_chrom(0, g, w.annotation)
        for k in range(5):
            e.set_next_read_index(2 * k * b.n)
            e.submit(b)
        e.end_chrom()
        dev = e.fetch()
    finally:
        e.close()
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, [b] * 5)
    # first-read indices carry the gaps of the interleaved numbering; everything else is what five plain submits give
    for k in ("cl_first_read", "sub_first_read"):
        v = dev[k]
        dev[k] = (v // (2 * b.n)) * b.n + v % (2 * b.n)
    assert_same(dev, ora)


@pytest.mark.parametrize("chrom", [20, 24])
def test_config5_host_chromosome(chrom):
    """configs[4]: one chromosome of the synthetic human-like model with the opts_host.txt flag set, source 0 = the transcripts
    themselves (firstIsTranscriptome); chr25 is the 20 kb 'chrM' (reference in shared memory), chr21 is 1.9 Mb (it is not)"""
    c = workloads.human_like()[chrom]
    g = synth.host_genome_codes(c)
    bs = [synth.generate_host(c, g, 20200500 + c.index, 30000, num_sources=2, transcriptome=True),
          synth.generate_host(c, g, 20200500 + c.index, 5000, first_index=30000, num_sources=2, transcriptome=True)]
    kw = dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, enforce_strand=0, record_depth=0, first_is_transcriptome=1,
              track_break_sums=1, num_sources=2, rna=[1, 1])
    dev = engine.run(engine.make_config(**kw), g, c.annotation, bs, chrom_index=c.index)
    ora = oracle.run(oracle.make_config(**kw), g, c.annotation, bs, c.index)
    assert_same(dev, ora)
    assert dev["n_clusters"] > 10


def test_many_clusters_default_coverage_capacity_and_reuse():
    """bin=1 on fuzzed reads: thousands of clusters with coverage rows.  The default capacity (what fits 12 GB, here the
    whole cluster table) holds them; the second chromosome on the same context finds the rows zero again (they are cleared
    after use at end of chromosome instead of memset per chromosome); an explicit small capacity is an error, not a drop"""
    from tests.fuzz import random_batch
    g = synth.genome_codes(workloads.SARS2)[:6000].copy()
    ann = workloads.Annotation.empty()   # host mode: coordinate keys, one cluster per distinct (start, end) pair at bin=1
    kw = dict(bin=1, break_thresh=30, coronavirus=0, include_start=1, record_depth=1, num_sources=2, rna=[1, 1])
    e = engine.Engine(engine.make_config(**kw))
    try:
        for seed, n in ((1, 6000), (2, 4000)):
            b = random_batch(np.random.default_rng(seed), n, len(g), g, num_sources=2, weird=0.0, max_ops=12, big_n=400)
            e.begin_chrom(3, g, ann)
            e.submit(b)
            e.end_chrom()
            dev = e.fetch()
            e.release()
            ora = oracle.run(oracle.make_config(**kw), g, ann, [b], 3)
            assert_same(dev, ora)
            assert dev["n_clusters"] > 2000
    finally:
        e.close()
    with pytest.raises(engine.NptError) as err:
        engine.run(engine.make_config(**dict(kw, max_coverage_clusters=256)), g, ann, [b], 3)
    assert err.value.code == -5 and "coverage rows" in str(err.value)
