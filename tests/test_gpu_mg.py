"""npt_mg_* — the in-process multi-device context behind the C ABI (csrc/comm.inl) — on ONE GPU: a device listed several
times gives several contexts whose collectives run as device-to-device copies and an add kernel; the submit split, the
key exchange, the global numbering, the summed counters and the re-interleaved per-read results must equal the oracle on
the unsharded input.  tests/mg_check.py runs the same check over NCCL on 2, 4 or 8 real devices."""
import numpy as np
import pytest

from nptranscript_b200 import engine, synth, workloads
from oracle import oracle
from tests.common import assert_same
from tests.fuzz import random_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ndev", [1, 2, 3])
def test_mg_virus_mode_several_batches(ndev):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    sizes = [9000, 1, 4000, 2]   # batches smaller than the device count leave devices without reads
    batches, first = [], 0
    for i, n in enumerate(sizes):
        batches.append(synth.generate(w, 400 + i, n, first_index=first, num_sources=2))
        first += n
    kw = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2, max_coverage_clusters=256)
    dev = engine.run_mg(engine.make_config(**kw), [0] * ndev, g, w.annotation, batches)
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, batches)
    assert_same(dev, ora)


def test_mg_host_mode_first_is_transcriptome():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(8)
    batches = [random_batch(rng, n, w.genome_len, g, num_sources=2, big_n=300) for n in (2500, 3500)]
    kw = dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1], first_is_transcriptome=1, num_sources=2)
    ann = workloads.Annotation.empty()
    dev = engine.run_mg(engine.make_config(**kw), [0, 0], g, ann, batches, chrom_index=2)
    ora = oracle.run(oracle.make_config(**kw), g, ann, batches, 2)
    assert_same(dev, ora)


def test_mg_two_chromosomes_on_one_context():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    kw = dict(bin=100, break_thresh=1000, record_depth=1)
    e = engine.MultiEngine(engine.make_config(**kw), [0, 0])
    try:
        for seed, n in ((1, 5000), (2, 3000)):
            b = synth.generate(w, seed, n)
            e.begin_chrom(0, g, w.annotation)
            e.submit(b)
            e.end_chrom()
            dev = e.fetch()
            e.release()
            assert_same(dev, oracle.run(oracle.make_config(**kw), g, w.annotation, [b]))
    finally:
        e.close()


def test_mg_packed_batches_split_on_even_bytes():
    """npt_mg_submit_packed: shares of a packed batch start on an even seq4 byte; wide-element and exception lists are cut and
    rebased per share"""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    batches = [synth.generate(w, 500, 7001, num_sources=2), synth.generate(w, 501, 3, first_index=7001, num_sources=2)]
    rng = np.random.default_rng(2)
    k = rng.integers(0, len(batches[0].seq4), 300)
    batches[0].seq4[k] = rng.integers(0, 256, 300).astype(np.uint8)   # non-ACGT codes -> exceptions in every share
    kw = dict(bin=100, break_thresh=1000, record_depth=1, num_sources=2, max_coverage_clusters=256)
    dev = engine.run_mg(engine.make_config(**kw), [0, 0, 0], g, w.annotation, batches, packed=True)
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, batches)
    assert_same(dev, ora)


def test_mg_a_device_without_reads_and_an_error_that_does_not_hang():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    one = synth.generate(w, 9, 1)
    kw = dict(bin=100, break_thresh=1000, record_depth=1)
    dev = engine.run_mg(engine.make_config(**kw), [0, 0], g, w.annotation, [one])   # the second context never sees a read
    assert_same(dev, oracle.run(oracle.make_config(**kw), g, w.annotation, [one]))
    # a capacity error on the contexts must come back as an error from npt_mg_end_chrom, not as a hang of the group
    b = synth.generate(w, 10, 4000)
    with pytest.raises(engine.NptError):
        engine.run_mg(engine.make_config(**dict(kw, bin=1, max_subclusters=16)), [0, 0], g, w.annotation, [b])
