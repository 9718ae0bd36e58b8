"""Compact transfer format (npt_packed_batch): the numpy packer round-trips on the CPU; on the GPU a packed submit gives the
same results as the expanded arrays and as the oracle."""
import numpy as np
import pytest

from nptranscript_b200 import packed, synth, workloads
from tests.fuzz import random_batch


def _roundtrip(b):
    pb = packed.pack(b)
    cig, seq4 = packed.unpack(pb)
    assert np.array_equal(cig, b.cigar)
    assert np.array_equal(seq4, b.seq4[: int(b.seq_off[-1])])
    return pb


def test_pack_unpack_identity_synthetic_reads():
    w = workloads.SARS2
    b = synth.generate(w, 7, 5000)
    pb = _roundtrip(b)
    # about half the bytes: 16-bit CIGAR words, 2-bit bases; only junction-sized N elements are wide
    assert pb.nbytes() < 0.53 * b.nbytes()
    assert 0 < len(pb.wide_idx) < 3 * b.n
    assert (pb.wide_word >> 4).min() >= 4095


def test_pack_unpack_identity_hand_built_reads():
    reads = [
        (10, 0, 0, "5S20M4095N7M1I3M", "ACGTN" + "ACGT" * 5 + "RYKMACG" + "T" + "=AC"),       # N, IUPAC, '=' and an escape at exactly 4095
        (1, 16, 0, "3M", "ACG"),                                                              # odd length: pad nibble is an exception
        (5, 0, 1, "4094M", "A" * 4094),                                                      # longest narrow element
        (5, 0, 1, "268435455M", "AC"),                                                       # 28-bit length
        (7, 0, 0, "", ""),                                                                   # no CIGAR, no bases
    ]
    b = synth.from_reads(reads)
    pb = _roundtrip(b)
    assert list(pb.wide_idx) == [2, 8]   # 5S 20M [4095N] 7M 1I 3M | 3M | 4094M | [268435455M]
    assert len(pb.exc_byte) >= 4


def test_pack_unpack_identity_fuzz():
    rng = np.random.default_rng(11)
    g = synth.genome_codes(workloads.SARS2)
    for _ in range(3):
        _roundtrip(random_batch(rng, 400, len(g), g, weird=0.5, bad_op=0.02, big_n=9000))


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [0, 1])
def test_packed_submit_equals_plain_and_oracle(depth):
    from nptranscript_b200 import engine
    from oracle import oracle
    from tests.common import assert_same
    w = workloads.SARS2
    g = synth.genome_codes(w)
    batches = [synth.generate(w, 31, 30000), synth.generate(w, 32, 1), synth.generate(w, 33, 12345, first_index=30001)]
    kw = dict(bin=100, break_thresh=1000, record_depth=depth)
    dev_p = engine.run(engine.make_config(**kw), g, w.annotation, batches, packed=True)
    dev = engine.run(engine.make_config(**kw), g, w.annotation, batches)
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, batches)
    assert_same(dev_p, ora)
    assert_same(dev_p, dev)


@pytest.mark.gpu
def test_packed_submit_fuzz_with_exceptions_and_wide_elements():
    from nptranscript_b200 import engine
    from oracle import oracle
    from tests.common import assert_same
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(5)
    batches = [random_batch(rng, 3000, len(g), g, weird=0.4, bad_op=0.01, big_n=9000) for _ in range(3)]
    for b in batches:   # non-ACGT codes in the reads
        k = rng.integers(0, len(b.seq4), 500)
        b.seq4[k] = rng.integers(0, 256, 500).astype(np.uint8)
    kw = dict(bin=10, break_thresh=50, record_depth=1)
    dev_p = engine.run(engine.make_config(**kw), g, w.annotation, batches, packed=True)
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, batches)
    assert_same(dev_p, ora)
