"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded reads.  Bit-exact."""
import numpy as np
import pytest

from nptranscript_b200 import abi, engine, synth, workloads
from oracle import oracle
from tests.common import assert_same

pytestmark = pytest.mark.gpu


def _both(cfg_kw, w, batches, annotation=None, chrom_index=0):
    g = synth.genome_codes(w)
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
    assert dev["n_slow_reads"] < 20000 // 200


def test_multi_batch_and_sources():
    w = workloads.HCOV229E
    bs = [synth.generate(w, 20200310 + k, 5000, first_index=0, num_sources=4, src_fixed=k) for k in range(4)]
    dev, ora = _both(dict(bin=10, break_thresh=1000, record_depth=1, num_sources=4), w, bs)
    assert_same(dev, ora)
