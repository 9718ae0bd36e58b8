"""The multi-rank protocol on ONE GPU: two contexts stand in for two ranks, keys travel by device pointer instead of NCCL,
counters are summed on the host.  Global ids, tables and coverage must equal the oracle on the unsharded input."""
import numpy as np
import pytest

from nptranscript_b200 import abi, engine, synth, workloads
from oracle import oracle
from tests.common import assert_same
from tests.fuzz import random_batch

pytestmark = pytest.mark.gpu

SUMMED = ("cl_read_count", "sub_count", "sub_break_sum", "sub_sums", "pair_count", "total_counts", "spliced", "unspliced", "depth", "errors",
          "map_start", "map_end")


def _two_ranks(kw, w, g, ann, full, cut, chrom_index=0):
    parts = [full.slice(0, cut), full.slice(cut, full.n)]
    engs = [engine.Engine(engine.make_config(**kw)) for _ in parts]
    try:
        for e, b, base in zip(engs, parts, (0, cut)):
            e.set_read_index_base(base)
            e.begin_chrom(chrom_index, g, ann)
            e.submit(b)
        bufs = [e.export_keys() for e in engs]
        engs[0].import_keys(*bufs[1])
        engs[1].import_keys(*bufs[0])
        res = []
        for e in engs:
            e.end_chrom()
            res.append(e.fetch())
        return res
    finally:
        for e in engs:
            e.close()


def _check(kw, w, ann, full, cut, chrom_index=0):
    g = synth.genome_codes(w)
    r0, r1 = _two_ranks(kw, w, g, ann, full, cut, chrom_index)
    ora = oracle.run(oracle.make_config(**kw), g, ann, [full], chrom_index)
    merged = {}
    for k, v in r0.items():
        if k in SUMMED and k in r1:
            merged[k] = v + r1[k]  # what the all-reduce does
        elif isinstance(v, np.ndarray) and k in ("cluster_id", "sub_id", "start_pos", "end_pos", "read_st", "read_en", "read_flags", "maps_sum", "err_sum"):
            merged[k] = np.concatenate([v, r1[k]])
        elif k == "breaks":
            merged[k] = np.concatenate([v, r1[k]])
        elif k == "break_off":
            merged[k] = np.concatenate([v, r1[k][1:] + v[-1]])
        else:
            merged[k] = v
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, r1[k]), f"{k} differs between the ranks"  # keys, first indices, offsets: identical everywhere
    merged["n_reads"] = r0["n_reads"] + r1["n_reads"]
    # cluster read counts are derived from the (not yet summed) sub counts at fetch time: recompute from the summed ones
    so = merged["cl_sub_off"].astype(np.int64)
    if merged["n_clusters"]:
        merged["cl_read_count"] = np.stack([merged["sub_count"][so[c]:so[c + 1]].sum(axis=0) for c in range(merged["n_clusters"])])
    assert_same(merged, ora)


@pytest.mark.parametrize("cut", [1, 7000, 19999])
def test_virus_mode_two_ranks(cut):
    w = workloads.SARS2
    full = synth.generate(w, 91, 20000, num_sources=2)
    _check(dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2, max_coverage_clusters=256), w, w.annotation, full, cut)


def test_host_mode_first_is_transcriptome_two_ranks():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(3)
    full = random_batch(rng, 6000, w.genome_len, g, num_sources=2, big_n=300)
    kw = dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, record_depth=0, rna=[1, 1], first_is_transcriptome=1, num_sources=2)
    _check(kw, w, workloads.Annotation.empty(), full, 2500, chrom_index=2)
