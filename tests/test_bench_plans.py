"""bench.py --config 1|3|4|5: every plan builds on the CPU box, its generator makes reads of the stated shape, and the oracle
clusters a small sample of them with the plan's flag set (what bench_configs.run checks the device against on the GPU box)."""
import types

import numpy as np
import pytest

import bench_configs
from oracle import oracle


@pytest.mark.parametrize("no,world,rank", [(1, 1, 0), (3, 1, 0), (3, 8, 5), (4, 2, 1), (5, 1, 0), (5, 8, 3)])
def test_plan_builds_and_its_reads_cluster(no, world, rank):
    args = types.SimpleNamespace(reads=0, host_scale=0.01, config=no)
    P = bench_configs._plan(no, rank, world, args)
    assert P["chroms"], "every rank of these plans owns at least one chromosome"
    ch = P["chroms"][0]
    b = ch["gen"](1500, ch["first"])
    assert b.n == 1500 and len(b.cigar_off) == 1501 and int(b.seq_off[-1]) == len(b.seq4)
    assert int(b.src.max()) < P["cfg_kw"]["num_sources"]
    if no == 3 and world == 8:
        assert (b.src == rank).all()          # GPU k holds BAM k
    if no == 4:
        assert P["scaling"] == "strong" and P["waves"] * ch["n"] * world <= 500_000_000
    if no == 5:
        assert not P["collective"] and all(c["index"] % world == rank for c in P["chroms"])
    res = oracle.run(oracle.make_config(**P["cfg_kw"]), ch["genome"], ch["annotation"], [b], ch["index"])
    assert res["n_clusters"] > 0 and (res["cluster_id"] >= 0).mean() > 0.95
