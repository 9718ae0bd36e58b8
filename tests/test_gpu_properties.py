"""Size-independent properties of the CUDA path at a size the oracle would need minutes for (1 M reads, the read shape of
BASELINE.json configs[1]): conservation sums, first-appearance numbering, batch-split invariance.  The oracle is not used."""
import numpy as np
import pytest

from nptranscript_b200 import engine, synth, workloads

pytestmark = pytest.mark.gpu

N = 1_000_000
KW = dict(bin=100, break_thresh=1000, record_depth=1, calc_breaks=1, track_break_sums=1, num_sources=1, max_coverage_clusters=2048)


@pytest.fixture(scope="module")
def big():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200301, N)
    res = engine.run(engine.make_config(**KW), g, w.annotation, [b])
    return w, g, b, res


def test_conservation(big):
    w, g, b, res = big
    ok = res["cluster_id"] >= 0
    n_ok = int(ok.sum())
    assert n_ok == N  # the generator emits valid records only
    assert int(res["total_counts"].sum()) == n_ok
    assert int(res["cl_read_count"].sum()) == n_ok and int(res["sub_count"].sum()) == n_ok
    so = res["cl_sub_off"].astype(np.int64)
    per_cluster = np.add.reduceat(res["sub_count"].sum(axis=1), so[:-1])
    assert np.array_equal(per_cluster, res["cl_read_count"].sum(axis=1))
    assert int(res["sub_break_sum"].sum()) == n_ok  # every read adds its breaks once (Count.addBreaks)
    # every emitted position is one depth count; every non-matching one an error count
    assert int(res["depth"].sum(dtype=np.int64)) == int(res["maps_sum"].sum(dtype=np.int64))
    assert int(res["errors"].sum(dtype=np.int64)) == int(res["err_sum"].sum(dtype=np.int64))
    assert (res["depth"] >= res["errors"]).all() and (res["depth"] >= 0).all()
    # mapStart and mapEnd receive the same number of events (one pair per break_p position + setStartEnd)
    assert int(res["map_start"].sum(dtype=np.int64)) == int(res["map_end"].sum(dtype=np.int64)) >= n_ok
    # break lists: start and end of the alignment, strictly increasing, even length
    bo = res["break_off"].astype(np.int64)
    nb = np.diff(bo)
    assert (nb % 2 == 0).all() and (nb >= 2).all()
    assert np.array_equal(res["breaks"][bo[:-1]], res["start_pos"]) and np.array_equal(res["breaks"][bo[1:] - 1], res["end_pos"])
    d = np.diff(res["breaks"].astype(np.int64))
    inner = np.ones(len(d), bool)
    inner[bo[1:-1] - 1] = False
    assert (d[inner] >= 0).all()
    # sum of the break sums of a sub-cluster = sum over its reads, position by position (spot check: total)
    assert int(res["sub_sums"].sum(dtype=np.int64)) == int(res["breaks"].sum(dtype=np.int64))
    # break-point pairs: one per internal gap > T of each read (BreakPoints.addBreakPoint)
    k_in_read = np.arange(len(res["breaks"]), dtype=np.int64) - np.repeat(bo[:-1], nb)
    junction = (k_in_read[:-1] % 2 == 1) & inner  # between the end of one block and the start of the next
    assert int(res["pair_count"].sum()) == int((junction & (d > KW["break_thresh"])).sum())


def test_first_appearance_numbering(big):
    w, g, b, res = big
    cl = res["cluster_id"].astype(np.int64)
    first = np.full(res["n_clusters"], N, np.int64)
    np.minimum.at(first, cl, np.arange(N))
    assert np.array_equal(first, res["cl_first_read"])
    assert (np.diff(res["cl_first_read"]) > 0).all()  # "ID<chr>.<n>": n-th distinct key in arrival order
    so = res["cl_sub_off"].astype(np.int64)
    gsub = so[cl] + res["sub_id"].astype(np.int64)
    sfirst = np.full(res["n_subs"], N, np.int64)
    np.minimum.at(sfirst, gsub, np.arange(N))
    assert np.array_equal(sfirst, res["sub_first_read"])
    for c in range(res["n_clusters"]):
        assert (np.diff(res["sub_first_read"][so[c]:so[c + 1]]) > 0).all()  # ".t<k>": k-th distinct binned break list of the cluster


def test_batch_split_invariance(big):
    w, g, b, res = big
    cuts = [0, 1, 777, 300_000, 300_001, 999_999, N]
    parts = [b.slice(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    res2 = engine.run(engine.make_config(**KW), g, w.annotation, parts)
    for k, v in res.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, res2[k]), k
        elif k != "n_slow_reads" and not k.startswith("_"):
            assert v == res2[k], k
