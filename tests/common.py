"""Shared helpers: field-by-field comparison of device results with the oracle."""
import numpy as np

READ_FIELDS = ["cluster_id", "sub_id", "start_pos", "end_pos", "read_st", "read_en", "maps_sum", "err_sum", "break_off", "breaks"]
TABLE_FIELDS = ["cl_first_read", "cl_key_off", "cl_key_tok", "cl_read_count", "cl_sub_off", "sub_first_read", "sub_key_off",
                "sub_key", "sub_count", "sub_break_sum", "sub_sum_off", "sub_sums", "sub_flags", "total_counts", "spliced",
                "unspliced", "pair_src", "pair_level", "pair_start", "pair_end", "pair_count"]
COV_FIELDS = ["depth", "errors", "map_start", "map_end"]
SLOW = 0x20


def assert_same(dev, ora, coverage=True, what="all"):
    assert dev["n_reads"] == ora["n_reads"]
    assert dev["n_clusters"] == ora["n_clusters"], (dev["n_clusters"], ora["n_clusters"])
    assert dev["n_subs"] == ora["n_subs"], (dev["n_subs"], ora["n_subs"])
    for k in READ_FIELDS + TABLE_FIELDS:
        a, b = np.asarray(dev[k]), np.asarray(ora[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a.ravel() != b.ravel())
            raise AssertionError(f"{k}: {len(bad)} mismatches, first at {bad[:5]}: dev={a.ravel()[bad[:5]]} oracle={b.ravel()[bad[:5]]}")
    fa = np.asarray(dev["read_flags"]) & ~np.uint32(SLOW)  # the slow-path bit is a diagnostic, not a result
    assert np.array_equal(fa, np.asarray(ora["read_flags"])), "read_flags"
    if coverage and "depth" in ora:
        for k in COV_FIELDS:
            a, b = dev[k], ora[k]
            assert a.shape == b.shape, (k, a.shape, b.shape)
            if not np.array_equal(a, b):
                bad = np.argwhere(a != b)
                raise AssertionError(f"{k}: {len(bad)} mismatches, first {bad[:5].tolist()} dev={a[tuple(bad[0])]} oracle={b[tuple(bad[0])]}")
