"""Frozen answers (tests/golden/cases.json, written by tests/golden/make_golden.py from the pure-Python restatement):
oracle.cpp on the CPU and the CUDA path on the GPU must both reproduce them — strings included, through render.py."""
import json
import os

import numpy as np
import pytest

from nptranscript_b200 import abi, render, synth, workloads
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "cases.json")))["cases"]


def _digest(rows, full):
    """rows: int array [S, stride] of one cluster -> the fixture's form."""
    s, p = np.nonzero(rows)
    v = rows[s, p].astype(np.int64)
    if full:
        return sorted([int(a), int(b), int(c)] for a, b, c in zip(s, p, v))
    M = (1 << 61) - 1
    return [len(v), int(v.sum()), int(sum((int(a) + 1) * (int(b) + 7) * int(c) for a, b, c in zip(s, p, v)) % M)]


def _check(case, res):
    ann = workloads.Annotation(kind=case["annotation"]["kind"], genes=case["annotation"]["genes"], start=case["annotation"]["start"],
                               end=case["annotation"]["end"], strand=case["annotation"]["strand"])
    kw, ci = case["flags"], case["chrom_index"]
    corona = bool(kw.get("coronavirus", 1))
    reads = case["reads"]
    names = [f"read{i}" for i in range(len(reads))]
    srcs = [r[2] for r in reads]
    keys = render.cluster_keys(res, ann, ci, coronavirus=corona)
    assert keys == [c["key"] for c in case["clusters"]]
    assert [render.cluster_id(ci, c) for c in range(res["n_clusters"])] == [c["id"] for c in case["clusters"]]
    subn = render.sub_names(res, srcs, names, ci, bool(kw.get("first_is_transcriptome", 0)))
    bo = res["break_off"]
    for i, row in enumerate(case["rows"]):
        if row is None:
            assert int(res["read_flags"][i]) & abi.NPT_RF_BADCIGAR
            continue
        f = int(res["read_flags"][i])
        got = dict(cluster=render.cluster_id(ci, int(res["cluster_id"][i])), sub=subn[i], start_pos=int(res["start_pos"][i]), end_pos=int(res["end_pos"][i]),
                   read_st=int(res["read_st"][i]), read_en=int(res["read_en"][i]), type_ind=f & 3, leader=bool(f & 4), neg=bool(f & 8),
                   breaks=res["breaks"][int(bo[i]):int(bo[i + 1])].tolist(), maps_sum=int(res["maps_sum"][i]), err_sum=int(res["err_sum"][i]))
        assert got == row, (i, got, row)
    assert res["total_counts"].tolist() == case["total_counts"]
    assert res["spliced"].tolist() == case["spliced"] and res["unspliced"].tolist() == case["unspliced"]
    got_pairs = sorted([int(a), int(b), int(c), int(d), int(e)] for a, b, c, d, e in
                       zip(res["pair_src"], res["pair_level"], res["pair_start"], res["pair_end"], res["pair_count"]))
    assert got_pairs == case["pairs"]
    so, ko = res["cl_sub_off"], res["sub_key_off"]
    full = case["name"].startswith("kats")
    for c, cl in enumerate(case["clusters"]):
        assert res["cl_read_count"][c].tolist() == cl["read_count"]
        assert int(so[c + 1]) - int(so[c]) == len(cl["subs"])
        for k, sub in enumerate(cl["subs"]):
            g = int(so[c]) + k
            assert res["sub_key"][int(ko[g]):int(ko[g + 1])].tolist() == sub["key"]
            assert res["sub_count"][g].tolist() == sub["count"] and int(res["sub_break_sum"][g]) == sub["break_sum"]
        if kw.get("record_depth"):
            for name in ("depth", "errors", "map_start", "map_end"):
                assert _digest(res[name][c], full) == cl[name], (case["name"], c, name)


def _inputs(case):
    w = workloads.SARS2
    g = np.full(w.genome_len, 1, np.uint8) if case["name"].startswith("kats") else synth.genome_codes(w)
    a = case["annotation"]
    ann = workloads.Annotation(kind=a["kind"], genes=a["genes"], start=a["start"], end=a["end"], strand=a["strand"])
    b = synth.from_reads([tuple(r) for r in case["reads"]])
    return g, ann, b


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_golden(case):
    g, ann, b = _inputs(case)
    res = oracle.run(oracle.make_config(**case["flags"]), g, ann, [b], case["chrom_index"])
    _check(case, res)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_reproduces_golden(case):
    from nptranscript_b200 import engine
    g, ann, b = _inputs(case)
    res = engine.run(engine.make_config(max_coverage_clusters=256, **case["flags"]), g, ann, [b], case["chrom_index"])
    _check(case, res)
