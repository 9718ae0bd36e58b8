"""The lane logic of the fast walk (nptranscript_b200/csrc/fastwalk.cuh) compiled for the host and checked against the
oracle on the CPU box: per-read break lists / read offsets / sums, and the coverage obtained by summing the reads' entries
and junction points per (cluster, source).  The serial device walker takes the reads the fast walk refuses; this test
runs the oracle on exactly the reads the fast walk accepted."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nptranscript_b200 import synth, workloads
from nptranscript_b200.synth import Batch
from oracle import oracle
from tests.fuzz import random_batch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libhostfastwalk.so")
RARE_WORDS, BRK_INLINE = 16, 6


def _lib():
    src = os.path.join(HERE, "host_fastwalk.cpp")
    deps = [src] + [os.path.join(ROOT, "nptranscript_b200", "csrc", f) for f in ("fastwalk.cuh", "npt_device.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        tmp = SO + f".{os.getpid()}.tmp"
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", tmp], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, SO)
    L = C.CDLL(SO)
    L.hfw_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong] + [C.c_void_p] * 19 + [C.c_uint]
    return L


def ref_grid(codes):
    """word W = positions 8W..8W+7 (1-based; position 0 is padding), first position in the top nibble, 4 words of padding"""
    G = len(codes)
    nw = (G >> 3) + 1 + 4
    nib = np.zeros(nw * 8, np.uint32)
    nib[1:G + 1] = codes & 0xF
    sh = (28 - 4 * (np.arange(nw * 8) & 7)).astype(np.uint32)
    return (nib << sh).reshape(nw, 8).sum(axis=1).astype(np.uint32)


def run_host(b, g, T, record, break_cap=0):
    L = _lib()
    n = b.n
    refg = ref_grid(g)
    out = dict(nbreaks=np.zeros(n, np.int32), arena=np.zeros(n * BRK_INLINE + break_cap, np.int32), read_st=np.zeros(n, np.int32),
               read_en=np.zeros(n, np.int32), maps=np.zeros(n, np.int32), errs=np.zeros(n, np.int32), rflags=np.zeros(n, np.uint32),
               bloc=np.zeros(n, np.uint32), fhdrA=np.zeros(n * 4, np.uint32), fmeta=np.zeros(n, np.uint32),
               frare=np.zeros(n * RARE_WORDS, np.uint32),
               fent=np.zeros(((int(b.seq_off[-1]) >> 4) + 8 * n + 8) * 4, np.uint32), slow=np.zeros(n + 1, np.uint32), bctr=np.zeros(8, np.uint32))
    seq4 = np.ascontiguousarray(b.seq4)
    cig = np.ascontiguousarray(b.cigar)
    nslow = L.hfw_run(T, len(g), record, refg.ctypes.data, len(refg), n, b.pos.ctypes.data, b.cigar_off.ctypes.data,
                      cig.ctypes.data, b.seq_off.ctypes.data, seq4.ctypes.data, out["nbreaks"].ctypes.data, out["arena"].ctypes.data,
                      out["read_st"].ctypes.data, out["read_en"].ctypes.data, out["maps"].ctypes.data, out["errs"].ctypes.data,
                      out["rflags"].ctypes.data, out["bloc"].ctypes.data, out["fhdrA"].ctypes.data, out["fmeta"].ctypes.data, out["frare"].ctypes.data,
                      out["fent"].ctypes.data,
                      out["slow"].ctypes.data, out["bctr"].ctypes.data, break_cap)
    out["nslow"] = nslow
    return out


def subset(b, keep):
    idx = np.flatnonzero(keep)
    co, so = b.cigar_off.astype(np.int64), b.seq_off.astype(np.int64)
    cig = np.concatenate([b.cigar[co[i]:co[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint32)
    seq = np.concatenate([b.seq4[so[i]:so[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
    ncoff = np.concatenate([[0], np.cumsum(co[idx + 1] - co[idx])]).astype(np.uint32)
    nsoff = np.concatenate([[0], np.cumsum(so[idx + 1] - so[idx])]).astype(np.uint64)
    return Batch(b.pos[idx].copy(), b.flag[idx].copy(), b.src[idx].copy(), ncoff, cig.astype(np.uint32), nsoff, seq.astype(np.uint8))


def coverage_from_entries(out, b, ora, G, S):
    """what cov_key_kernel + accumulate_kernel compute, in numpy"""
    nc, stride = ora["n_clusters"], G + 2
    cov = {k: np.zeros((nc, S, stride), np.int64) for k in ("depth", "errors", "map_start", "map_end")}
    fhdrA, fmeta, frare = out["fhdrA"].reshape(-1, 4), out["fmeta"], out["frare"].reshape(-1, RARE_WORDS)
    fent = out["fent"].reshape(-1, 4)
    sh = (28 - 4 * np.arange(8)).astype(np.uint32)
    for r in range(b.n):
        c, s = int(ora["cluster_id"][r]), int(b.src[r])
        if c < 0:
            continue
        h, meta, rare = fhdrA[r], int(fmeta[r]), frare[r]
        nblk, njunc, nextra = meta & 0xFF, (meta >> 8) & 0xFF, (meta >> 16) & 0xFF
        e0 = int(h[0])
        assert all(int(h[1 + k]) == 0 for k in range(nblk, 3))
        for k in range(nblk):
            first, cnt = int(h[1 + k]) & 0xFFFF, int(h[1 + k]) >> 16
            words = fent[e0:e0 + cnt].reshape(-1)               # 8 positions per word
            nib = (words[:, None] >> sh[None, :]) & 0xF          # [word][nibble from the top]
            flags = nib.reshape(-1)
            p0 = first * 32
            m = min(len(flags), stride - p0)
            covb, qrk, mis = flags[:m] & 1, (flags[:m] >> 1) & 1, (flags[:m] >> 3) & 1
            assert not flags[m:].any() and not (flags & 4).any()
            cov["depth"][c, s, p0:p0 + m] += covb + qrk
            cov["errors"][c, s, p0:p0 + m] += mis + qrk
            e0 += cnt
        cov["map_start"][c, s, ora["start_pos"][r]] += 1
        cov["map_end"][c, s, ora["end_pos"][r]] += 1
        for j in range(njunc):
            cov["map_end"][c, s, int(rare[4 + 2 * j])] += 1
            cov["map_start"][c, s, int(rare[5 + 2 * j])] += 1
        for d in range((meta >> 24) & 0xFF):
            cov["depth"][c, s, int(rare[d])] += 1
            cov["errors"][c, s, int(rare[d])] += 1
        for e in range(nextra):
            a, p = int(rare[8 + 2 * e]), int(rare[9 + 2 * e])
            if a & 0x80000000:
                cov["depth"][c, s, a & 0x7FFFFFFF] += 1
                cov["errors"][c, s, a & 0x7FFFFFFF] += 1
            else:
                cov["map_start"][c, s, a] += 1
                cov["map_end"][c, s, p] += 1
    return cov


def check(b, g, w, T, S=1, min_fast=0.5, **cfg_kw):
    first = run_host(b, g, T, 1)
    slow = (first["rflags"] & 0x20) != 0
    assert first["nslow"] == slow.sum()
    assert set(first["slow"][:first["nslow"]].tolist()) == set(np.flatnonzero(slow).tolist())
    assert (~slow).mean() >= min_fast, f"only {(~slow).mean():.3f} of the reads took the fast walk"
    fb = subset(b, ~slow)
    out = run_host(fb, g, T, 1)
    assert out["nslow"] == 0
    kw = dict(bin=10, break_thresh=T, record_depth=1, num_sources=S)
    kw.update(cfg_kw)
    ora = oracle.run(oracle.make_config(**kw), g, w.annotation, [fb])
    nb = np.diff(ora["break_off"].astype(np.int64))
    # oracle reads flagged bad (0x10) never reach here: the fast walk hands those to the serial path
    assert not (ora["read_flags"] & 0x10).any()
    assert np.array_equal(out["nbreaks"], nb)
    arena = out["arena"].reshape(-1, BRK_INLINE)
    for r in range(fb.n):
        o = int(ora["break_off"][r])
        assert np.array_equal(arena[r, :nb[r]], ora["breaks"][o:o + nb[r]]), (r, arena[r], ora["breaks"][o:o + nb[r]])
    for k, f in (("read_st", "read_st"), ("read_en", "read_en"), ("maps", "maps_sum"), ("errs", "err_sum")):
        bad = np.flatnonzero(out[k] != ora[f])
        assert len(bad) == 0, (f, bad[:5], out[k][bad[:5]], ora[f][bad[:5]])
    if kw.get("coronavirus", 1):   # host mode trims the first / last exon: start / end points then come from cluster_kernel
        cov = coverage_from_entries(out, fb, ora, len(g), S)
        for k in ("depth", "errors", "map_start", "map_end"):
            bad = np.argwhere(cov[k] != ora[k])
            assert len(bad) == 0, (k, bad[:5].tolist(), [int(cov[k][tuple(x)]) for x in bad[:5]], [int(ora[k][tuple(x)]) for x in bad[:5]])
    # without depth recording the walk writes the same breaks and read offsets
    out0 = run_host(fb, g, T, 0)
    assert out0["nslow"] == 0 and np.array_equal(out0["nbreaks"], nb) and np.array_equal(out0["arena"], out["arena"])
    assert np.array_equal(out0["read_st"], out["read_st"]) and np.array_equal(out0["read_en"], out["read_en"])
    return float((~slow).mean())


@pytest.mark.parametrize("T", [1000, 100, 16])
def test_synthetic_reads(T):
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200301, 3000)
    frac = check(b, g, w, T, min_fast=0.95)
    assert frac > 0.95


def test_multi_source():
    w = workloads.HCOV229E
    g = synth.genome_codes(w)
    b = synth.generate(w, 20200310, 2000, num_sources=3)
    check(b, g, w, 1000, S=3, min_fast=0.95)


@pytest.mark.parametrize("seed,T", [(1, 1000), (2, 50), (3, 200), (4, 16)])
def test_fuzzed_cigars(seed, T):
    """every op, zero-length ops, long mismatch runs, alignments past the reference end: whatever the fast walk keeps must
    be exact, and it must hand over everything it cannot restate"""
    w = workloads.SARS2
    g = synth.genome_codes(w)
    rng = np.random.default_rng(seed)
    b = random_batch(rng, 1500, w.genome_len, g, num_sources=2, weird=0.05, max_ops=40, long_reads=(seed == 1))
    check(b, g, w, T, S=2, min_fast=0.05)


def test_known_answers():
    """SURVEY §8(c) KAT1-KAT3/KAT5 on the all-A reference through the fast walk (KAT4 uses =/X: serial path)"""
    w = workloads.SARS2
    G = w.genome_len
    g = np.full(G, 1, np.uint8)
    A = "A"
    def kat2(after):
        k = list(A * 1665)
        for off in (63, 64) + after:
            k[off] = "C"
        return "".join(k)
    reads = [
        (5, 0, 0, "10S65M28190N1600M30S", A * 1705),
        (5, 0, 0, "65M28190N1600M", kat2((65, 66))),
        (100, 0, 0, "5M2D5M", A * 10),
        (100, 0, 0, "10M2000N10M", A * 10 + "CC" + A * 8),
    ]
    b = synth.from_reads(reads)
    out = run_host(b, g, 1000, 1)
    assert out["nslow"] == 0
    arena = out["arena"].reshape(-1, BRK_INLINE)
    assert arena[0, :4].tolist() == [5, 69, 28260, 29859] and out["read_st"][0] == 11 and out["read_en"][0] == 1675
    assert arena[1, :4].tolist() == [5, 67, 28262, 29859]
    assert arena[2, :2].tolist() == [100, 111]
    assert arena[3, :4].tolist() == [100, 109, 2112, 2119]
    # KAT2 proper: three mismatching emissions right after the junction (three of the four extras a header holds)
    k2 = run_host(synth.from_reads([(5, 0, 0, "65M28190N1600M", kat2((65, 66, 67)))]), g, 1000, 1)
    assert k2["nslow"] == 0 and k2["arena"][:4].tolist() == [5, 67, 28263, 29859] and (int(k2["fmeta"][0]) >> 16) & 0xFF == 3
    # five of them are more than a header holds -> serial path
    assert run_host(synth.from_reads([(5, 0, 0, "65M28190N1600M", kat2((65, 66, 67, 68, 69)))]), g, 1000, 1)["nslow"] == 1
    check(synth.from_reads([(5, 0, 0, "65M28190N1600M", kat2((65, 66, 67)))]), g, w, 1000, min_fast=1.0)
    check(b, g, w, 1000, min_fast=1.0)


def test_long_break_lists_without_depth_recording():
    """host-mode shape: spliced reads with many junctions.  Without depth recording the fast walk keeps them (any number of
    aligned blocks, up to 127 junctions) and moves lists longer than the inline slot to the overflow area; with a too small
    overflow area it says so (bctr[1]) and reports the need (bctr[0])"""
    c = workloads.human_like()[20]
    g = synth.host_genome_codes(c)
    b = synth.generate_host(c, g, 9, 3000, num_sources=2, transcriptome=True)
    T = 50
    ora = oracle.run(oracle.make_config(bin=10, break_thresh=T, record_depth=0, num_sources=2), g, workloads.SARS2.annotation, [b])
    nb = np.diff(ora["break_off"].astype(np.int64))
    assert (nb > BRK_INLINE).mean() > 0.3
    cap = int(nb[nb > BRK_INLINE].sum())
    out = run_host(b, g, T, 0, break_cap=cap)
    slow = (out["rflags"] & 0x20) != 0
    assert slow.mean() < 0.05 and out["bctr"][1] == 0 and out["bctr"][0] == nb[(nb > BRK_INLINE) & ~slow].sum()
    for r in np.flatnonzero(~slow):
        o = int(ora["break_off"][r])
        want = ora["breaks"][o:o + nb[r]]
        assert out["nbreaks"][r] == nb[r]
        if nb[r] <= BRK_INLINE:
            assert out["bloc"][r] == 0xFFFFFFFF and np.array_equal(out["arena"][r * BRK_INLINE:r * BRK_INLINE + nb[r]], want)
        else:
            at = b.n * BRK_INLINE + int(out["bloc"][r])
            assert np.array_equal(out["arena"][at:at + nb[r]], want), r
    for k, f in (("read_st", "read_st"), ("read_en", "read_en")):
        assert np.array_equal(out[k][~slow], ora[f][~slow]), f
    small = run_host(b, g, T, 0, break_cap=cap // 2)
    assert small["bctr"][1] == 1 and small["bctr"][0] >= cap - nb[(nb > BRK_INLINE) & slow].sum()
    # with depth recording those reads still go to the serial walker
    rec = run_host(b, g, T, 1)
    assert np.array_equal((rec["rflags"] & 0x20) != 0, (nb > BRK_INLINE) | slow | ((rec["rflags"] & 0x20) != 0)) and ((rec["rflags"] & 0x20) != 0)[nb > BRK_INLINE].all()
