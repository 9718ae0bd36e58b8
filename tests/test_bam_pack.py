"""libnptbam (BAM -> SoA packer + the reference's record filter chain) against an independent pure-Python BAM reader and
restatement of the chain, on BAM files written here from synthetic reads."""
import os

import numpy as np
import pytest

from nptranscript_b200 import bam, synth, workloads
from oracle import py_bam


def _case(tmp_path, n=3000, seed=5, block_bytes=0xFF00):
    w = workloads.SARS2
    rng = np.random.default_rng(seed)
    b0 = synth.generate(w, 100 + seed, n)
    b1 = synth.generate(workloads.HCOV229E, 200 + seed, n // 2)
    recs = []
    for ref_id, b in ((0, b0), (1, b1), (0, b0.slice(0, 50))):   # a third run goes back to reference 0
        rs = bam.records_from_batch(b, ref_id=ref_id)
        for i, r in enumerate(rs):
            u = rng.random()
            r["name"] = f"r{ref_id}_{len(recs) + i}"
            r["mapq"] = int(rng.integers(0, 61))
            if u < 0.05:
                r["flag"] |= 0x4
                r["ref_id"] = -1 if rng.random() < 0.5 else r["ref_id"]
            elif u < 0.10:
                r["flag"] |= 0x100
            elif u < 0.13:
                r["flag"] |= 0x800
            v = rng.random()
            if v < 0.1:
                r["qual"] = None
            elif v < 0.2:
                r["qual"] = bytes(rng.integers(0, 6, r["l_seq"]).astype(np.uint8))      # low quality
            else:
                r["qual"] = bytes(rng.integers(5, 40, r["l_seq"]).astype(np.uint8))
        recs += rs
    # a one-base read and a read with a 200-element CIGAR on purpose
    recs.insert(7, dict(ref_id=0, pos=5, flag=0, mapq=60, name="tiny", cigar=[(1 << 4) | 0], seq4=b"\x10", l_seq=1, qual=b"\x1e"))
    path = os.path.join(tmp_path, "t.bam")
    refs = [(w.name, w.genome_len), ("hcov229e", workloads.HCOV229E.genome_len)]
    bam.write_bam(path, refs, recs, block_bytes=block_bytes)
    return path, refs, recs


def _expected_batches(path, **flt):
    refs, recs = py_bam.read_bam(path)
    stream, counts = py_bam.filter_chain(recs, **flt)
    batches, cur = [], None
    for r, q1, kept in stream:
        if cur is None or (r["ref_id"] != cur["ref_id"] and cur["recs"]):
            cur = dict(ref_id=r["ref_id"], recs=[])
            batches.append(cur)
        elif r["ref_id"] != cur["ref_id"]:
            cur["ref_id"] = r["ref_id"]
        if kept:
            cur["recs"].append((r, q1))
    return refs, [b for b in batches if b["recs"]], counts


@pytest.mark.parametrize("flt,threads,block,max_batch", [
    (dict(), 1, 0xFF00, 1 << 20),
    (dict(fail_thresh=7.0, min_mapq=10), 4, 4000, 700),
    (dict(fail_thresh=7.0, min_mapq=0, max_reads=1234), 3, 0xFF00, 1 << 20),
    (dict(ref_include=[0, 1]), 2, 333, 1 << 20),
])
def test_packer_matches_python_reading(tmp_path, flt, threads, block, max_batch):
    path, refs, _ = _case(str(tmp_path), block_bytes=block)
    erefs, ebatches, ecounts = _expected_batches(path, **flt)
    rd = bam.Reader(path, nthreads=threads)
    assert rd.refs == erefs == refs
    got = list(rd.batches(src=3, max_batch_reads=max_batch, decode_names=True, **flt))
    # re-split the expectation at max_batch
    exp = []
    for b in ebatches:
        for o in range(0, len(b["recs"]), max_batch):
            exp.append((b["ref_id"], b["recs"][o:o + max_batch]))
    assert [(r, b.n) for r, b, _ in got] == [(r, len(rs)) for r, rs in exp]
    for (ref_id, b, extra), (_, rs) in zip(got, exp):
        assert b.pos.tolist() == [r["pos"] for r, _ in rs]
        assert b.flag.tolist() == [r["flag"] for r, _ in rs]
        assert (b.src == 3).all()
        assert extra["names"] == [r["name"] for r, _ in rs]
        assert extra["mapq"].tolist() == [r["mapq"] for r, _ in rs]
        assert np.array_equal(extra["q1"], np.array([q for _, q in rs]))  # same libm pow/log10, same summation order
        assert b.cigar.tolist() == [c for r, _ in rs for c in r["cigar"]]
        assert b.cigar_off.tolist() == np.concatenate([[0], np.cumsum([len(r["cigar"]) for r, _ in rs])]).tolist()
        assert b.seq4.tobytes() == b"".join(r["seq4"] for r, _ in rs)
        assert b.seq_off.tolist() == np.concatenate([[0], np.cumsum([len(r["seq4"]) for r, _ in rs])]).tolist()
    t = rd.totals
    assert (t["n_unmapped"], t["n_excluded_ref"], t["n_low_quality"], t["n_short"], t["n_low_mapq"], t["n_secondary"], bool(t["stopped_at_max_reads"])) == \
           (ecounts["unmapped"], ecounts["excluded"], ecounts["lowq"], ecounts["short"], ecounts["lowmapq"], ecounts["secondary"], ecounts["stopped"])
    rd.close()


def test_round_trip_of_a_synthetic_batch(tmp_path):
    """No filtering: what the generator made is what the packer hands to npt_submit (odd-length reads lose only their pad nibble)."""
    w = workloads.SARS2
    b = synth.generate(w, 77, 2000)
    path = os.path.join(str(tmp_path), "rt.bam")
    bam.write_bam(path, [(w.name, w.genome_len)], bam.records_from_batch(b))
    (ref_id, got, extra), = list(bam.Reader(path, nthreads=2).batches())
    assert ref_id == 0 and got.n == b.n
    for k in ("pos", "flag", "cigar_off", "cigar", "seq_off", "seq4"):
        assert np.array_equal(getattr(got, k), getattr(b, k)), k
    assert (extra["q1"] == 500).all()


@pytest.mark.parametrize("threads", [1, 5])
def test_compact_format_from_bam_equals_the_numpy_packer(tmp_path, threads):
    """nptb_pack (C++, on the reader's threads) and packed.pack (numpy) give the same arrays, and unpacking them gives back the
    batch; reads with N / IUPAC bases and junction-sized N elements take the exception and wide lists"""
    from nptranscript_b200 import abi, packed
    w = workloads.SARS2
    b = synth.generate(w, 78, 3000)
    rng = np.random.default_rng(1)
    k = rng.integers(0, len(b.seq4), 200)
    b.seq4[k] = rng.integers(0, 256, 200).astype(np.uint8)
    path = os.path.join(str(tmp_path), "pk.bam")
    bam.write_bam(path, [(w.name, w.genome_len)], bam.records_from_batch(b))
    (ref_id, got, extra), = list(bam.Reader(path, nthreads=threads).batches(packed=True))
    pb, want = extra["packed"], packed.pack(got)
    for x, y in zip(pb.arrays(), want.arrays()):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    cig, seq4 = packed.unpack(pb)
    assert np.array_equal(cig, got.cigar) and np.array_equal(seq4, got.seq4)
    assert len(pb.wide_idx) > 0 and len(pb.exc_byte) > 0 and pb.nbytes() < 0.55 * got.nbytes()
    # nptb_packed is npt_packed_batch field for field (a pointer to one is what npt_submit_packed takes)
    assert [(n, getattr(bam.PackedStruct, n).offset) for n, _ in bam.PackedStruct._fields_] == \
           [(n, getattr(abi.npt_packed_batch, n).offset) for n, _ in abi.npt_packed_batch._fields_]


def test_malformed_files_are_reported(tmp_path):
    p = os.path.join(str(tmp_path), "bad.bam")
    open(p, "wb").write(b"not a bam file at all, not even gzip..........")
    with pytest.raises(bam.BamError):
        bam.Reader(p)
    path, _, _ = _case(str(tmp_path), n=300)
    data = bytearray(open(path, "rb").read())
    data[len(data) // 2] ^= 0x55  # corrupt one compressed byte: inflate or the CRC must notice
    open(p, "wb").write(bytes(data))
    with pytest.raises(bam.BamError):
        list(bam.Reader(p).batches())
    open(p, "wb").write(bytes(open(path, "rb").read()[:-100]))  # truncated
    with pytest.raises(bam.BamError):
        list(bam.Reader(p).batches())


def test_library_exports_every_declared_symbol():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "npt_bam.h")).read()
    declared = set(re.findall(r"\b(nptb_[a-z_]+)\s*\(", hdr))
    assert declared == set(bam.SYMBOLS)
    L = bam.lib()
    for s in bam.SYMBOLS:
        assert hasattr(L, s)


def test_struct_layout_matches_c():
    import ctypes as C
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "npt_bam.h"\nint main(){\n'
    probes = []
    for cname, st in (("nptb_filter", bam.Filter), ("nptb_batch", bam.BatchStruct), ("nptb_packed", bam.PackedStruct)):
        src += f'printf("%zu\\n", sizeof({cname}));\n'
        probes.append((cname, None, C.sizeof(st)))
        for fname, _ in st._fields_:
            src += f'printf("%zu\\n", offsetof({cname}, {fname}));\n'
            probes.append((cname, fname, getattr(st, fname).offset))
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        out = subprocess.check_output([os.path.join(d, "p")], text=True).split()
    for (n, f, py), c in zip(probes, out):
        assert py == int(c), (n, f, py, c)
