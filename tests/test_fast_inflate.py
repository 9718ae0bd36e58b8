"""bam/fast_inflate.h — the BAM packer's own DEFLATE decoder — against zlib: every stream it accepts must decode to exactly
what zlib gives; what it refuses (-1) is the packer's cue to use zlib, so refusing is always safe, corrupting never is."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from nptranscript_b200 import bam, synth, workloads


def _raw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return co.compress(data) + co.flush()


def _inflate(comp, n, mode):
    L = bam.lib()
    out = C.create_string_buffer(max(n, 1) + 64)
    got = L.nptb_inflate_raw(comp, len(comp), out, n, mode)
    return got, out.raw[:max(got, 0)]


def _payloads():
    rng = np.random.default_rng(0)
    w = workloads.SARS2
    b = synth.generate(w, 3, 400)
    yield "bam-like", b.seq4.tobytes()[:60000]
    yield "cigar words", b.cigar.tobytes()[:64000]
    yield "random", rng.integers(0, 256, 50000, dtype=np.uint8).tobytes()
    yield "two symbols", rng.integers(0, 2, 65536, dtype=np.uint8).tobytes()
    yield "zeros", bytes(65536)
    yield "one byte", b"x"
    yield "empty", b""
    yield "text", (b"the quick brown fox jumps over the lazy dog. " * 1500)[:65536]
    yield "long matches", (bytes(range(256)) * 300)[:65536]
    yield "skewed", rng.choice(np.arange(256, dtype=np.uint8), 65000, p=np.r_[[0.5], np.full(255, 0.5 / 255)]).tobytes()
    yield "short periods", b"".join(bytes([i % 7, (i * 3) % 5]) for i in range(30000))


@pytest.mark.parametrize("level", [0, 1, 3, 6, 9])
def test_agrees_with_zlib(level):
    accepted = 0
    for name, data in _payloads():
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            comp = _raw(data, level, strategy)
            n0, ref = _inflate(comp, len(data), 0)
            assert n0 == len(data) and ref == data
            n1, got = _inflate(comp, len(data), 1)
            assert n1 in (-1, len(data)), (name, level, strategy, n1)
            if n1 >= 0:
                assert got == data, (name, level, strategy)
                accepted += 1
    assert accepted > 20       # it takes the usual streams; refusing is allowed only for the odd ones


def test_damaged_and_truncated_streams_never_decode_to_wrong_bytes_silently():
    rng = np.random.default_rng(1)
    data = (b"ACGTTGCA" * 4000) + rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    comp = bytearray(_raw(data, 6))
    for cut in (1, 5, len(comp) // 2, len(comp) - 1):
        n1, got = _inflate(bytes(comp[:cut]), len(data), 1)
        assert n1 == -1 or got == data[:n1]     # a prefix at most, never garbage ... and the packer compares the length
        assert n1 != len(data)
    for trial in range(200):
        bad = bytearray(comp)
        k = int(rng.integers(0, len(bad)))
        bad[k] ^= 1 << int(rng.integers(0, 8))
        n1, got = _inflate(bytes(bad), len(data), 1)   # must not crash or write past the buffer; the CRC decides the rest
        assert -1 <= n1 <= len(data)
    # an output buffer that is too small is refused, not overrun
    n1, _ = _inflate(bytes(comp), len(data) - 10, 1)
    assert n1 == -1


def test_reader_gives_the_same_batches_with_and_without_it(tmp_path, monkeypatch):
    w = workloads.SARS2
    b = synth.generate(w, 5, 3000)
    path = os.path.join(str(tmp_path), "x.bam")
    bam.write_bam(path, [(w.name, w.genome_len)], bam.records_from_batch(b), level=6)
    (r1, fast, _), = list(bam.Reader(path, nthreads=3).batches())
    monkeypatch.setenv("NPTB_ZLIB_ONLY", "1")
    (r0, ref, _), = list(bam.Reader(path, nthreads=3).batches())
    for k in ("pos", "flag", "cigar_off", "cigar", "seq_off", "seq4"):
        assert np.array_equal(getattr(fast, k), getattr(ref, k)) and np.array_equal(getattr(fast, k), getattr(b, k)), k
