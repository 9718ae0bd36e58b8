"""BAM (BGZF) -> structure-of-arrays batches through libnptbam (include/npt_bam.h), plus a small BAM writer used by the
tests and the examples (there is no samtools/pysam in this environment).

The packer restates the record loop of J/run/ViralTranscriptAnalysisCmd2.java:712-830, 936-959 up to the call of
profile.identity1: which records reach the hot path, in which order, with which q1."""
import ctypes as C
import os
import struct
import zlib

import numpy as np

from . import _build
from .synth import Batch

_lib = None


class Filter(C.Structure):
    _fields_ = [("fail_thresh", C.c_double), ("min_mapq", C.c_int32), ("max_reads", C.c_int64), ("ref_include", C.POINTER(C.c_uint8))]


class BatchStruct(C.Structure):
    _fields_ = [("ref_id", C.c_int32), ("n", C.c_int64), ("pos", C.POINTER(C.c_int32)), ("flag", C.POINTER(C.c_uint16)),
                ("src", C.POINTER(C.c_uint16)), ("cigar_off", C.POINTER(C.c_uint32)), ("cigar", C.POINTER(C.c_uint32)),
                ("seq_off", C.POINTER(C.c_uint64)), ("seq4", C.POINTER(C.c_uint8)), ("q1", C.POINTER(C.c_double)),
                ("mapq", C.POINTER(C.c_uint8)), ("name_off", C.POINTER(C.c_uint32)), ("names", C.POINTER(C.c_char)),
                ("n_records", C.c_int64), ("n_unmapped", C.c_int64), ("n_excluded_ref", C.c_int64), ("n_low_quality", C.c_int64),
                ("n_short", C.c_int64), ("n_low_mapq", C.c_int64), ("n_secondary", C.c_int64), ("stopped_at_max_reads", C.c_int32)]


class PackedStruct(C.Structure):
    """nptb_packed = npt_packed_batch, field for field"""
    _fields_ = [("n", C.c_int64), ("pos", C.c_void_p), ("flag", C.c_void_p), ("src", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar16", C.c_void_p),
                ("n_wide", C.c_int64), ("wide_idx", C.c_void_p), ("wide_word", C.c_void_p), ("seq_off", C.c_void_p), ("seq2", C.c_void_p),
                ("n_exc", C.c_int64), ("exc_byte", C.c_void_p), ("exc_val", C.c_void_p)]


SYMBOLS = ["nptb_open", "nptb_close", "nptb_n_refs", "nptb_ref_name", "nptb_ref_len", "nptb_next_batch", "nptb_last_error", "nptb_pack", "nptb_inflate_raw"]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_bam())
        L.nptb_open.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        L.nptb_open.restype = C.c_void_p
        L.nptb_close.argtypes = [C.c_void_p]
        L.nptb_close.restype = None
        L.nptb_n_refs.argtypes = [C.c_void_p]
        L.nptb_n_refs.restype = C.c_int32
        L.nptb_ref_name.argtypes = [C.c_void_p, C.c_int32]
        L.nptb_ref_name.restype = C.c_char_p
        L.nptb_ref_len.argtypes = [C.c_void_p, C.c_int32]
        L.nptb_ref_len.restype = C.c_int64
        L.nptb_next_batch.argtypes = [C.c_void_p, C.POINTER(Filter), C.c_int32, C.c_int64, C.POINTER(BatchStruct)]
        L.nptb_next_batch.restype = C.c_int
        L.nptb_last_error.argtypes = [C.c_void_p]
        L.nptb_last_error.restype = C.c_char_p
        L.nptb_pack.argtypes = [C.c_void_p, C.POINTER(BatchStruct), C.POINTER(PackedStruct)]
        L.nptb_pack.restype = C.c_int
        L.nptb_inflate_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.nptb_inflate_raw.restype = C.c_long
        _lib = L
    return _lib


class BamError(RuntimeError):
    pass


class Reader:
    """for ref_id, batch, extra in Reader(path).batches(...): engine.submit(batch)"""

    def __init__(self, path, nthreads=0, pinned_from=None):
        """pinned_from: a loaded libnpt (engine.lib()) whose npt_alloc_pinned / npt_free_pinned back the batch buffers."""
        L = lib()
        err = C.create_string_buffer(512)
        a = f = None
        if pinned_from is not None:
            a = C.cast(pinned_from.npt_alloc_pinned, C.c_void_p)
            f = C.cast(pinned_from.npt_free_pinned, C.c_void_p)
        self.h = L.nptb_open(os.fsencode(path), nthreads, a, f, err, len(err))
        if not self.h:
            raise BamError(err.value.decode())
        self.refs = [(L.nptb_ref_name(self.h, i).decode(), L.nptb_ref_len(self.h, i)) for i in range(L.nptb_n_refs(self.h))]
        self.totals = {}

    def close(self):
        if self.h:
            lib().nptb_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def batches(self, src=0, max_batch_reads=1 << 20, fail_thresh=0.0, min_mapq=0, max_reads=0, ref_include=None, copy=True, decode_names=False,
                packed=False):
        """Yields (ref_id, synth.Batch, dict(q1, mapq, name_off, names_bytes[, names])).  copy=False hands out views valid until the
        next batch.  packed=True adds extra["packed"] = the batch in the compact transfer format (packed.PackedBatch, packed by
        nptb_pack on the reader's threads; copies when copy=True, else views valid until the next batch)."""
        L = lib()
        flt = Filter(fail_thresh, min_mapq, max_reads, None)
        keep = None
        if ref_include is not None:
            keep = np.ascontiguousarray(ref_include, dtype=np.uint8)
            flt.ref_include = keep.ctypes.data_as(C.POINTER(C.c_uint8))
        b = BatchStruct()
        while True:
            rc = L.nptb_next_batch(self.h, C.byref(flt), src, max_batch_reads, C.byref(b))
            self.totals = {k: getattr(b, k) for k in ("n_records", "n_unmapped", "n_excluded_ref", "n_low_quality", "n_short", "n_low_mapq",
                                                      "n_secondary", "stopped_at_max_reads")}
            if rc < 0:
                raise BamError(L.nptb_last_error(self.h).decode())
            if rc == 0:
                return
            n = b.n
            view = lambda p, m, dt: (np.ctypeslib.as_array(p, shape=(max(m, 1),))[:m].view(dt) if m >= 0 else None)
            co = view(b.cigar_off, n + 1, np.uint32)
            so = view(b.seq_off, n + 1, np.uint64)
            no = view(b.name_off, n + 1, np.uint32)
            arrs = [view(b.pos, n, np.int32), view(b.flag, n, np.uint16), view(b.src, n, np.uint16), co, view(b.cigar, int(co[-1]), np.uint32), so,
                    view(b.seq4, int(so[-1]), np.uint8)]
            q1, mq = view(b.q1, n, np.float64), view(b.mapq, n, np.uint8)
            extra = dict(name_off=no.copy(), names_bytes=C.string_at(b.names, int(no[-1])))
            if decode_names:
                nm = extra["names_bytes"]
                extra["names"] = [nm[int(no[i]):int(no[i + 1])].decode() for i in range(n)]
            if copy:
                arrs = [a.copy() for a in arrs]
                q1, mq = q1.copy(), mq.copy()
            extra.update(q1=q1, mapq=mq)
            if packed:
                from .packed import PackedBatch
                ps = PackedStruct()
                if L.nptb_pack(self.h, C.byref(b), C.byref(ps)) != 0:
                    raise BamError(L.nptb_last_error(self.h).decode())
                ncig, nseq = int(co[-1]), int(so[-1])

                def pv(ptr, m, dt):
                    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(1, m * np.dtype(dt).itemsize),))[:m * np.dtype(dt).itemsize].view(dt)
                    return a.copy() if copy else a
                extra["packed"] = PackedBatch(arrs[0], arrs[1], arrs[2], arrs[3], pv(ps.cigar16, ncig, np.uint16), pv(ps.wide_idx, ps.n_wide, np.uint32),
                                              pv(ps.wide_word, ps.n_wide, np.uint32), arrs[5], pv(ps.seq2, (nseq + 1) // 2, np.uint8),
                                              pv(ps.exc_byte, ps.n_exc, np.uint64), pv(ps.exc_val, ps.n_exc, np.uint8))
            yield b.ref_id, Batch(*arrs), extra


# ---------------------------------------------------------------------------------------------------- writer (tests)
def _bgzf_block(data, level):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    assert bsize <= 65535
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bam(path, refs, records, block_bytes=0xFF00, level=1, header_text="@HD\tVN:1.6\tSO:unsorted\n"):
    """refs: [(name, length)].  records: iterable of dict(ref_id, pos (1-based), flag, mapq, name, cigar (uint32 words),
    seq4 (packed bytes), l_seq, qual (bytes of length l_seq, or None for absent qualities)).  BGZF with blocks of
    block_bytes uncompressed bytes (records straddle block boundaries, as in real files)."""
    out = bytearray()
    text = header_text.encode()
    out += b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for name, ln in refs:
        nb = name.encode() + b"\x00"
        out += struct.pack("<I", len(nb)) + nb + struct.pack("<I", ln)
    for r in records:
        name = r["name"].encode() + b"\x00"
        cig = np.asarray(r["cigar"], dtype="<u4").tobytes()
        l_seq = int(r["l_seq"])
        seq = bytes(r["seq4"])
        assert len(seq) == (l_seq + 1) // 2
        qual = bytes(r["qual"]) if r.get("qual") is not None else b"\xff" * l_seq
        assert len(qual) == l_seq
        body = struct.pack("<iiBBHHHIiii", r["ref_id"], r["pos"] - 1, len(name), r.get("mapq", 60), 4680, len(cig) // 4, r["flag"], l_seq, -1, -1, 0)
        body += name + cig + seq + qual + r.get("tags", b"")
        out += struct.pack("<I", len(body)) + body
    with open(path, "wb") as f:
        for o in range(0, len(out), block_bytes):
            f.write(_bgzf_block(bytes(out[o:o + block_bytes]), level))
        f.write(_bgzf_block(b"", level))  # EOF marker


def records_from_batch(batch, ref_id=0, quals=None, mapq=None, names=None):
    """synth.Batch -> write_bam records (bases need per-read base counts: the generator pads odd lengths with code 0,
    so l_seq is taken from the CIGAR: M/I/S/=/X lengths)."""
    recs = []
    for i in range(batch.n):
        cg = batch.cigar[int(batch.cigar_off[i]):int(batch.cigar_off[i + 1])]
        op, ln = cg & 15, cg >> 4
        l_seq = int(ln[(op == 0) | (op == 1) | (op == 4) | (op == 7) | (op == 8)].sum())
        sq = batch.seq4[int(batch.seq_off[i]):int(batch.seq_off[i + 1])]
        sq = sq[:(l_seq + 1) // 2]
        recs.append(dict(ref_id=ref_id, pos=int(batch.pos[i]), flag=int(batch.flag[i]), mapq=60 if mapq is None else int(mapq[i]),
                         name=f"read{i}" if names is None else names[i], cigar=cg, seq4=sq.tobytes(), l_seq=l_seq,
                         qual=None if quals is None else quals[i]))
    return recs
