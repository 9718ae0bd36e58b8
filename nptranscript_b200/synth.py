"""ctypes front-end of synth/synth.cpp: deterministic synthetic reads as npt_batch structure-of-arrays."""
import ctypes as C
import os

import numpy as np

from . import _build
from .workloads import Workload

_MAXA = 32


class _Params(C.Structure):
    _fields_ = [
        ("genome_len", C.c_int32), ("polya_len", C.c_int32), ("donor", C.c_int32), ("n_acceptors", C.c_int32),
        ("acceptor", C.c_int32 * _MAXA), ("acceptor_w", C.c_double * _MAXA),
        ("w_canonical", C.c_double), ("w_leaderless", C.c_double), ("w_genomic", C.c_double), ("w_double", C.c_double),
        ("p_mismatch", C.c_double), ("p_ins", C.c_double), ("p_del", C.c_double), ("indel_mean", C.c_double),
        ("p_neg", C.c_double), ("num_sources", C.c_int32), ("src_fixed", C.c_int32), ("use_eqx", C.c_int32),
        ("leaderless_len", C.c_double), ("genomic_len", C.c_double),
    ]


class _HostParams(C.Structure):
    _fields_ = [
        ("chrom_len", C.c_int32), ("n_genes", C.c_int32), ("gene_exon_off", C.c_void_p), ("exon_start", C.c_void_p), ("exon_end", C.c_void_p),
        ("p_skip", C.c_double), ("p_trunc5", C.c_double), ("p_intergenic", C.c_double), ("gene_skew", C.c_double),
        ("p_mismatch", C.c_double), ("p_ins", C.c_double), ("p_del", C.c_double), ("indel_mean", C.c_double), ("p_neg", C.c_double),
        ("num_sources", C.c_int32), ("src_fixed", C.c_int32), ("transcriptome", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_synth())
        _lib.npts_generate.restype = C.c_void_p
        _lib.npts_generate.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int]
        _lib.npts_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3
        _lib.npts_copy.argtypes = [C.c_void_p] * 8
        _lib.npts_free.argtypes = [C.c_void_p]
        _lib.npts_genome.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]
        _lib.npts_generate_host.restype = C.c_void_p
        _lib.npts_generate_host.argtypes = [C.POINTER(_HostParams), C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int]
    return _lib


def genome_codes(w: Workload) -> np.ndarray:
    if w.genome_file:   # the real reference (BAM 4-bit codes of its letters)
        from .workloads import read_fasta, ref_char_to_code
        return ref_char_to_code(read_fasta(w.genome_file)[2])
    out = np.empty(w.genome_len, dtype=np.uint8)
    lib().npts_genome(w.genome_seed, w.genome_len, w.polya_len, out.ctypes.data)
    return out


class Batch:
    """Host arrays of one npt_batch.  Arrays may be numpy-owned or views over pinned memory."""

    def __init__(self, pos, flag, src, cigar_off, cigar, seq_off, seq4):
        self.pos, self.flag, self.src = pos, flag, src
        self.cigar_off, self.cigar, self.seq_off, self.seq4 = cigar_off, cigar, seq_off, seq4

    @property
    def n(self):
        return len(self.pos)

    def nbytes(self):
        return sum(a.nbytes for a in (self.pos, self.flag, self.src, self.cigar_off, self.cigar, self.seq_off, self.seq4))

    def slice(self, a, b):
        co, so = self.cigar_off, self.seq_off
        return Batch(self.pos[a:b], self.flag[a:b], self.src[a:b], (co[a:b + 1] - co[a]).astype(np.uint32),
                     self.cigar[co[a]:co[b]], (so[a:b + 1] - so[a]).astype(np.uint64), self.seq4[so[a]:so[b]])


def generate(w: Workload, seed: int, n: int, first_index: int = 0, num_sources: int = 1, src_fixed: int = -1,
             use_eqx: bool = False, nthreads: int = 0, alloc=None) -> Batch:
    """alloc(nbytes) -> writable buffer object (e.g. pinned memory); default numpy."""
    p = _Params()
    p.genome_len, p.polya_len, p.donor, p.n_acceptors = w.genome_len, w.polya_len, w.donor, len(w.acceptors)
    for i, (a, wt) in enumerate(zip(w.acceptors, w.acceptor_w)):
        p.acceptor[i], p.acceptor_w[i] = a, wt
    p.w_canonical, p.w_leaderless, p.w_genomic, p.w_double = w.w_mix
    p.p_mismatch, p.p_ins, p.p_del, p.indel_mean, p.p_neg = w.p_mismatch, w.p_ins, w.p_del, w.indel_mean, w.p_neg
    p.num_sources, p.src_fixed, p.use_eqx = num_sources, src_fixed, int(use_eqx)
    p.leaderless_len, p.genomic_len = w.leaderless_len, w.genomic_len
    g = genome_codes(w)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    L = lib()
    h = L.npts_generate(C.byref(p), g.ctypes.data, seed, first_index, n, nthreads)
    try:
        nn, nc, ns = C.c_int64(), C.c_int64(), C.c_int64()
        L.npts_sizes(h, C.byref(nn), C.byref(nc), C.byref(ns))

        def mk(count, dt):
            if alloc is None:
                return np.empty(count, dtype=dt)
            return np.frombuffer(alloc(max(1, count * np.dtype(dt).itemsize)), dtype=dt, count=count)

        b = Batch(mk(n, np.int32), mk(n, np.uint16), mk(n, np.uint16), mk(n + 1, np.uint32), mk(nc.value, np.uint32),
                  mk(n + 1, np.uint64), mk(ns.value, np.uint8))
        L.npts_copy(h, b.pos.ctypes.data, b.flag.ctypes.data, b.src.ctypes.data, b.cigar_off.ctypes.data,
                    b.cigar.ctypes.data if nc.value else None, b.seq_off.ctypes.data, b.seq4.ctypes.data if ns.value else None)
    finally:
        L.npts_free(h)
    return b


def _collect(L, h, n, alloc):
    try:
        nn, nc, ns = C.c_int64(), C.c_int64(), C.c_int64()
        L.npts_sizes(h, C.byref(nn), C.byref(nc), C.byref(ns))

        def mk(count, dt):
            if alloc is None:
                return np.empty(count, dtype=dt)
            return np.frombuffer(alloc(max(1, count * np.dtype(dt).itemsize)), dtype=dt, count=count)

        b = Batch(mk(n, np.int32), mk(n, np.uint16), mk(n, np.uint16), mk(n + 1, np.uint32), mk(nc.value, np.uint32),
                  mk(n + 1, np.uint64), mk(ns.value, np.uint8))
        L.npts_copy(h, b.pos.ctypes.data, b.flag.ctypes.data, b.src.ctypes.data, b.cigar_off.ctypes.data,
                    b.cigar.ctypes.data if nc.value else None, b.seq_off.ctypes.data, b.seq4.ctypes.data if ns.value else None)
    finally:
        L.npts_free(h)
    return b


def host_genome_codes(c) -> np.ndarray:
    """Seeded uniform ACGT for a workloads.HostChromosome (no polyA tail)."""
    out = np.empty(c.length, dtype=np.uint8)
    lib().npts_genome(c.seed * 1000003 + c.index, c.length, 0, out.ctypes.data)
    return out


def generate_host(c, genome, seed: int, n: int, first_index: int = 0, num_sources: int = 1, src_fixed: int = -1, transcriptome: bool = False,
                  nthreads: int = 0, alloc=None) -> Batch:
    """Spliced long reads on one chromosome of the synthetic gene model (workloads.HostChromosome); read i depends only on
    (seed, first_index + i).  transcriptome=True: source 0 carries the full transcripts aligned error-free."""
    p = _HostParams()
    off = np.ascontiguousarray(c.gene_exon_off, np.int32); es = np.ascontiguousarray(c.exon_start, np.int32); ee = np.ascontiguousarray(c.exon_end, np.int32)
    p.chrom_len, p.n_genes = c.length, c.n_genes
    p.gene_exon_off, p.exon_start, p.exon_end = off.ctypes.data, es.ctypes.data, ee.ctypes.data
    p.p_skip, p.p_trunc5, p.p_intergenic, p.gene_skew = c.p_skip, c.p_trunc5, c.p_intergenic, c.gene_skew
    p.p_mismatch, p.p_ins, p.p_del, p.indel_mean, p.p_neg = c.p_mismatch, c.p_ins, c.p_del, c.indel_mean, c.p_neg
    p.num_sources, p.src_fixed, p.transcriptome = num_sources, src_fixed, int(transcriptome)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    L = lib()
    g = np.ascontiguousarray(genome, np.uint8)
    h = L.npts_generate_host(C.byref(p), g.ctypes.data, seed, first_index, n, nthreads)
    return _collect(L, h, n, alloc)


def from_reads(reads) -> Batch:
    """Hand-built reads for known-answer tests: iterable of (pos, flag, src, cigar_string, bases_string)."""
    ops = "MIDNSHP=X"
    code = {ch: i for i, ch in enumerate("=ACMGRSVTWYHKDBN")}
    pos, flag, src, coff, cig, soff, seq = [], [], [], [0], [], [0], bytearray()
    import re
    for (p_, f_, s_, cs, bs) in reads:
        pos.append(p_); flag.append(f_); src.append(s_)
        for ln, op in re.findall(r"(\d+)([A-Za-z=?])", cs):
            o = ops.index(op) if op in ops else 9  # '?' etc. → invalid op code 9
            cig.append((int(ln) << 4) | o)
        coff.append(len(cig))
        nib = [code[c.upper()] for c in bs]
        if len(nib) % 2:
            nib.append(0)
        for k in range(0, len(nib), 2):
            seq.append((nib[k] << 4) | nib[k + 1])
        soff.append(len(seq))
    return Batch(np.array(pos, np.int32), np.array(flag, np.uint16), np.array(src, np.uint16), np.array(coff, np.uint32),
                 np.array(cig, np.uint32), np.array(soff, np.uint64), np.frombuffer(bytes(seq), dtype=np.uint8).copy())
