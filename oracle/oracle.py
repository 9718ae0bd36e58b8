"""ctypes front-end of oracle/liboracle.so (oracle.cpp): the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no runnable reference, no golden vectors in the reference; see
oracle.cpp header and DESIGN.md).  Returns the same field names as nptranscript_b200.results.Results so tests can
compare field by field."""
import ctypes as C
import os
import subprocess

import numpy as np

from nptranscript_b200 import abi  # type definitions of include/npt.h only

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "npt.h")
    stale = lambda: not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or stale():
        import fcntl
        with open(LIB + ".lock", "w") as f:  # one builder at a time (parallel test workers, ranks)
            fcntl.flock(f, fcntl.LOCK_EX)
            if force or stale():
                r = subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.npto_create.restype = C.c_void_p
        L.npto_create.argtypes = [C.POINTER(abi.npt_config), C.c_int, C.c_void_p, C.c_int64, C.POINTER(abi.npt_annotation), C.c_int]
        L.npto_destroy.argtypes = [C.c_void_p]
        L.npto_submit.argtypes = [C.c_void_p, C.POINTER(abi.npt_batch), C.c_int]
        for f in ("npto_n_reads", "npto_n_pairs", "npto_n_breaks", "npto_key_tokens", "npto_sub_key_len", "npto_sub_sum_len"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("npto_n_clusters", "npto_n_subs"):
            getattr(L, f).restype = C.c_int32
            getattr(L, f).argtypes = [C.c_void_p]
        L.npto_get_reads.argtypes = [C.c_void_p] * 12
        L.npto_get_tables.argtypes = [C.c_void_p] * 18
        L.npto_get_pairs.argtypes = [C.c_void_p] * 6
        L.npto_get_coverage.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4
        L.npto_sub_consensus.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        _lib = L
    return _lib


def make_config(**kw):
    """npt_config with the reference's defaults for the path (ViralTranscriptAnalysisCmd2.java:118-190, 336-372)."""
    c = abi.npt_config()
    c.abi_version = abi.NPT_ABI_VERSION
    d = dict(bin=1, break_thresh=1000, include_start=1, coronavirus=1, start_thresh=100, end_thresh=100, enforce_strand=0,
             record_depth=0, calc_breaks=1, first_is_transcriptome=0, track_break_sums=1, num_sources=1, device=0,
             max_clusters=0, max_subclusters=0, max_break_pairs=0, max_coverage_clusters=0)
    d.update({k: v for k, v in kw.items() if k != "rna"})
    for k, v in d.items():
        setattr(c, k, int(v))
    rna = kw.get("rna", [1] * abi.NPT_MAX_SOURCES)
    for i in range(abi.NPT_MAX_SOURCES):
        c.rna[i] = int(rna[i]) if i < len(rna) else 0
    return c


def _annot_struct(annotation):
    st = np.asarray(annotation.start, dtype=np.int32)
    en = np.asarray(annotation.end, dtype=np.int32)
    sd = np.asarray(annotation.strand, dtype=np.uint8)
    ni = np.asarray(annotation.name_id, dtype=np.int32)
    nh = np.asarray(annotation.name_hash if annotation.kind == abi.NPT_ANNOT_GFF else [], dtype=np.uint32)
    a = abi.npt_annotation()
    a.kind, a.n = annotation.kind, len(st)
    a.start = st.ctypes.data_as(C.POINTER(C.c_int32)); a.end = en.ctypes.data_as(C.POINTER(C.c_int32))
    a.strand = sd.ctypes.data_as(C.POINTER(C.c_uint8)); a.name_id = ni.ctypes.data_as(C.POINTER(C.c_int32))
    a.name_hash = nh.ctypes.data_as(C.POINTER(C.c_uint32)) if len(nh) else None
    return a, (st, en, sd, ni, nh)


def _batch_struct(b):
    s = abi.npt_batch()
    s.n, s.location = b.n, abi.NPT_MEM_HOST
    s.pos, s.flag, s.src = b.pos.ctypes.data, b.flag.ctypes.data, b.src.ctypes.data
    s.cigar_off, s.cigar, s.seq_off, s.seq4 = b.cigar_off.ctypes.data, b.cigar.ctypes.data, b.seq_off.ctypes.data, b.seq4.ctypes.data
    return s


class OracleRun:
    """One chromosome's worth of sequential processing (an IdentityProfileHolder + CigarClusters + BreakPoints)."""

    def __init__(self, cfg, ref_codes, annotation, chrom_index=0, keep_reads=True):
        self.cfg = cfg
        self.S = cfg.num_sources
        self.ref = np.ascontiguousarray(ref_codes, dtype=np.uint8)
        self.n_orf = 0 if annotation.kind == abi.NPT_ANNOT_GFF else len(annotation.start)
        a, self._keep = _annot_struct(annotation)
        self.h = lib().npto_create(C.byref(cfg), chrom_index, self.ref.ctypes.data, len(self.ref), C.byref(a), int(keep_reads))
        self.keep_reads = keep_reads

    def submit(self, batch, nthreads=1):
        s = _batch_struct(batch)
        lib().npto_submit(self.h, C.byref(s), nthreads)

    def close(self):
        if self.h:
            lib().npto_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def results(self, coverage=True):
        L, h, S = lib(), self.h, self.S
        r = {}
        n = L.npto_n_reads(h)
        r["n_reads"] = n
        if self.keep_reads:
            nb = L.npto_n_breaks(h)
            names = ["cluster_id", "sub_id", "start_pos", "end_pos", "read_st", "read_en", "read_flags", "maps_sum", "err_sum"]
            arrs = [np.empty(n, dtype=np.uint32 if k == "read_flags" else np.int32) for k in names]
            boff = np.empty(n + 1, dtype=np.uint64)
            br = np.empty(nb, dtype=np.int32)
            L.npto_get_reads(h, *[a.ctypes.data for a in arrs], boff.ctypes.data, br.ctypes.data)
            r.update(dict(zip(names, arrs)))
            r["break_off"], r["breaks"] = boff, br
        nc, ns = L.npto_n_clusters(h), L.npto_n_subs(h)
        nk, nsk, nss = L.npto_key_tokens(h), L.npto_sub_key_len(h), L.npto_sub_sum_len(h)
        r["n_clusters"], r["n_subs"] = nc, ns
        t = dict(cl_first_read=np.empty(nc, np.int64), cl_key_off=np.empty(nc + 1, np.uint32), cl_key_tok=np.empty(nk, np.int32),
                 cl_read_count=np.empty((nc, S), np.int32), cl_sub_off=np.empty(nc + 1, np.uint32),
                 sub_first_read=np.empty(ns, np.int64), sub_key_off=np.empty(ns + 1, np.uint32), sub_key=np.empty(nsk, np.int32),
                 sub_count=np.empty((ns, S), np.int32), sub_break_sum=np.empty(ns, np.int32), sub_sum_off=np.empty(ns + 1, np.uint32),
                 sub_sums=np.empty(nss, np.int64), sub_true_breaks=np.empty(nss, np.float64), sub_flags=np.empty(ns, np.uint8),
                 total_counts=np.empty(S, np.int32), spliced=np.empty((S, self.n_orf), np.int32),
                 unspliced=np.empty((S, self.n_orf), np.int32))
        L.npto_get_tables(h, *[a.ctypes.data for a in t.values()])
        r.update(t)
        npairs = L.npto_n_pairs(h)
        p = [np.empty(npairs, np.int32) for _ in range(5)]
        L.npto_get_pairs(h, *[a.ctypes.data for a in p])
        r["n_pairs"] = npairs
        r.update(dict(zip(["pair_src", "pair_level", "pair_start", "pair_end", "pair_count"], p)))
        if coverage and self.cfg.record_depth:
            stride = len(self.ref) + 2
            cov = [np.zeros((nc, S, stride), np.int32) for _ in range(4)]
            for c in range(nc):
                L.npto_get_coverage(h, c, *[a[c].ctypes.data for a in cov])
            r["cov_stride"] = stride
            r.update(dict(zip(["depth", "errors", "map_start", "map_end"], cov)))
        return r

    def sub_consensus(self, sub_index):
        out = np.empty(256, np.int32)
        k = lib().npto_sub_consensus(self.h, sub_index, out.ctypes.data, 256)
        return None if k < 0 else out[:k].copy()


def run(cfg, ref_codes, annotation, batches, chrom_index=0, nthreads=1, coverage=True):
    o = OracleRun(cfg, ref_codes, annotation, chrom_index)
    for b in batches:
        o.submit(b, nthreads)
    r = o.results(coverage)
    r["_run"] = o
    return r
