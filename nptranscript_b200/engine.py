"""Thin ctypes wrapper over libnpt.so (include/npt.h).  Product path: fails loudly when the CUDA library is missing,
cannot be loaded, or finds no sm_100 device — there is no CPU fallback and nothing here touches oracle/."""
import ctypes as C
import os

import numpy as np

from . import _build, abi

_lib = None


class NptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"npt error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("NPT_LIBNPT") or _build.build_libnpt()  # NPT_LIBNPT: an experimental build (profiles/NOTES.md)
        _lib = abi.declare(C.CDLL(path))
        if _lib.npt_abi_version() != abi.NPT_ABI_VERSION:
            raise RuntimeError("libnpt.so ABI version mismatch")
    return _lib


def make_config(**kw):
    """npt_config with the reference's defaults (ViralTranscriptAnalysisCmd2.java:118-190, 336-372)."""
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


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).copy()


class Results(dict):
    """npt_results copied into numpy arrays (same field names)."""

    @classmethod
    def from_struct(cls, r, S, what):
        o = cls()
        n = r.n_reads
        o["n_reads"], o["n_clusters"], o["n_subs"], o["n_orf"] = n, r.n_clusters, r.n_subs, r.n_orf
        o["n_slow_reads"] = r.n_slow_reads
        if what & abi.NPT_FETCH_READS:
            for k in ("cluster_id", "sub_id", "start_pos", "end_pos", "read_st", "read_en", "maps_sum", "err_sum"):
                o[k] = _np(getattr(r, k), n, np.int32)
            o["read_flags"] = _np(r.read_flags, n, np.uint32)
            o["break_off"] = _np(r.break_off, n + 1, np.uint64) if n else np.zeros(1, np.uint64)
            o["breaks"] = _np(r.breaks, int(o["break_off"][-1]), np.int32)
        if what & abi.NPT_FETCH_TABLES:
            nc, ns = r.n_clusters, r.n_subs
            o["cl_first_read"] = _np(r.cl_first_read, nc, np.int64)
            o["cl_key_off"] = _np(r.cl_key_off, nc + 1, np.uint32)
            o["cl_key_tok"] = _np(r.cl_key_tok, int(o["cl_key_off"][-1]) if nc else 0, np.int32)
            o["cl_read_count"] = _np(r.cl_read_count, nc * S, np.int32).reshape(nc, S)
            o["cl_sub_off"] = _np(r.cl_sub_off, nc + 1, np.uint32)
            o["sub_first_read"] = _np(r.sub_first_read, ns, np.int64)
            o["sub_key_off"] = _np(r.sub_key_off, ns + 1, np.uint32)
            o["sub_key"] = _np(r.sub_key, int(o["sub_key_off"][-1]) if ns else 0, np.int32)
            o["sub_count"] = _np(r.sub_count, ns * S, np.int32).reshape(ns, S)
            o["sub_break_sum"] = _np(r.sub_break_sum, ns, np.int32)
            o["sub_sum_off"] = _np(r.sub_sum_off, ns + 1, np.uint32)
            o["sub_sums"] = _np(r.sub_sums, int(o["sub_sum_off"][-1]) if ns else 0, np.int64)
            o["sub_flags"] = _np(r.sub_flags, ns, np.uint8)
            o["total_counts"] = _np(r.total_counts, S, np.int32)
            o["spliced"] = _np(r.spliced, S * r.n_orf, np.int32).reshape(S, r.n_orf)
            o["unspliced"] = _np(r.unspliced, S * r.n_orf, np.int32).reshape(S, r.n_orf)
            o["n_pairs"] = r.n_pairs
            for k in ("pair_src", "pair_level", "pair_start", "pair_end", "pair_count"):
                o[k] = _np(getattr(r, k), r.n_pairs, np.int32)
        if (what & abi.NPT_FETCH_COVERAGE) and r.depth:
            nc, st = r.n_clusters, r.cov_stride
            o["cov_stride"] = st
            for k in ("depth", "errors", "map_start", "map_end"):
                o[k] = _np(getattr(r, k), nc * S * st, np.int32).reshape(nc, S, st)
        return o


class Engine:
    """One npt_ctx.  begin_chrom → submit* → end_chrom → fetch."""

    def __init__(self, cfg):
        self.L = lib()
        self.cfg = cfg
        self.S = cfg.num_sources
        h = C.c_void_p()
        rc = self.L.npt_create(C.byref(cfg), C.byref(h))
        if rc:
            raise NptError(rc, self.L.npt_last_error(None).decode())
        self.h = h
        self._keep = []

    def _ck(self, rc):
        if rc:
            raise NptError(rc, self.L.npt_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.npt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin_chrom(self, chrom_index, ref_codes, annotation):
        ref = np.ascontiguousarray(ref_codes, dtype=np.uint8)
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
        self._ck(self.L.npt_begin_chrom(self.h, chrom_index, ref.ctypes.data, len(ref), C.byref(a)))
        self._keep = []

    def set_read_index_base(self, base):
        self._ck(self.L.npt_set_read_index_base(self.h, base))

    def set_next_read_index(self, index):
        self._ck(self.L.npt_set_next_read_index(self.h, index))

    def submit(self, batch):
        """batch: synth.Batch-like with host numpy arrays (kept alive until wait/end_chrom)."""
        s = abi.npt_batch()
        s.n, s.location = batch.n, abi.NPT_MEM_HOST
        s.pos, s.flag, s.src = batch.pos.ctypes.data, batch.flag.ctypes.data, batch.src.ctypes.data
        s.cigar_off, s.cigar = batch.cigar_off.ctypes.data, batch.cigar.ctypes.data
        s.seq_off, s.seq4 = batch.seq_off.ctypes.data, batch.seq4.ctypes.data
        self._keep.append(batch)
        self._ck(self.L.npt_submit(self.h, C.byref(s)))

    def submit_packed(self, pb):
        """pb: packed.PackedBatch (compact transfer format, host arrays kept alive until wait/end_chrom)."""
        s = abi.npt_packed_batch()
        s.n = pb.n
        s.pos, s.flag, s.src = pb.pos.ctypes.data, pb.flag.ctypes.data, pb.src.ctypes.data
        s.cigar_off, s.cigar16 = pb.cigar_off.ctypes.data, pb.cigar16.ctypes.data
        s.n_wide, s.wide_idx, s.wide_word = len(pb.wide_idx), pb.wide_idx.ctypes.data, pb.wide_word.ctypes.data
        s.seq_off, s.seq2 = pb.seq_off.ctypes.data, pb.seq2.ctypes.data
        s.n_exc, s.exc_byte, s.exc_val = len(pb.exc_byte), pb.exc_byte.ctypes.data, pb.exc_val.ctypes.data
        self._keep.append(pb)
        self._ck(self.L.npt_submit_packed(self.h, C.byref(s)))

    def submit_device(self, n, pos, flag, src, cigar_off, cigar, seq_off, seq4):
        """Device pointers (ints) of a batch already resident in HBM."""
        s = abi.npt_batch()
        s.n, s.location = n, abi.NPT_MEM_DEVICE
        s.pos, s.flag, s.src, s.cigar_off, s.cigar, s.seq_off, s.seq4 = pos, flag, src, cigar_off, cigar, seq_off, seq4
        self._ck(self.L.npt_submit(self.h, C.byref(s)))

    def wait(self):
        self._ck(self.L.npt_wait(self.h))
        self._keep = []

    def end_chrom(self):
        self._ck(self.L.npt_end_chrom(self.h))
        self._keep = []

    def fetch(self, what=abi.NPT_FETCH_ALL):
        r = abi.npt_results()
        self._ck(self.L.npt_fetch_results(self.h, what, C.byref(r)))
        return Results.from_struct(r, self.S, what)

    def fetch_raw(self, what=abi.NPT_FETCH_ALL):
        """No numpy copies: the npt_results struct with pointers into the context's (pinned) buffers."""
        r = abi.npt_results()
        self._ck(self.L.npt_fetch_results(self.h, what, C.byref(r)))
        return r

    def release(self):
        self._ck(self.L.npt_release_results(self.h))

    def kernel_ms(self):
        w, f, wl, tl = C.c_float(), C.c_float(), C.c_int64(), C.c_int64()
        self._ck(self.L.npt_last_kernel_ms(self.h, C.byref(w), C.byref(f), C.byref(wl), C.byref(tl)))
        st = C.c_float()
        self._ck(self.L.npt_last_step_ms(self.h, C.byref(st)))
        stg = (C.c_float * 6)()
        self._ck(self.L.npt_last_stage_ms(self.h, stg))
        return dict(walk_ms=w.value, finalize_ms=f.value, walk_launches=wl.value, total_launches=tl.value, step_ms=st.value,
                    fastwalk_ms=stg[0], walk0_ms=stg[1], walk2_ms=stg[2], cluster_ms=stg[3], walk1_ms=stg[4], fastcov_ms=stg[5])

    def device_view(self):
        v = abi.npt_device_view()
        self._ck(self.L.npt_device_view_get(self.h, C.byref(v)))
        return v

    def export_keys(self):
        """(device pointer, int32 words) of this rank's packed key buffer; valid until the next export or npt_destroy."""
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.L.npt_export_keys(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def import_keys(self, dev_ptr, n_words):
        self._ck(self.L.npt_import_keys(self.h, dev_ptr, n_words))

    # ---- collectives inside the library (csrc/comm.inl)
    def comm_init(self, id_bytes, rank, world):
        buf = C.create_string_buffer(bytes(id_bytes), abi.NPT_COMM_ID_BYTES)
        self._ck(self.L.npt_comm_init(self.h, buf, rank, world))

    def end_chrom_all(self):
        """Collective end of chromosome: key exchange + end_chrom + sum of the counters, all inside libnpt over NCCL."""
        self._ck(self.L.npt_end_chrom_all(self.h))
        self._keep = []

    def exchange_ms(self):
        k, r = C.c_float(), C.c_float()
        self._ck(self.L.npt_last_exchange_ms(self.h, C.byref(k), C.byref(r)))
        return k.value, r.value


def comm_unique_id():
    buf = C.create_string_buffer(abi.NPT_COMM_ID_BYTES)
    rc = lib().npt_comm_unique_id(buf)
    if rc:
        raise NptError(rc, lib().npt_last_error(None).decode())
    return buf.raw


def _annotation_struct(annotation):
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


class MultiEngine:
    """One npt_mg: the in-process multi-GPU context (one npt_ctx and one host thread per device inside libnpt, NCCL between
    them).  Same call sequence as Engine; host batches are split over the devices by the library."""

    def __init__(self, cfg, devices):
        self.L = lib()
        self.S = cfg.num_sources
        dv = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.L.npt_mg_create(C.byref(cfg), dv, len(devices), C.byref(h))
        if rc:
            raise NptError(rc, self.L.npt_mg_last_error(None).decode())
        self.h = h
        self._keep = []

    def _ck(self, rc):
        if rc:
            raise NptError(rc, self.L.npt_mg_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.npt_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_read_index_base(self, base):
        self._ck(self.L.npt_mg_set_read_index_base(self.h, base))

    def begin_chrom(self, chrom_index, ref_codes, annotation):
        ref = np.ascontiguousarray(ref_codes, dtype=np.uint8)
        a, keep = _annotation_struct(annotation)
        self._ck(self.L.npt_mg_begin_chrom(self.h, chrom_index, ref.ctypes.data, len(ref), C.byref(a)))
        self._keep = []

    def submit(self, batch):
        s = abi.npt_batch()
        s.n, s.location = batch.n, abi.NPT_MEM_HOST
        s.pos, s.flag, s.src = batch.pos.ctypes.data, batch.flag.ctypes.data, batch.src.ctypes.data
        s.cigar_off, s.cigar = batch.cigar_off.ctypes.data, batch.cigar.ctypes.data
        s.seq_off, s.seq4 = batch.seq_off.ctypes.data, batch.seq4.ctypes.data
        self._keep.append(batch)
        self._ck(self.L.npt_mg_submit(self.h, C.byref(s)))

    def submit_packed(self, pb):
        s = abi.npt_packed_batch()
        s.n = pb.n
        s.pos, s.flag, s.src = pb.pos.ctypes.data, pb.flag.ctypes.data, pb.src.ctypes.data
        s.cigar_off, s.cigar16 = pb.cigar_off.ctypes.data, pb.cigar16.ctypes.data
        s.n_wide, s.wide_idx, s.wide_word = len(pb.wide_idx), pb.wide_idx.ctypes.data, pb.wide_word.ctypes.data
        s.seq_off, s.seq2 = pb.seq_off.ctypes.data, pb.seq2.ctypes.data
        s.n_exc, s.exc_byte, s.exc_val = len(pb.exc_byte), pb.exc_byte.ctypes.data, pb.exc_val.ctypes.data
        self._keep.append(pb)
        self._ck(self.L.npt_mg_submit_packed(self.h, C.byref(s)))

    def end_chrom(self):
        self._ck(self.L.npt_mg_end_chrom(self.h))
        self._keep = []

    def fetch(self, what=abi.NPT_FETCH_ALL):
        r = abi.npt_results()
        self._ck(self.L.npt_mg_fetch_results(self.h, what, C.byref(r)))
        return Results.from_struct(r, self.S, what)

    def fetch_raw(self, what=abi.NPT_FETCH_ALL):
        r = abi.npt_results()
        self._ck(self.L.npt_mg_fetch_results(self.h, what, C.byref(r)))
        return r

    def release(self):
        self._ck(self.L.npt_mg_release_results(self.h))

    def step_ms(self):
        s, k, r = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.npt_mg_last_step_ms(self.h, C.byref(s), C.byref(k), C.byref(r)))
        return dict(step_ms=s.value, key_exchange_ms=k.value, reduce_ms=r.value)


def run_mg(cfg, devices, ref_codes, annotation, batches, chrom_index=0, what=abi.NPT_FETCH_ALL, packed=False):
    e = MultiEngine(cfg, devices)
    try:
        e.begin_chrom(chrom_index, ref_codes, annotation)
        for b in batches:
            if packed:
                from .packed import pack
                e.submit_packed(pack(b))
            else:
                e.submit(b)
        e.end_chrom()
        r = e.fetch(what)
        r["_timing"] = e.step_ms()
        return r
    finally:
        e.close()


def run(cfg, ref_codes, annotation, batches, chrom_index=0, what=abi.NPT_FETCH_ALL, packed=False):
    """packed=True sends every batch in the compact transfer format (npt_submit_packed)."""
    e = Engine(cfg)
    try:
        e.begin_chrom(chrom_index, ref_codes, annotation)
        for b in batches:
            if packed:
                from .packed import pack
                e.submit_packed(pack(b))
            else:
                e.submit(b)
        e.end_chrom()
        r = e.fetch(what)
        r["_timing"] = e.kernel_ms()
        return r
    finally:
        e.close()
