"""ctypes mirror of include/npt.h (the C ABI).  Field order and types must match the header exactly;
tests/test_abi.py checks sizes/offsets against a C probe and that the built library exports every symbol."""
import ctypes as C

NPT_ABI_VERSION = 1
NPT_MAX_SOURCES = 16

NPT_OK, NPT_EINVAL, NPT_ENODEVICE, NPT_ECUDA, NPT_ENOMEM, NPT_ECAPACITY, NPT_ESTATE = 0, -1, -2, -3, -4, -5, -6
NPT_ANNOT_CSV, NPT_ANNOT_EMPTY, NPT_ANNOT_GFF = 0, 1, 2
NPT_MEM_HOST, NPT_MEM_DEVICE = 0, 1
NPT_TOK_ST, NPT_TOK_END, NPT_TOK_PLUS, NPT_TOK_MINUS, NPT_TOK_NOENDS, NPT_TOK_GENESET = -1, -2, -3, -4, -5, -6
NPT_RF_TYPE_MASK, NPT_RF_LEADER, NPT_RF_NEG, NPT_RF_BADCIGAR, NPT_RF_SLOWPATH = 0x3, 0x4, 0x8, 0x10, 0x20
NPT_FETCH_READS, NPT_FETCH_TABLES, NPT_FETCH_COVERAGE, NPT_FETCH_ALL = 1, 2, 4, 7

i32, i64, u8, u16, u32, u64 = C.c_int32, C.c_int64, C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64
P = C.POINTER


class npt_config(C.Structure):
    _fields_ = [
        ("abi_version", i32), ("bin", i32), ("break_thresh", i32), ("include_start", i32), ("coronavirus", i32),
        ("start_thresh", i32), ("end_thresh", i32), ("enforce_strand", i32), ("record_depth", i32), ("calc_breaks", i32),
        ("first_is_transcriptome", i32), ("track_break_sums", i32), ("num_sources", i32), ("rna", i32 * NPT_MAX_SOURCES),
        ("device", i32), ("max_clusters", i32), ("max_subclusters", i32), ("max_break_pairs", i32),
        ("max_coverage_clusters", i32),
    ]


class npt_annotation(C.Structure):
    _fields_ = [("kind", i32), ("n", i32), ("start", P(i32)), ("end", P(i32)), ("strand", P(u8)), ("name_id", P(i32)),
                ("name_hash", P(u32))]


class npt_batch(C.Structure):
    _fields_ = [
        ("n", i64), ("location", i32), ("reserved", i32), ("pos", C.c_void_p), ("flag", C.c_void_p), ("src", C.c_void_p),
        ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("seq_off", C.c_void_p), ("seq4", C.c_void_p),
    ]


class npt_packed_batch(C.Structure):
    _fields_ = [
        ("n", i64), ("pos", C.c_void_p), ("flag", C.c_void_p), ("src", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar16", C.c_void_p),
        ("n_wide", i64), ("wide_idx", C.c_void_p), ("wide_word", C.c_void_p), ("seq_off", C.c_void_p), ("seq2", C.c_void_p),
        ("n_exc", i64), ("exc_byte", C.c_void_p), ("exc_val", C.c_void_p),
    ]


class npt_results(C.Structure):
    _fields_ = [
        ("n_reads", i64), ("cluster_id", P(i32)), ("sub_id", P(i32)), ("start_pos", P(i32)), ("end_pos", P(i32)),
        ("read_st", P(i32)), ("read_en", P(i32)), ("read_flags", P(u32)), ("maps_sum", P(i32)), ("err_sum", P(i32)),
        ("break_off", P(u64)), ("breaks", P(i32)),
        ("n_clusters", i32), ("cl_first_read", P(i64)), ("cl_key_off", P(u32)), ("cl_key_tok", P(i32)),
        ("cl_read_count", P(i32)), ("cl_sub_off", P(u32)),
        ("n_subs", i32), ("sub_first_read", P(i64)), ("sub_key_off", P(u32)), ("sub_key", P(i32)), ("sub_count", P(i32)),
        ("sub_break_sum", P(i32)), ("sub_sum_off", P(u32)), ("sub_sums", P(i64)), ("sub_flags", P(u8)),
        ("total_counts", P(i32)), ("n_orf", i32), ("spliced", P(i32)), ("unspliced", P(i32)),
        ("n_pairs", i64), ("pair_src", P(i32)), ("pair_level", P(i32)), ("pair_start", P(i32)), ("pair_end", P(i32)),
        ("pair_count", P(i32)),
        ("cov_stride", i32), ("depth", P(i32)), ("errors", P(i32)), ("map_start", P(i32)), ("map_end", P(i32)),
        ("n_slow_reads", i64),
    ]


class npt_device_view(C.Structure):
    _fields_ = [("n_clusters", i32), ("cov_stride", i32), ("coverage", C.c_void_p), ("coverage_elems", i64),
                ("small_counts", C.c_void_p), ("small_elems", i64), ("n_subs", i32), ("n_pairs", i32),
                ("sub_count", C.c_void_p), ("sub_count_elems", i64), ("sub_break_sum", C.c_void_p), ("sub_break_sum_elems", i64),
                ("sub_sums", C.c_void_p), ("sub_sums_elems", i64), ("pair_count", C.c_void_p), ("pair_count_elems", i64)]


# every symbol include/npt.h declares (checked by tests/test_abi.py against the header text and the built .so)
SYMBOLS = [
    "npt_abi_version", "npt_create", "npt_destroy", "npt_last_error", "npt_begin_chrom", "npt_submit", "npt_submit_packed", "npt_wait",
    "npt_end_chrom", "npt_fetch_results", "npt_release_results", "npt_device_view_get",
    "npt_set_read_index_base", "npt_set_next_read_index", "npt_alloc_pinned", "npt_free_pinned", "npt_last_kernel_ms", "npt_last_step_ms", "npt_last_stage_ms",
    "npt_export_keys", "npt_import_keys",
    "npt_comm_unique_id", "npt_comm_init", "npt_end_chrom_all", "npt_last_exchange_ms",
    "npt_mg_create", "npt_mg_destroy", "npt_mg_last_error", "npt_mg_set_read_index_base", "npt_mg_begin_chrom", "npt_mg_submit", "npt_mg_submit_packed", "npt_mg_wait",
    "npt_mg_end_chrom", "npt_mg_fetch_results", "npt_mg_release_results", "npt_mg_last_step_ms",
]
NPT_COMM_ID_BYTES = 128


def declare(lib):
    """Attach argtypes/restypes to a loaded libnpt.so."""
    vp = C.c_void_p
    lib.npt_abi_version.restype = C.c_int
    lib.npt_create.argtypes = [P(npt_config), P(vp)]
    lib.npt_destroy.argtypes = [vp]
    lib.npt_destroy.restype = None
    lib.npt_last_error.argtypes = [vp]
    lib.npt_last_error.restype = C.c_char_p
    lib.npt_begin_chrom.argtypes = [vp, i32, vp, i64, P(npt_annotation)]
    lib.npt_submit.argtypes = [vp, P(npt_batch)]
    lib.npt_submit_packed.argtypes = [vp, P(npt_packed_batch)]
    lib.npt_wait.argtypes = [vp]
    lib.npt_end_chrom.argtypes = [vp]
    lib.npt_fetch_results.argtypes = [vp, C.c_int, P(npt_results)]
    lib.npt_release_results.argtypes = [vp]
    lib.npt_device_view_get.argtypes = [vp, P(npt_device_view)]
    lib.npt_set_read_index_base.argtypes = [vp, i64]
    lib.npt_set_next_read_index.argtypes = [vp, i64]
    lib.npt_set_next_read_index.restype = C.c_int
    lib.npt_alloc_pinned.argtypes = [C.c_size_t]
    lib.npt_alloc_pinned.restype = vp
    lib.npt_free_pinned.argtypes = [vp]
    lib.npt_free_pinned.restype = None
    lib.npt_last_kernel_ms.argtypes = [vp, P(C.c_float), P(C.c_float), P(i64), P(i64)]
    lib.npt_last_step_ms.argtypes = [vp, P(C.c_float)]
    lib.npt_last_stage_ms.argtypes = [vp, P(C.c_float)]
    lib.npt_last_stage_ms.restype = C.c_int
    lib.npt_export_keys.argtypes = [vp, P(vp), P(i64)]
    lib.npt_import_keys.argtypes = [vp, vp, i64]
    lib.npt_comm_unique_id.argtypes = [vp]
    lib.npt_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.npt_end_chrom_all.argtypes = [vp]
    lib.npt_last_exchange_ms.argtypes = [vp, P(C.c_float), P(C.c_float)]
    lib.npt_mg_create.argtypes = [P(npt_config), P(i32), i32, P(vp)]
    lib.npt_mg_destroy.argtypes = [vp]
    lib.npt_mg_destroy.restype = None
    lib.npt_mg_last_error.argtypes = [vp]
    lib.npt_mg_last_error.restype = C.c_char_p
    lib.npt_mg_set_read_index_base.argtypes = [vp, i64]
    lib.npt_mg_begin_chrom.argtypes = [vp, i32, vp, i64, P(npt_annotation)]
    lib.npt_mg_submit.argtypes = [vp, P(npt_batch)]
    lib.npt_mg_submit_packed.argtypes = [vp, P(npt_packed_batch)]
    lib.npt_mg_wait.argtypes = [vp]
    lib.npt_mg_end_chrom.argtypes = [vp]
    lib.npt_mg_fetch_results.argtypes = [vp, C.c_int, P(npt_results)]
    lib.npt_mg_release_results.argtypes = [vp]
    lib.npt_mg_last_step_ms.argtypes = [vp, P(C.c_float), P(C.c_float), P(C.c_float)]
    for s in ("npt_comm_unique_id", "npt_comm_init", "npt_end_chrom_all", "npt_last_exchange_ms", "npt_mg_create", "npt_mg_set_read_index_base",
              "npt_mg_begin_chrom", "npt_mg_submit", "npt_mg_submit_packed", "npt_mg_wait", "npt_mg_end_chrom", "npt_mg_fetch_results", "npt_mg_release_results",
              "npt_mg_last_step_ms"):
        getattr(lib, s).restype = C.c_int
    for s in ("npt_create", "npt_begin_chrom", "npt_submit", "npt_submit_packed", "npt_wait", "npt_end_chrom", "npt_fetch_results",
              "npt_release_results", "npt_device_view_get", "npt_set_read_index_base",
              "npt_last_kernel_ms", "npt_last_step_ms", "npt_export_keys", "npt_import_keys"):
        getattr(lib, s).restype = C.c_int
    return lib
