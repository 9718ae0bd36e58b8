/*
 * npt.h — C ABI of the B200-native npTranscript read→transcript-cluster hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own; the seam this
 * ABI replaces is the three-method surface of IdentityProfileHolder
 * (J = /root/reference/java/src/main/java/npTranscript):
 *
 *   npt_create        <- the static configuration copied in J/run/ViralTranscriptAnalysisCmd2.java:221-386, 496-511
 *   npt_begin_chrom   <- new Annotation(...) + new IdentityProfileHolder(genomes, chr, outp, in_nmes,
 *                        calcBreaks1, currentIndex, annot)   J/run/ViralTranscriptAnalysisCmd2.java:872-886,
 *                        J/cluster/IdentityProfileHolder.java:106-128
 *   npt_submit        <- profile.identity1(readSeq, sam, source_index, ...) called once per surviving record
 *                        J/run/ViralTranscriptAnalysisCmd2.java:967, J/cluster/IdentityProfileHolder.java:170-196
 *                        (batched: one call carries n records as structure-of-arrays)
 *   npt_end_chrom     <- waitOnThreads + printBreakPoints() + getConsensus()
 *                        J/run/ViralTranscriptAnalysisCmd2.java:840-843, 989-991,
 *                        J/cluster/IdentityProfileHolder.java:101-104, 152-157
 *   npt_fetch_results <- what the writers consume: Outputs.printRead rows (J/cluster/IdentityProfile1.java:128-144),
 *                        CigarClusters.process1 (J/cluster/CigarClusters.java:138-196), Outputs.writeDepthH5
 *                        (J/cluster/Outputs.java:559-593), Outputs.printMatrix (J/cluster/Outputs.java:461-489),
 *                        Annotation.print (J/cluster/Annotation.java:40-58)
 *   npt_destroy       <- IdentityProfileHolder.shutDownExecutor()  J/cluster/IdentityProfileHolder.java:198-204
 *
 * Semantics: results are defined as if the submitted records were processed one at a time in
 * submit order (the reference's --maxThreads=1 behaviour; with its default thread pool the
 * reference's own cluster numbering is a race, J/cluster/IdentityProfileHolder.java:180-193).
 *
 * There is no CPU fallback: every entry point that computes needs an sm_100 device and
 * returns NPT_ENODEVICE without one.
 *
 * Plain C: pointers and sizes only.  All integers little-endian host order.
 */
#ifndef NPT_H
#define NPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPT_ABI_VERSION 1
#define NPT_MAX_SOURCES 16

/* error codes (every call returns 0 or one of these; text via npt_last_error) */
#define NPT_OK 0
#define NPT_EINVAL -1     /* bad argument / unsupported flag combination          */
#define NPT_ENODEVICE -2  /* no usable CUDA device (no CPU fallback)              */
#define NPT_ECUDA -3      /* CUDA runtime error                                   */
#define NPT_ENOMEM -4     /* host or device allocation failed                     */
#define NPT_ECAPACITY -5  /* a table (clusters, sub-clusters, pairs, ...) is full */
#define NPT_ESTATE -6     /* call out of order                                    */

/* annotation kinds: J/cluster/Annotation.java (CSV ORF table), J/cluster/EmptyAnnotation.java */
#define NPT_ANNOT_CSV 0
#define NPT_ANNOT_EMPTY 1
#define NPT_ANNOT_GFF 2   /* exon table of one chromosome, J/cluster/GFFAnnotation.java:368-459 (host mode, coronavirus=0) */

/* where the arrays of an npt_batch live */
#define NPT_MEM_HOST 0
#define NPT_MEM_DEVICE 1

/* key tokens.  A cluster key (CigarHash.secondKey, J/cluster/CigarHash.java:42-62) is a string built
 * from annotation names; on the device it is the list of these int32 tokens and the host renders it.
 *   token >= 0          CSV: canonical name id of the ORF row; EMPTY: floor(pos/bin)
 *   NPT_TOK_ST / _END   the "st"/"end" fall-through names of J/cluster/Annotation.java:71,96
 *   NPT_TOK_PLUS/_MINUS the '+'/'-' suffix appended when enforceStrand (J/cluster/IdentityProfile1.java:325-327)
 * The position inside the list decides the delimiter (J/cluster/IdentityProfile1.java:266-292). */
#define NPT_TOK_ST (-1)
#define NPT_TOK_END (-2)
#define NPT_TOK_PLUS (-3)
#define NPT_TOK_MINUS (-4)
/* first token when the 5'/3' end names are left out of the key (!includeStart and more than two breaks,
 * IdentityProfile1.java:267,289): "a,b_" (4 breaks) must not collide with "a_b" (2 breaks). */
#define NPT_TOK_NOENDS (-5)
/* GFF mode: first token of a key made of gene names; the following tokens are name ids in the iteration order of the
 * reference's HashSet<String> (J/cluster/IdentityProfile1.java:301-318), each rendered as name + '_'.  A key without the
 * marker is the single coordinate token floor(pos/bin) rendered "<chr>.<bin>" (GFFAnnotation.getCoord :132-135). */
#define NPT_TOK_GENESET (-6)

/* per-read flag bits in npt_results.read_flags */
#define NPT_RF_TYPE_MASK 0x3u  /* Annotation.getTypeInd: 0=5_3 1=5_no3 2=no5_3 3=no5_no3 (J/cluster/Annotation.java:233-236) */
#define NPT_RF_LEADER 0x4u     /* hasLeaderBreak (J/cluster/IdentityProfile1.java:256, 281) */
#define NPT_RF_NEG 0x8u        /* read negative strand flag (SAM flag 0x10) */
#define NPT_RF_BADCIGAR 0x10u  /* pos <= 0, CIGAR op > 8 (J/cluster/IdentityProfile1.java:577-578) or an M element past the end of the read
                                 (getBase, :531-551): the reference throws; read not clustered */
#define NPT_RF_SLOWPATH 0x20u  /* diagnostic: read was walked by the exact serial device path */

typedef struct npt_config {
  int32_t abi_version;            /* NPT_ABI_VERSION */
  int32_t bin;                    /* --bin: CigarHash2.round = Annotation.tolerance (ViralTranscriptAnalysisCmd2.java:496-497) */
  int32_t break_thresh;           /* --breakThresh: IdentityProfile1.break_thresh (:565) */
  int32_t include_start;          /* --includeStart: IdentityProfile1.includeStartEnd (:300) */
  int32_t coronavirus;            /* --coronavirus (:302); also selects min_first_last_exon_length 0/10 (:339,360) */
  int32_t start_thresh;           /* --startThresh (:508) */
  int32_t end_thresh;             /* --endThresh (:509) */
  int32_t enforce_strand;         /* --enforceStrand: Annotation.enforceStrand (:241) */
  int32_t record_depth;           /* --recordDepthByPosition: CigarCluster.recordDepthByPosition = recordStartEnd (:333-334) */
  int32_t calc_breaks;            /* calcBreaks (:346,370); the library additionally requires break_thresh < seqlen (:882) */
  int32_t first_is_transcriptome; /* --firstIsTranscriptome: Outputs.firstIsTranscriptome (:376) */
  int32_t track_break_sums;       /* writeGFF || writeIsoforms: Count.addBreaks is called (CigarCluster.java:384,676) */
  int32_t num_sources;            /* number of input BAMs (in_nmes.length) */
  int32_t rna[NPT_MAX_SOURCES];   /* ViralTranscriptAnalysisCmd2.RNA[src] (:243-257) */
  int32_t device;                 /* CUDA device ordinal */
  /* capacities of the device tables; 0 = library default.  Exceeding one is NPT_ECAPACITY, never a silent drop. */
  int32_t max_clusters;
  int32_t max_subclusters;
  int32_t max_break_pairs;
  int32_t max_coverage_clusters;  /* clusters that get dense coverage rows when record_depth */
} npt_config;

/* ORF table of one chromosome (J/cluster/Annotation.java:138-179) as parallel arrays, in file row order
 * (rows with Direction=both are already duplicated by the caller when enforce_strand, :160-164).
 * NPT_ANNOT_GFF: one row per exon line of the chromosome in file order (start, end, strand, gene = the exon's parent
 * attribute, J/cluster/GFFAnnotation.java:409-428); up to 2^22 rows. */
typedef struct npt_annotation {
  int32_t kind;            /* NPT_ANNOT_* */
  int32_t n;               /* rows (0 for EMPTY) */
  const int32_t* start;    /* Minimum */
  const int32_t* end;      /* Maximum */
  const uint8_t* strand;   /* 1 = forward */
  const int32_t* name_id;  /* rows with equal gene strings share an id (index of first row with that string) */
  const uint32_t* name_hash; /* GFF only, per row: h ^ (h >>> 16) of Java's String.hashCode() of the row's gene string
                              (decides HashSet iteration order = order of the names inside the key); NULL otherwise */
} npt_annotation;

/* One batch of surviving records (after the filter chain of ViralTranscriptAnalysisCmd2.java:712-959),
 * structure-of-arrays, exactly the fields the path reads from SAMRecord:
 *   pos        sam.getAlignmentStart(), 1-based
 *   flag       SAM flag word (only 0x10 is read: getReadNegativeStrandFlag)
 *   src        source_index (SequenceUtils.src_tag, :756)
 *   cigar      BAM encoding (len<<4)|op, op in MIDNSHP=X = 0..8
 *   seq4       BAM 4-bit bases "=ACMGRSVTWYHKDBN", two per byte, high nibble first, each read byte-aligned
 * cigar_off/seq_off are batch-relative element/byte offsets with n+1 entries. */
typedef struct npt_batch {
  int64_t n;
  int32_t location;        /* NPT_MEM_HOST or NPT_MEM_DEVICE */
  int32_t reserved;
  const int32_t* pos;
  const uint16_t* flag;
  const uint16_t* src;
  const uint32_t* cigar_off;
  const uint32_t* cigar;
  const uint64_t* seq_off;
  const uint8_t* seq4;
} npt_batch;

/* The same records in the compact transfer format (host memory only).  The end-to-end path is bound by the host-to-device
 * link (1 681 B per SARS-CoV-2 dRNA read as npt_batch); this form moves 852 B per read and is expanded into the npt_batch
 * layout by two copy kernels on the device (unpack.cuh), bit-identical to submitting the expanded arrays:
 *   cigar16   one 16-bit word per CIGAR element: (len<<4)|op for len < 4095, else 0xFFF0|op and the element appears in the
 *             wide list (wide_idx = element index in the batch, ascending; wide_word = its full 32-bit BAM word)
 *   seq2      two bits per base, A C G T = 0 1 2 3, four bases per byte, first base in the top bits; position-aligned with
 *             seq4 (nibble k of the seq4 stream = base k of the seq2 stream, so seq2 has ceil(seq_off[n] / 2) bytes)
 *   exc_byte / exc_val   seq4 bytes that hold any code other than A C G T (N, IUPAC, '=', the pad nibble of an odd-length
 *             read when it is not one of 1 2 4 8): index into the seq4 stream and the byte to put there
 * cigar_off / seq_off are those of the equivalent npt_batch (elements / seq4 bytes).  The packer for BAM records is
 * nptbam (include/npt_bam.h); nptranscript_b200/packed.py is the numpy one the tests and bench use. */
typedef struct npt_packed_batch {
  int64_t n;
  const int32_t* pos;
  const uint16_t* flag;
  const uint16_t* src;
  const uint32_t* cigar_off;
  const uint16_t* cigar16;
  int64_t n_wide;
  const uint32_t* wide_idx;
  const uint32_t* wide_word;
  const uint64_t* seq_off;
  const uint8_t* seq2;
  int64_t n_exc;
  const uint64_t* exc_byte;
  const uint8_t* exc_val;
} npt_packed_batch;

/* Everything the reference's writers need, in host memory owned by the context
 * (valid until npt_release_results / npt_begin_chrom / npt_destroy). */
typedef struct npt_results {
  /* ---- per read, in submit order (the 0.readToCluster row, IdentityProfile1.java:133-136) ---- */
  int64_t n_reads;
  const int32_t* cluster_id;    /* n of "ID<chr>.<n>": first-appearance rank (CigarClusters.java:83); -1 if not clustered */
  const int32_t* sub_id;        /* k of ".t<k>": first-appearance rank inside the cluster (CigarCluster.java:660) */
  const int32_t* start_pos;     /* breaks[0] after trimming */
  const int32_t* end_pos;       /* breaks[last] */
  const int32_t* read_st;       /* sam.getReadPositionAtReferencePosition(alignmentStart) (:634) */
  const int32_t* read_en;       /* ... (alignmentEnd) (:635) */
  const uint32_t* read_flags;   /* NPT_RF_* */
  const int32_t* maps_sum;      /* maps[src].valsum of the read's scratch (errorRatio denominator); 0 if !record_depth */
  const int32_t* err_sum;       /* errors[src].valsum (numerator) */
  const uint64_t* break_off;    /* n_reads+1 offsets into breaks */
  const int32_t* breaks;        /* CigarCluster.breaks of each read (1-based positions) */
  /* ---- per cluster, index = cluster_id ---- */
  int32_t n_clusters;
  const int64_t* cl_first_read; /* global submit index of the read that created it */
  const uint32_t* cl_key_off;   /* n_clusters+1 offsets into cl_key_tok */
  const int32_t* cl_key_tok;    /* key tokens */
  const int32_t* cl_read_count; /* [n_clusters][num_sources] readCount */
  const uint32_t* cl_sub_off;   /* n_clusters+1 offsets into the sub-cluster arrays (subs of a cluster are contiguous, in .t order) */
  /* ---- per sub-cluster ---- */
  int32_t n_subs;
  const int64_t* sub_first_read;
  const uint32_t* sub_key_off;  /* n_subs+1 offsets into sub_key */
  const int32_t* sub_key;       /* binned breaks, CigarCluster.cloneBreaks (CigarCluster.java:620-631) */
  const int32_t* sub_count;     /* [n_subs][num_sources] Count.count */
  const int32_t* sub_break_sum; /* Count.break_sum */
  const uint32_t* sub_sum_off;  /* n_subs+1 offsets into sub_sums */
  const int64_t* sub_sums;      /* exact integer sums of the raw breaks of the reads counted in break_sum */
  const uint8_t* sub_flags;     /* bit0: forward of the first read, bit1: neg-strand flag of the first read */
  /* ---- per chromosome aggregates ---- */
  const int32_t* total_counts;  /* [num_sources] CigarClusters.totalCounts */
  int32_t n_orf;
  const int32_t* spliced;       /* [num_sources][n_orf] Annotation.spliced_count */
  const int32_t* unspliced;     /* [num_sources][n_orf] */
  int64_t n_pairs;              /* break-point matrix entries (BreakPoints.java:102-108), sorted by (src, level, start, end) */
  const int32_t* pair_src;
  const int32_t* pair_level;
  const int32_t* pair_start;
  const int32_t* pair_end;
  const int32_t* pair_count;
  /* ---- coverage (record_depth only): dense rows of seqlen+2 ints, index = 1-based position ---- */
  int32_t cov_stride;           /* seqlen + 2 */
  const int32_t* depth;         /* [n_clusters][num_sources][cov_stride] maps[src] */
  const int32_t* errors;        /* [n_clusters][num_sources][cov_stride] errors[src] */
  const int32_t* map_start;     /* [n_clusters][num_sources][cov_stride] mapStart */
  const int32_t* map_end;       /* [n_clusters][num_sources][cov_stride] mapEnd */
  int64_t n_slow_reads;         /* diagnostic: reads that took the serial device path */
} npt_results;

typedef struct npt_ctx npt_ctx;

int npt_abi_version(void);
int npt_create(const npt_config* cfg, npt_ctx** out);
void npt_destroy(npt_ctx* ctx);
const char* npt_last_error(const npt_ctx* ctx); /* ctx may be NULL: error of the last failed npt_create on this thread */

/* ref_codes: one BAM 4-bit code per reference base (same code table as seq4), host memory. */
int npt_begin_chrom(npt_ctx* ctx, int32_t chrom_index, const uint8_t* ref_codes, int64_t ref_len, const npt_annotation* annot);
/* Asynchronous: returns once the batch is enqueued.  Host arrays (NPT_MEM_HOST) must stay untouched until npt_wait.
 * Device arrays (NPT_MEM_DEVICE) are read in place and must stay valid and unchanged until npt_end_chrom has returned
 * (with first_is_transcriptome the end-of-chromosome pass reads src again).  A record whose src is >= num_sources makes
 * npt_end_chrom fail with NPT_EINVAL. */
int npt_submit(npt_ctx* ctx, const npt_batch* batch);
/* npt_submit for a batch in the compact transfer format (host arrays, untouched until npt_wait). */
int npt_submit_packed(npt_ctx* ctx, const npt_packed_batch* batch);
int npt_wait(npt_ctx* ctx);
/* Drains, assigns first-appearance ids (sort-and-rank), rewrites per-read ids, scans the coverage difference arrays. */
int npt_end_chrom(npt_ctx* ctx);
/* Copies results to host buffers owned by ctx.  what = bitmask of NPT_FETCH_*. */
#define NPT_FETCH_READS 1
#define NPT_FETCH_TABLES 2
#define NPT_FETCH_COVERAGE 4
#define NPT_FETCH_ALL 7
int npt_fetch_results(npt_ctx* ctx, int what, npt_results* out);
int npt_release_results(npt_ctx* ctx);

/* Multi-GPU (reads shard across ranks in contiguous ranges of the submit order, npt_set_read_index_base).
 *
 * Device-side protocol, one exchange of keys and one sum of counters (nptranscript_b200/multigpu.py drives it over NCCL):
 *   every rank:  npt_submit ...; npt_export_keys -> packed int32 device buffer with the keys of its cluster, sub-cluster
 *                and break-point-pair tables and their first-appearance indices;
 *   all-gather the buffers; npt_import_keys(buffer of every OTHER rank): the keys are inserted without counts, so every
 *                rank now holds the same key set and the same minima;
 *   npt_end_chrom: the sort-and-rank numbers clusters and sub-clusters identically on every rank (CigarClusters.java:83,
 *                CigarCluster.java:660 as if all reads had gone through one process in global submit order); per-read ids
 *                are global;
 *   all-reduce (sum) the device arrays of npt_device_view in place; then npt_fetch_results returns the global tables. */
typedef struct npt_device_view {
  int32_t n_clusters;           /* clusters after npt_end_chrom */
  int32_t cov_stride;
  void* coverage;               /* device int32 [4][n_clusters][num_sources][cov_stride], rows in cluster-id order */
  int64_t coverage_elems;
  void* small_counts;           /* device int32: total_counts[S], spliced[S][n_orf], unspliced[S][n_orf] */
  int64_t small_elems;
  int32_t n_subs, n_pairs;
  void* sub_count;              /* device int32 [n_subs][num_sources], sub-clusters in result order */
  int64_t sub_count_elems;
  void* sub_break_sum;          /* device int32 [n_subs] */
  int64_t sub_break_sum_elems;
  void* sub_sums;               /* device int64 [sub_sum_off[n_subs]] */
  int64_t sub_sums_elems;
  void* pair_count;             /* device int32 [n_pairs], pairs in result order (sorted by src, level, start, end) */
  int64_t pair_count_elems;
} npt_device_view;
int npt_export_keys(npt_ctx* ctx, const int32_t** dev_buf, int64_t* n_words);
int npt_import_keys(npt_ctx* ctx, const int32_t* dev_buf, int64_t n_words);
/* The same protocol with the collectives inside the library (NCCL, loaded at run time with dlopen("libnccl.so.2"); csrc/comm.inl).
 * Multi-process, one rank per GPU: rank 0 calls npt_comm_unique_id and hands the 128 bytes to the other ranks by any means
 * (MPI, torch.distributed, a file); every rank calls npt_comm_init on its context; then npt_end_chrom_all replaces
 * npt_end_chrom: export -> all-gather of the headers -> one group of broadcasts -> import kernels -> npt_end_chrom ->
 * all-reduce of the counters, with two host waits in total.  npt_last_step_ms then covers the whole span. */
#define NPT_COMM_ID_BYTES 128
int npt_comm_unique_id(void* id128);
int npt_comm_init(npt_ctx* ctx, const void* id128, int rank, int world);
int npt_end_chrom_all(npt_ctx* ctx);
int npt_last_exchange_ms(npt_ctx* ctx, float* key_exchange_ms, float* reduce_ms);

/* In-process multi-GPU context: what a single-threaded host (the Java shim: one caller, one holder,
 * J/cluster/IdentityProfileHolder.java:170-196) drives.  One npt_ctx per listed device, one host thread per device inside
 * every call, NCCL communicators from ncclCommInitAll.  Each host batch is split into contiguous read ranges with about
 * the same number of bases, one per device; results are defined exactly as for one context (sequential in submit order).
 * npt_mg_fetch_results returns the global tables / coverage once and the per-read arrays of all devices back in submit order.
 * A device may be listed more than once (several contexts on one GPU; used by the single-GPU tests): the collectives then
 * run as device-to-device copies and an add kernel instead of NCCL. */
typedef struct npt_mg npt_mg;
int npt_mg_create(const npt_config* cfg, const int32_t* devices, int32_t n_devices, npt_mg** out);   /* cfg->device is ignored */
void npt_mg_destroy(npt_mg* mg);
const char* npt_mg_last_error(const npt_mg* mg);   /* mg may be NULL: error of the last failed npt_mg_create on this thread */
int npt_mg_set_read_index_base(npt_mg* mg, int64_t base);
int npt_mg_begin_chrom(npt_mg* mg, int32_t chrom_index, const uint8_t* ref_codes, int64_t ref_len, const npt_annotation* annot);
int npt_mg_submit(npt_mg* mg, const npt_batch* batch);   /* NPT_MEM_HOST only; arrays untouched until npt_mg_wait / npt_mg_end_chrom */
int npt_mg_submit_packed(npt_mg* mg, const npt_packed_batch* batch);   /* the same in the compact transfer format */
int npt_mg_wait(npt_mg* mg);
int npt_mg_end_chrom(npt_mg* mg);
int npt_mg_fetch_results(npt_mg* mg, int what, npt_results* out);
int npt_mg_release_results(npt_mg* mg);
int npt_mg_last_step_ms(npt_mg* mg, float* step_ms, float* key_exchange_ms, float* reduce_ms);   /* max over the devices */

int npt_device_view_get(npt_ctx* ctx, npt_device_view* out);
/* Offset added to the batch-relative read index to form the global submit index (sharded runs). */
int npt_set_read_index_base(npt_ctx* ctx, int64_t base);
/* Global submit index of the first read of the NEXT npt_submit (any time inside a chromosome): sharded runs whose ranks
 * take turns in the submit order, e.g. wave k of rank r starts at (k * ranks + r) * wave_size. */
int npt_set_next_read_index(npt_ctx* ctx, int64_t index);

/* pinned host memory for batches (cudaHostAlloc) */
void* npt_alloc_pinned(size_t bytes);
void npt_free_pinned(void* p);

/* timing of the device work of the last npt_submit..npt_end_chrom span, measured with CUDA events
 * on the library's compute stream (milliseconds); used by bench.py for the roofline numbers. */
int npt_last_kernel_ms(npt_ctx* ctx, float* walk_ms, float* finalize_ms, int64_t* walk_launches, int64_t* total_launches);
/* CUDA-event span on the compute stream from the first device work of npt_begin_chrom to the last of npt_end_chrom. */
int npt_last_step_ms(npt_ctx* ctx, float* step_ms);
/* Per-stage CUDA-event durations summed over the batches of the last chromosome, in launch order:
 * out6 = { fast_walk_kernel (breaks + coverage entries of the reads it can restate), walk_kernel<0> (serial walker: all
 * reads when the fast walk is off, else the reads it refused), walk_kernel<2> (long break lists), cluster_kernel,
 * walk_kernel<1> (coverage of the serial walker's reads), coverage of the fast reads (keys + radix sort + accumulate_kernel) }. */
int npt_last_stage_ms(npt_ctx* ctx, float* out6);

#ifdef __cplusplus
}
#endif
#endif /* NPT_H */
