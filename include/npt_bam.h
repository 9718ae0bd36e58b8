/* npt_bam.h — BAM (BGZF) -> structure-of-arrays packer feeding npt_submit.
 *
 * Replaces, for the hot path's caller, what the reference does per record between opening the BAM and
 * profile.identity1(...):
 *   htsjdk SamReaderFactory...open(bam).iterator()                J/run/ViralTranscriptAnalysisCmd2.java:623-627
 *   the per-record filter chain, in this order                   J/run/ViralTranscriptAnalysisCmd2.java:712-830, 936-959
 *     unmapped flag (:716) -> chromosome include list (:721-730) -> mean-error-probability quality q1 from the base
 *     qualities, dropped if q1 < fail_thresh (:738-752) -> read length <= 1 (:797) -> numReads++ / maxReads stop (:804-808)
 *     -> MAPQ < --qual (:813) -> secondary / supplementary alignments dropped (allowSuppAlignments is hard-wired false,
 *     :394, :938-959)
 *   chromosome switch on a change of reference index (:830): a batch never spans two reference sequences.
 * Not restated (host-side options that default to off): readList / pattern name filters (:732-737, :759), probInclude
 * random sub-sampling (:817), the fastq re-mapping branch (:904-935).
 *
 * BAM stores CIGAR as (len<<4)|op words and SEQ as 4-bit codes, two per byte, high nibble first — the layout npt_batch
 * takes — so packing is one memcpy per field.  BGZF blocks are inflated by a pool of threads (zlib raw inflate).
 *
 * Host code only: no CUDA, no dependency on libnpt.  Buffers come from malloc, or from the allocator pair given to
 * nptb_open (pass npt_alloc_pinned / npt_free_pinned for page-locked batches).
 */
#ifndef NPT_BAM_H
#define NPT_BAM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nptb_reader nptb_reader;

typedef struct nptb_filter {
  double fail_thresh;          /* --fail_thresh (first value): records with q1 < fail_thresh are dropped; 0 keeps all */
  int32_t min_mapq;            /* --qual */
  int64_t max_reads;           /* --maxReads: stop after this many records passed the filters up to the length test; <=0 = no limit */
  const uint8_t* ref_include;  /* optional [n_refs]: 0 = skip records of that reference (chromsToInclude); NULL = all */
} nptb_filter;

typedef struct nptb_batch {
  int32_t ref_id;              /* reference index of every record of the batch */
  int64_t n;
  /* npt_batch fields (host memory owned by the reader, valid until the next nptb_next_batch / nptb_close) */
  const int32_t* pos;          /* 1-based alignment start (BAM pos + 1 = SAMRecord.getAlignmentStart) */
  const uint16_t* flag;
  const uint16_t* src;         /* the src given to nptb_next_batch (SequenceUtils.src_tag) */
  const uint32_t* cigar_off;   /* [n+1] */
  const uint32_t* cigar;
  const uint64_t* seq_off;     /* [n+1], bytes */
  const uint8_t* seq4;
  /* what the host keeps for its writers */
  const double* q1;            /* -10*log10(mean(10^(-q/10))) over the base qualities, 500 when they are absent (:739-746) */
  const uint8_t* mapq;
  const uint32_t* name_off;    /* [n+1] */
  const char* names;           /* read names, not NUL-terminated */
  /* running totals since nptb_open */
  int64_t n_records, n_unmapped, n_excluded_ref, n_low_quality, n_short, n_low_mapq, n_secondary;
  int32_t stopped_at_max_reads;
} nptb_batch;

/* nthreads <= 0: hardware concurrency.  alloc/free_: NULL = malloc/free.  Returns NULL on error (message in err). */
nptb_reader* nptb_open(const char* path, int nthreads, void* (*alloc)(size_t), void (*free_)(void*), char* err, size_t errlen);
void nptb_close(nptb_reader* r);
int32_t nptb_n_refs(const nptb_reader* r);
const char* nptb_ref_name(const nptb_reader* r, int32_t i);
int64_t nptb_ref_len(const nptb_reader* r, int32_t i);
/* Next batch of up to max_batch_reads records that pass the filters, all of one reference.  Returns 1 with *out filled,
 * 0 at end of file (or after max_reads), < 0 on a malformed file (message from nptb_last_error). */
int nptb_next_batch(nptb_reader* r, const nptb_filter* f, int32_t src, int64_t max_batch_reads, nptb_batch* out);
const char* nptb_last_error(const nptb_reader* r);

/* The batch in the compact transfer format of npt_submit_packed (include/npt.h): 16-bit CIGAR words + the list of elements of
 * 4 095 or longer, 2-bit bases + the list of seq4 bytes that hold anything but A C G T.  Field for field the layout of
 * npt_packed_batch (tests/test_bam_pack.py checks the offsets), so that a pointer to it can be handed to npt_submit_packed.
 * Buffers belong to the reader (same allocator as the batches: pinned when nptb_open got npt_alloc_pinned) and are valid
 * until the next nptb_pack / nptb_close; pos, flag, src, cigar_off and seq_off are the batch's own arrays. */
typedef struct nptb_packed {
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
} nptb_packed;
int nptb_pack(nptb_reader* r, const nptb_batch* b, nptb_packed* out);

/* Test hook: one raw DEFLATE stream (the payload of a BGZF block) through the packer's own table-driven decoder (mode 1,
 * bam/fast_inflate.h: it refuses what it does not handle and the packer then uses zlib) or through zlib (mode 0).
 * Returns the number of bytes written, or -1. */
long nptb_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, int mode);

#ifdef __cplusplus
}
#endif
#endif
