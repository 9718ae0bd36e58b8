// npt_device.cuh — device-side data structures shared by the kernels (sm_100a).
#pragma once
#include <cstdint>

namespace npt {

// counters[] slots (device uint32)
enum {
  CN_CLUSTERS = 0,   // dense cluster ids handed out (coverage row index)
  CN_TOK_USED,       // cluster record arena (ints)
  CN_SUBS,           // sub-cluster records created
  CN_SUBKEY_USED,    // sub-cluster record arena (ints)
  CN_SUMS_USED,      // int64 break-sum arena
  CN_PAIRS,          // distinct break-point pairs
  CN_ERROR,          // sticky error bits (ERR_*)
  CN_SLOW,           // windows that took the exact per-base path (diagnostic)
  CN_COUNT
};

enum { ERR_CL_FULL = 1, ERR_SUB_FULL = 2, ERR_PAIR_FULL = 4, ERR_COV_FULL = 8, ERR_ARENA_FULL = 16, ERR_TOO_MANY_BREAKS = 32, ERR_GENESET = 64, ERR_BAD_SRC = 128 };
constexpr int GFF_MAX_GENES = 32;  // distinct genes one read may hit (GFF mode)

// Hash-table slot (clusters and sub-clusters): `word` = 0 while empty, else (hash32 << 32) | record offset in the arena.
// The record is fully written before the CAS that publishes it, so readers never wait.  The slot index is the identity
// of the (sub-)cluster until end of chromosome.
struct Slot {
  unsigned long long word;
  unsigned long long min_idx;  // first-appearance global read index (atomicMin); sub-clusters: (index << 2) | first-read flags
};
// cluster record in cl_rec:  [0] dense id (coverage row; written by the creator right after the CAS)  [1] ntok  [2..] tokens
// sub record in sb_rec:      [0] cluster slot  [1] nkey  [2] offset into sub_sums  [3] nsum  [4..] binned breaks
constexpr int CL_REC_HDR = 2, SB_REC_HDR = 4;

struct PairEntry {  // break-point matrix cell  (src, level, start, end) -> count
  unsigned long long key;  // 0 = empty
  int count;
  int pad;
};

struct DevCfg {
  // ---- semantics
  int T, bin, include_start, coronavirus, start_thresh, end_thresh, enforce_strand, record_depth, calc_breaks1;
  int fit, track_sums, S, min_fl_exon;
  unsigned rna_mask;
  int annot_kind, n_orf;
  int G, stride;
  // ---- reference + annotation
  const uint32_t* ref_words;  // 8 bases per word, first base in bits 31..28; padded by 4 words
  int ref_nwords;             // words incl. padding
  const uint32_t* refg_words; // the same bases on the POSITION grid of the fast walk: word W = positions 8W..8W+7 (1-based,
  int refg_nwords;            // position 0 is padding), first in the top nibble; padded by 4 words
  uint32_t* fw_scratch;       // fast walk without depth recording: FW_BRK_BUF words per resident thread (long break lists)
  const int* orf_start;
  const int* orf_name;
  const uint8_t* orf_strand;
  // GFF mode: exon rows sorted by start: {start, end, (file row << 1) | strand, name id}; gx_hash is indexed by name id
  const int4* gx; int gx_n;
  const unsigned* gx_hash;
  // ---- tables
  Slot* cl; unsigned cl_mask; int cl_cap;
  int* cl_rec; unsigned tok_cap;
  Slot* sb; unsigned sb_mask; int sb_cap;
  int* sb_rec; unsigned sbkey_cap;
  int* sub_count;            // [sb slots][S]
  int* sub_break_sum;        // [sb slots]
  unsigned long long* sub_min0;  // [sb slots] firstIsTranscriptome: first global index of a source-0 read of the sub-cluster
  unsigned long long* sub_sums; unsigned sums_cap;
  PairEntry* pr; unsigned pr_mask; int pr_cap;
  int* cov; int cov_cap;     // [9][cov_cap][S][stride]: depth-diff, errors, mapStart, mapEnd, deletions of length 1..4 by first deleted
                             // position (serial walker), absolute depth credited by accumulate_kernel (fast walk)
  int* small;                // total_counts[S], spliced[S][n_orf], unspliced[S][n_orf]
  unsigned* counters;
};

struct BatchDev {
  long long n;
  long long idx_base;  // global submit index of read 0 of this batch
  const int* pos;
  const uint16_t* flag;
  const uint16_t* src;
  const uint32_t* cigar_off;
  const uint32_t* cigar;
  const unsigned long long* seq_off;
  const uint8_t* seq4;
  long long seq_bytes;
  long long cigar_words;
  // outputs
  int* cluster; int* sub; int* start_pos; int* end_pos; int* read_st; int* read_en; unsigned* rflags;
  int* maps_sum; int* err_sum; int* nbreaks; unsigned* break_loc;
  int* break_arena;     // [n][BRK_INLINE] inline slots, then break_cap ints of overflow area
  unsigned break_cap;
  unsigned* bctr;       // per batch: [0] overflow-area cursor, [1] overflow-area full flag, [2..4] read cursors of walk_kernel<0/1/2>,
                        // [5] read cursor of fast_walk_kernel, [6] tile cursor of accumulate_kernel, [7] reads in slow_list,
                        // [8] overflow-area full flag of the serial walker's chain when it runs beside a fast walk (12 words)
  // fast walk (fastwalk.cuh): when `fast` is set every read is first offered to fast_walk_kernel and only the reads it lists
  // in slow_list are walked by walk_kernel<0/1/2>
  int fast;
  unsigned* slow_list;  // [n] read indices flagged NPT_RF_SLOWPATH
  uint4* fhdrA;         // [n] first entry of the read in the arena + its (up to three) blocks of entries
  unsigned* fmeta;      // [n] nblk | njunc << 8 | nextra << 16 | ndbl << 24
  unsigned* frare;      // [n][12] rare per-read events (junction points, extras; fastwalk.cuh), valid where fmeta says so
  uint4* fent;          // entry arena: 32 positions x 4 flag bits per entry
};

}  // namespace npt
