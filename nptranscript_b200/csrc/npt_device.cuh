// npt_device.cuh — device-side data structures shared by the kernels (sm_100a).
#pragma once
#include <cstdint>

namespace npt {

// counters[] slots (device uint32 unless noted)
enum {
  CN_CLUSTERS = 0,   // provisional cluster ids handed out
  CN_TOK_USED,       // cluster key token arena
  CN_SUBS,           // provisional sub-cluster ids handed out
  CN_SUBKEY_USED,    // sub-cluster key arena
  CN_SUMS_USED,      // int64 break-sum arena
  CN_PAIRS,          // distinct break-point pairs
  CN_ERROR,          // sticky error bits (ERR_*)
  CN_SLOW,           // reads that took the serial path
  CN_COUNT
};

enum { ERR_CL_FULL = 1, ERR_SUB_FULL = 2, ERR_PAIR_FULL = 4, ERR_COV_FULL = 8, ERR_ARENA_FULL = 16, ERR_TOO_MANY_BREAKS = 32 };

struct ClEntry {   // cluster hash-table slot (key = token list)
  unsigned long long tag;      // 0 = empty, else 64-bit key hash (never 0)
  unsigned long long min_idx;  // first-appearance global read index (atomicMin)
  int prov;                    // provisional dense id, -1 until published
  unsigned tok_off, ntok;
  unsigned pad;
};

struct SubEntry {  // sub-cluster hash-table slot (key = cluster prov + binned break list)
  unsigned long long tag;
  unsigned long long min_idx;
  int prov;
  unsigned key_off, nkey;
  int cl_prov;
};

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
  int force_serial;  // T too small for the windowed fast path
  // ---- reference + annotation
  const uint32_t* ref_words;  // 8 bases per word, first base in bits 31..28; padded by 4 words
  int ref_nwords;             // words incl. padding
  const int* orf_start;
  const int* orf_name;
  const uint8_t* orf_strand;
  // ---- tables
  ClEntry* cl; unsigned cl_mask; int cl_cap;
  int* cl_tok; unsigned tok_cap;
  SubEntry* sb; unsigned sb_mask; int sb_cap;
  int* sb_key; unsigned sbkey_cap;
  int* sub_count;            // [sb_cap][S]
  int* sub_break_sum;        // [sb_cap]
  unsigned* sub_sum_off;     // [sb_cap]
  unsigned* sub_nsum;        // [sb_cap]
  unsigned char* sub_flags;  // [sb_cap]
  unsigned long long* sub_sums; unsigned sums_cap;
  PairEntry* pr; unsigned pr_mask; int pr_cap;
  int* cov; int cov_cap;     // [4][cov_cap][S][stride]: depth-diff, errors, mapStart, mapEnd
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
  // outputs
  int* cluster; int* sub; int* start_pos; int* end_pos; int* read_st; int* read_en; unsigned* rflags;
  int* maps_sum; int* err_sum; int* nbreaks; unsigned* break_loc;
  int* break_arena;     // [n][BRK_INLINE] inline slots, then break_cap ints of overflow area
  unsigned break_cap;
  unsigned* bctr;       // per batch: [0] overflow-area cursor, [1] overflow-area full flag
};

}  // namespace npt
