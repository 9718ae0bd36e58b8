// npt.cu — host side of libnpt.so: context, device memory, batch staging, end-of-chromosome finalisation, C ABI.
// Product code.  Never includes or links anything under oracle/.  No CPU compute fallback: without an sm_100 device
// every computing entry point fails with NPT_ENODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/npt.h"
#include "npt_device.cuh"
#include "kernels.cuh"
#include "fastcov.cuh"
#include "unpack.cuh"

using namespace npt;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DBuf {  // grow-only device buffer
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct HBuf {  // grow-only pinned host buffer
  T* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap && p) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    if (cudaHostAlloc(&p, std::max<size_t>(n, 1) * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return false; }
    cap = std::max<size_t>(n, 1);
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Staging {  // device copy of one host batch
  DBuf<unsigned char> buf;
  cudaEvent_t done = nullptr;  // kernel that consumed it finished
  bool busy = false;
  int batch_index = -1;
};

struct Batch {
  BatchDev d{};
  int64_t n = 0;
  void* out_block = nullptr;  // one allocation for all per-read outputs
  size_t out_bytes = 0, arena_bytes = 0, bctr_bytes = 0;
  unsigned* bctr = nullptr;
  int staging = -1;           // index into ctx->stg, or -1 for caller-owned device memory
  bool retired = false;
  cudaEvent_t done = nullptr;
  // fast walk scratch (returned to the pool when the batch retires): slow list, headers, entry arena, sort keys / values
  void* f_slow = nullptr; size_t f_slow_bytes = 0;
  void* f_hdr = nullptr; size_t f_hdr_bytes = 0;
  void* f_ent = nullptr; size_t f_ent_bytes = 0;
  void* f_sort = nullptr; size_t f_sort_bytes = 0;   // keys, vals, sorted keys, sorted vals, cub temp
  unsigned *k_in = nullptr, *v_in = nullptr, *k_out = nullptr, *v_out = nullptr;
  void* cub_tmp = nullptr; size_t cub_bytes = 0;
  int key_bits = 0;
};

}  // namespace

struct npt_ctx {
  npt_config cfg{};
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  std::string err;
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_aux = nullptr;   // s_aux: the serial walker's reads beside the bulk
  bool in_chrom = false, finalized = false;
  // chromosome state
  int chrom_index = 0;
  int64_t G = 0;
  int n_orf = 0;
  DevCfg dc{};
  bool ref_smem = false;
  DBuf<uint32_t> d_ref;
  DBuf<uint8_t> d_ref_codes;
  DBuf<int> d_ostart, d_oname;
  DBuf<uint8_t> d_ostrand;
  DBuf<int4> d_gx; DBuf<unsigned> d_gxhash;
  DBuf<Slot> d_cl;
  DBuf<int> d_cl_rec;
  DBuf<Slot> d_sb;
  DBuf<int> d_sb_rec, d_sub_count, d_sub_break_sum;
  DBuf<unsigned long long> d_sub_sums, d_sub_min0;
  int walk_blocks[3] = {0, 0, 0};   // grid sizes of walk_kernel<0/1/2>
  size_t walk_smem = 0, cluster_smem = 0;
  int cluster_occ = 8, covkey_occ = 6;   // resident cluster_kernel / cov_key_kernel blocks per SM
  // fast walk (fastwalk.cuh / fastcov.cuh)
  DBuf<uint32_t> d_refg, d_fw_scratch;
  bool fast = false, fast_record = false;
  int fw_blocks = 0, fw_threads = 0, acc_blocks = 0, acc_pos_shift = 0;
  size_t fw_smem = 0;
  DBuf<PairEntry> d_pr;
  DBuf<int> d_cov, d_cov_out;
  DBuf<int> d_small;
  DBuf<unsigned> d_counters;
  // finalize scratch
  DBuf<unsigned long long> d_keys, d_keys2;
  DBuf<int> d_vals, d_vals2, d_cl_slot, d_sb_slot, d_cl_rank, d_cl_slot_rank, d_sb_grank, d_cl_sub_off;
  DBuf<unsigned char> d_cub;
  DBuf<long long> d_break_off;
  DBuf<int> d_breaks_out;
  // device block pool: cudaFree of large blocks is synchronous and slow (hundreds of ms per chromosome in the e2e trace),
  // so per-batch output blocks are recycled across batches and chromosomes and only freed in npt_destroy
  std::vector<std::pair<void*, size_t>> pool_free;
  // batches
  std::vector<Batch> batches;
  Staging stg[2];
  int64_t idx_base = 0;       // global index of the next read
  int64_t idx_user_base = 0;  // npt_set_read_index_base
  // finalize products
  unsigned h_counters[CN_COUNT]{};
  int n_clusters = 0, n_subs = 0;
  int cov_rows = 0;           // rows of d_cov_out (n_clusters)
  int64_t n_pairs = 0;
  int64_t n_total = 0;
  int64_t n_slow = 0;            // reads the fast walk handed to the serial walker (this chromosome)
  bool split_streams = true;
  float slow_share = 0;          // reads of the last retired batch that went to the serial walker, as a share of the batch
  bool cov_dirty = true;         // d_cov may hold counts (fresh allocation, or a chromosome that was not finished)
  float brk_per_read = 0;        // largest overflow need (ints per read) of a retired batch of this chromosome
  long long total_breaks = 0;
  DBuf<long long> d_ex_first;
  DBuf<unsigned> d_ex_u[5];
  DBuf<int> d_ex_i[2];
  DBuf<unsigned char> d_ex_flags;
  // end-of-chromosome columns in final (rank) order that a multi-rank caller sums in place before fetching
  DBuf<long long> d_exc_first; DBuf<unsigned> d_exc_u[2];  // cluster columns (fetch)
  DBuf<unsigned> d_fin_sum_off;              // [n_subs + 1]
  DBuf<long long> d_fin_sums;                // [total_sums]
  long long total_sums = 0;
  DBuf<unsigned long long> d_pair_keys, d_pair_keys2;  // compacted / sorted break-point pair keys
  DBuf<int> d_pair_cnt, d_pair_cnt2;
  // multi-rank key exchange
  DBuf<int> d_kx, d_kx_slotmap;
  DBuf<unsigned> d_kx_ctr;
  int kx_hdr[16] = {0};          // header of the last npt_export_keys (KX_HDR ints; [13] = words of the buffer)
  struct NptComm* comm = nullptr;   // collectives of this rank (comm.inl); null = single rank
  int comm_rank = 0, comm_world = 1;
  DBuf<int> d_kx_all, d_kx_hdrs;
  float kx_ms = 0, reduce_ms = 0;
  npt_results res{};
  // host results (pinned for the large per-read arrays)
  HBuf<int> r_cluster, r_sub, r_start, r_end, r_rst, r_ren, r_maps, r_err, r_breaks;
  HBuf<unsigned> r_flags;
  HBuf<unsigned long long> r_boff;
  HBuf<int> r_cov;
  std::vector<int64_t> t_cl_first, t_sub_first, t_sub_sums;
  std::vector<uint32_t> t_cl_key_off, t_cl_sub_off, t_sub_key_off, t_sub_sum_off;
  std::vector<int32_t> t_cl_key_tok, t_cl_read_count, t_sub_key, t_sub_count, t_sub_break_sum, t_small;
  std::vector<uint8_t> t_sub_flags;
  std::vector<int32_t> t_pair[5];
  // timing
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
  float step_ms = 0;
  float walk_ms = 0, fin_ms = 0;
  int64_t walk_launches = 0, total_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> walk_events;
  std::vector<cudaEvent_t> stage_events;  // 12 per batch: (begin, end) of fast walk, walk<0>, walk<2> (+ clustering of its reads), cluster, walk<1>, coverage (keys + sort + accumulate)
  std::vector<cudaEvent_t> order_events;  // fork / join of the second compute stream, 3 per batch
  float stage_ms[6] = {0, 0, 0, 0, 0, 0};
};

static void comm_release(npt_ctx* ctx);   // comm.inl

namespace {

int fail(npt_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}
#define CK(call)                                                                                           \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? NPT_ENOMEM : NPT_ECUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                            \
  } while (0)

unsigned pow2_at_least(unsigned long long v) { unsigned p = 1; while (p < v && p < (1u << 31)) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void init_tables_kernel(Slot* cl, unsigned ncl, Slot* sb, unsigned nsb, PairEntry* pr, unsigned npr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < ncl; k += stride) { cl[k].word = 0; cl[k].min_idx = ~0ULL; }
  for (size_t k = i; k < nsb; k += stride) { sb[k].word = 0; sb[k].min_idx = ~0ULL; }
  for (size_t k = i; k < npr; k += stride) { pr[k].key = 0; pr[k].count = 0; pr[k].pad = 0; }
}

// reference codes (one byte per base) -> packed words: ref[w] = bases 8w..8w+7 (0-based), grid[W] = positions 8W..8W+7 (1-based),
// first in the top nibble, zero beyond the reference
__global__ void pack_reference_kernel(const uint8_t* __restrict__ codes, long long n, uint32_t* ref, int nwords, uint32_t* grid, int gwords) {
  const int total = nwords > gwords ? nwords : gwords;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    uint32_t a = 0, g = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const long long i = 8ll * w + k;          // 0-based base index = position i + 1
      if (i < n) a |= (uint32_t)(codes[i] & 0xF) << (28 - 4 * k);
      const long long j = i - 1;                // grid word w, nibble k = position 8w + k = base index 8w + k - 1
      if (j >= 0 && j < n) g |= (uint32_t)(codes[j] & 0xF) << (28 - 4 * k);
    }
    if (w < nwords) ref[w] = a;
    if (w < gwords) grid[w] = g;
  }
}

// cluster table slots -> dense arrays indexed by the cluster's dense id
__global__ void collect_clusters_kernel(const Slot* cl, unsigned nslots, const int* cl_rec, int n, unsigned long long* keys, int* vals,
                                        int* slot_of) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long w = cl[s].word;
    if (w) {
      const int p = cl_rec[(unsigned)(w & 0xffffffffu)];
      if (p >= 0 && p < n) { keys[p] = cl[s].min_idx; vals[p] = p; slot_of[p] = (int)s; }
    }
  }
}
// sub-cluster slots -> (sort key, slot) pairs in arbitrary dense order; key = (cluster rank << 40) | first read index
__global__ void collect_subs_kernel(const Slot* sb, unsigned nslots, const int* sb_rec, const Slot* cl, const int* cl_rec, const int* cl_rank,
                                    int n, unsigned long long* keys, int* vals, unsigned* counter) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long w = sb[s].word;
    if (w) {
      const int cslot = sb_rec[(unsigned)(w & 0xffffffffu)];
      const int cprov = cl_rec[(unsigned)(cl[cslot].word & 0xffffffffu)];
      const unsigned o = atomicAdd(counter, 1u);
      if ((int)o < n) { keys[o] = ((unsigned long long)(unsigned)cl_rank[cprov] << 40) | ((sb[s].min_idx >> 2) & ((1ULL << 40) - 1)); vals[o] = (int)s; }
    }
  }
}
// rank_of_slot[slot] = rank for clusters (via dense id) / global sub rank for sub-clusters
__global__ void cluster_slot_ranks_kernel(const int* slot_of_prov, const int* rank_of_prov, int n, int* rank_of_slot) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) rank_of_slot[slot_of_prov[p]] = rank_of_prov[p];
}
__global__ void ranks_kernel(const int* sorted_vals, int n, int* rank_of) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) rank_of[sorted_vals[r]] = r;
}
__global__ void sub_offsets_kernel(const unsigned long long* sorted_keys, int n, int n_clusters, int* cl_sub_off) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int c = (int)(sorted_keys[r] >> 40);
    if (r == 0 || (int)(sorted_keys[r - 1] >> 40) != c) cl_sub_off[c] = r;
    if (r == n - 1) cl_sub_off[n_clusters] = n;
  }
}
__global__ void remap_reads_kernel(int64_t n, int* cluster, int* sub, const int* cl_rank, const int* sb_grank, const int* cl_sub_off) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int cp = cluster[r], sp = sub[r];  // table slots
    int cr = -1, sr = -1;
    if (cp >= 0) { cr = cl_rank[cp]; if (sp >= 0) sr = sb_grank[sp] - cl_sub_off[cr]; }
    cluster[r] = cr; sub[r] = sr;
  }
}
__global__ void gather_breaks_kernel(int64_t n, const int* nbreaks, const unsigned* break_loc, const int* arena, long long* off,
                                     long long off_base, int* out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int nb = nbreaks[r];
    const unsigned loc = break_loc[r];
    const int* src = (loc == 0xFFFFFFFFu) ? arena + r * BRK_INLINE : arena + n * BRK_INLINE + loc;
    const long long o = off_base + off[r];
    off[r] = o;  // batch-relative -> absolute
    int* dst = out + o;
    for (int i = 0; i < nb; i++) dst[i] = src[i];
  }
}
// dense, rank-ordered copies of the table columns the host needs
__global__ void export_clusters_kernel(const Slot* cl, const int* cl_rec, const int* slot_of, const int* rank_of, int n, long long* first,
                                       unsigned* tok_off, unsigned* ntok) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const Slot& e = cl[slot_of[p]];
    const unsigned off = (unsigned)(e.word & 0xffffffffu);
    const int r = rank_of[p];
    first[r] = (long long)e.min_idx; tok_off[r] = off + CL_REC_HDR; ntok[r] = (unsigned)cl_rec[off + 1];
  }
}
// sorted_slots[g] = slot of the sub-cluster with global rank g
__global__ void export_subs_kernel(const Slot* sb, const int* sb_rec, const int* sorted_slots, int n, int S, const int* sub_count,
                                   const int* sub_break_sum, long long* first, unsigned* key_off, unsigned* nkey, int* count, int* break_sum,
                                   unsigned* sum_off, unsigned* nsum, unsigned char* flags) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
    const int s = sorted_slots[g];
    const Slot& e = sb[s];
    const unsigned off = (unsigned)(e.word & 0xffffffffu);
    first[g] = (long long)(e.min_idx >> 2); flags[g] = (unsigned char)(e.min_idx & 3);
    key_off[g] = off + SB_REC_HDR; nkey[g] = (unsigned)sb_rec[off + 1];
    sum_off[g] = (unsigned)sb_rec[off + 2]; nsum[g] = (unsigned)sb_rec[off + 3];
    for (int k = 0; k < S; k++) count[(size_t)g * S + k] = sub_count[(size_t)s * S + k];
    break_sum[g] = sub_break_sum[s];
  }
}

__global__ void compact_pairs_kernel(const PairEntry* pr, unsigned nslots, unsigned long long* keys, int* counts, unsigned* counter, int cap) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long k = pr[s].key;
    if (k) {
      const unsigned o = atomicAdd(counter, 1u);
      if ((int)o < cap) { keys[o] = k; counts[o] = pr[s].count; }
    }
  }
}
// break sums of the sub-clusters, from arena order to final (rank) order
__global__ void gather_sums_kernel(int n, const unsigned* arena_off, const unsigned* nsum, const unsigned* fin_off, const unsigned long long* arena,
                                   long long* out) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
    for (unsigned i = 0; i < nsum[g]; i++) out[fin_off[g] + i] = (long long)arena[arena_off[g] + i];
}
// multi-rank key exchange: live slots of the three tables as lists (order irrelevant)
__global__ void kx_list_clusters_kernel(const Slot* cl, unsigned nslots, int* list, unsigned* counter, int cap) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long w = cl[s].word;
    if (w) {
      const unsigned o = atomicAdd(counter, 1u);
      if ((int)o < cap) { list[4 * o] = (int)s; list[4 * o + 1] = (int)(w & 0xffffffffu); list[4 * o + 2] = (int)(cl[s].min_idx & 0xffffffffu); list[4 * o + 3] = (int)(cl[s].min_idx >> 32); }
    }
  }
}
__global__ void kx_list_subs_kernel(const Slot* sb, unsigned nslots, const unsigned long long* min0, int* list, unsigned* counter, int cap) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long w = sb[s].word;
    if (w) {
      const unsigned o = atomicAdd(counter, 1u);
      if ((int)o < cap) {
        const unsigned long long m = sb[s].min_idx, m0 = min0 ? min0[s] : ~0ULL;
        list[6 * o] = (int)(w & 0xffffffffu); list[6 * o + 1] = (int)(m & 0xffffffffu); list[6 * o + 2] = (int)(m >> 32);
        list[6 * o + 3] = (int)(m0 & 0xffffffffu); list[6 * o + 4] = (int)(m0 >> 32); list[6 * o + 5] = 0;
      }
    }
  }
}
__global__ void kx_list_pairs_kernel(const PairEntry* pr, unsigned nslots, int* keys, unsigned* counter, int cap) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
    const unsigned long long k = pr[s].key;
    if (k) {
      const unsigned o = atomicAdd(counter, 1u);
      if ((int)o < cap) { keys[2 * o] = (int)(k & 0xffffffffu); keys[2 * o + 1] = (int)(k >> 32); }
    }
  }
}

// the first `used` elements of each of the COV_ARRAYS accumulation arrays back to zero (blockIdx.y = array)
__global__ void clear_rows_kernel(int* cov, size_t arr_stride, size_t used) {
  int* a = cov + (size_t)blockIdx.y * arr_stride;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < used; i += (size_t)gridDim.x * blockDim.x * 4)
    for (int k = 0; k < 4; k++) if (i + k < used) a[i + k] = 0;
}

// Coverage rows: out[arr][row_of(prov)][src][:] = arr == 0 ? inclusive scan of the difference row : copy.
// One block per (arr, prov, src) row; the "segmented prefix scan" of the design (segments = rows).
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8;
__global__ void __launch_bounds__(SCAN_THREADS) coverage_rows_kernel(const int* cov, int cov_cap, int n_prov, int S, int stride,
                                                                     const int* row_of_prov, int n_rows, int* out) {
  const int row = blockIdx.x;  // over arr * n_prov * S
  const int src = row % S, prov = (row / S) % n_prov, arr = row / (S * n_prov);
  const int orow = row_of_prov[prov];
  if (orow < 0) return;
  const size_t arr_stride = (size_t)cov_cap * S * stride;
  const size_t rowoff = ((size_t)prov * S + src) * stride;
  const int* in = cov + arr * arr_stride + rowoff;
  int* o = out + (((size_t)arr * n_rows + orow) * S + src) * stride;
  // deletion histogram rows (arrays 4..): dh[l-1][p] = deletions of length l whose first deleted position is p.  The walk
  // carried the depth run across them, so: remove the deleted positions [p, p+l-1], add the positions the reference emits
  // after advancing, [p+l, p+2l-1], to depth and to errors (the D quirk, IdentityProfile1.java:513-522).
  auto dh = [&](int l, int p) -> int { return (p >= 0 && p < stride) ? cov[(size_t)(3 + l) * arr_stride + rowoff + p] : 0; };
  if (arr == 1) {
    for (int i = threadIdx.x; i < stride; i += SCAN_THREADS) {
      int v = in[i];
      for (int l = 1; l <= DEL_HIST; l++) for (int j = 0; j < l; j++) v += dh(l, i - l - j);
      o[i] = v;
    }
    return;
  }
  if (arr != 0) {
    for (int i = threadIdx.x; i < stride; i += SCAN_THREADS) o[i] = in[i];
    return;
  }
  __shared__ int warp_tot[SCAN_THREADS / 32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < stride; base += SCAN_THREADS * SCAN_ITEMS) {
    int v[SCAN_ITEMS], absd[SCAN_ITEMS];   // absd: depth credited as absolute counts by accumulate_kernel (array 8)
    int sum = 0;
    const int i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
      int x = 0;
      absd[j] = 0;
      if (i0 + j < stride) {
        x = in[i0 + j];
        absd[j] = cov[8 * arr_stride + rowoff + i0 + j];
        for (int l = 1; l <= DEL_HIST; l++) x += -dh(l, i0 + j) + 2 * dh(l, i0 + j - l) - dh(l, i0 + j - 2 * l);
      }
      v[j] = x; sum += v[j]; v[j] = sum;
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (unsigned w2 = 0; w2 < warp; w2++) woff += warp_tot[w2];
    const int carry = carry_s;
    const int excl = carry + woff + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) if (i0 + j < stride) o[i0 + j] = excl + v[j] + absd[j];
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl + sum;
    __syncthreads();
  }
}

// The per-batch kernel sequence on the compute stream.  from_stage: 0 = everything, 1 = from the store-all pass on
// (used when the overflow area of a batch had to be enlarged).
int launch_batch(npt_ctx* ctx, const Batch& B, int from_stage) {
  const BatchDev& bd = B.d;
  const DevCfg& dc = ctx->dc;
  const int T = WALK_THREADS;
  cudaStream_t sc = ctx->s_comp, sa = ctx->s_aux;
  // 12 events per batch: (begin, end) of the six stages of npt_last_stage_ms
  cudaEvent_t ev[12];
  for (auto& e : ev) { CK(cudaEventCreate(&e)); ctx->stage_events.push_back(e); }
  cudaEvent_t fork_ev, join_cl, join_cov;   // ordering only
  CK(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&join_cl, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&join_cov, cudaEventDisableTiming));
  ctx->order_events.push_back(fork_ev); ctx->order_events.push_back(join_cl); ctx->order_events.push_back(join_cov);
  // exactly two waves of resident blocks (grid-stride loop inside): a grid that is not a multiple of what fits the device ends
  // in a part-filled last wave (2 368 blocks at 7 instead of 8 blocks per SM cost 1.5x, r02i); one wave alone is 15 % slower
  // than two (5.4 against 4.7 ms per 10 M reads: the second wave evens out the per-read differences)
  const int cgrid = (int)std::max<long long>(1, std::min<long long>((bd.n + CLUSTER_THREADS - 1) / CLUSTER_THREADS, 2ll * ctx->sm_count * ctx->cluster_occ));
  // ---- the fast walk: every read it can restate exactly (breaks, read offsets, coverage entries); the others go to slow_list
  CK(cudaEventRecord(ev[0], sc));
  // The serial walker's few reads (walk<0>, walk<2>, their clustering, walk<1>) run beside the bulk on a second stream: each
  // is one long dependent chain in a handful of warps (0.65 + 0.9 ms per batch whatever its size), hidden behind
  // cluster_kernel and the coverage stage (-1.1 ms per 10 M reads; NPT_SPLIT=0 puts everything on one stream).
  const bool split = bd.fast == 2 && from_stage == 0;
  if (bd.fast && from_stage == 0) {
    const bool rec = ctx->fast_record;
    if (ctx->ref_smem) {
      if (rec) fast_walk_kernel<true, true><<<ctx->fw_blocks, ctx->fw_threads, ctx->fw_smem, sc>>>(dc, bd);
      else fast_walk_kernel<true, false><<<ctx->fw_blocks, ctx->fw_threads, ctx->fw_smem, sc>>>(dc, bd);
    } else {
      if (rec) fast_walk_kernel<false, true><<<ctx->fw_blocks, ctx->fw_threads, ctx->fw_smem, sc>>>(dc, bd);
      else fast_walk_kernel<false, false><<<ctx->fw_blocks, ctx->fw_threads, ctx->fw_smem, sc>>>(dc, bd);
    }
    ctx->total_launches++;
  }
  CK(cudaEventRecord(ev[1], sc));
  cudaStream_t sw = split ? sa : sc;   // where the serial walker runs
  if (split) { CK(cudaEventRecord(fork_ev, sc)); CK(cudaStreamWaitEvent(sa, fork_ev, 0)); }
  // ---- the exact serial walker (all reads when the fast walk is off, else the reads of slow_list)
  CK(cudaEventRecord(ev[2], sw));
  if (from_stage == 0) {
    if (ctx->ref_smem) walk_kernel<0, true><<<ctx->walk_blocks[0], T, ctx->walk_smem, sw>>>(dc, bd);
    else walk_kernel<0, false><<<ctx->walk_blocks[0], T, ctx->walk_smem, sw>>>(dc, bd);
  }
  CK(cudaEventRecord(ev[3], sw));
  CK(cudaEventRecord(ev[4], sw));
  if (ctx->ref_smem) walk_kernel<2, true><<<ctx->walk_blocks[2], T, ctx->walk_smem, sw>>>(dc, bd);
  else walk_kernel<2, false><<<ctx->walk_blocks[2], T, ctx->walk_smem, sw>>>(dc, bd);
  if (split) cluster_kernel<2><<<std::min(cgrid, ctx->sm_count), CLUSTER_THREADS, ctx->cluster_smem, sw>>>(dc, bd);
  CK(cudaEventRecord(ev[5], sw));
  if (split) CK(cudaEventRecord(join_cl, sa));
  // ---- clustering of the bulk
  CK(cudaEventRecord(ev[6], sc));
  if (split) cluster_kernel<1><<<cgrid, CLUSTER_THREADS, ctx->cluster_smem, sc>>>(dc, bd);
  else cluster_kernel<0><<<cgrid, CLUSTER_THREADS, ctx->cluster_smem, sc>>>(dc, bd);
  CK(cudaEventRecord(ev[7], sc));
  ctx->total_launches += (from_stage == 0 ? 3 : 2) + (split ? 1 : 0);
  // ---- coverage of the serial walker's reads (needs their clusters: after cluster_kernel<2> on its own stream)
  CK(cudaEventRecord(ev[8], sw));
  if (dc.record_depth) {
    if (ctx->ref_smem) walk_kernel<1, true><<<ctx->walk_blocks[1], T, ctx->walk_smem, sw>>>(dc, bd);
    else walk_kernel<1, false><<<ctx->walk_blocks[1], T, ctx->walk_smem, sw>>>(dc, bd);
    ctx->total_launches++;
  }
  CK(cudaEventRecord(ev[9], sw));
  if (split) CK(cudaEventRecord(join_cov, sa));
  // ---- coverage of the fast reads: key = (coverage row, source) -> radix sort -> positional popcount per group
  CK(cudaEventRecord(ev[10], sc));
  if (dc.record_depth && bd.fast) {
    const int kgrid = (int)std::max<long long>(1, std::min<long long>((bd.n + 255) / 256, (long long)ctx->sm_count * ctx->covkey_occ));
    cov_key_kernel<<<kgrid, 256, 0, sc>>>(dc, bd, B.k_in, B.v_in, ctx->acc_pos_shift);
    size_t tmp = B.cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(B.cub_tmp, tmp, B.k_in, B.k_out, B.v_in, B.v_out, (int)bd.n, 0, B.key_bits, sc));
    accumulate_kernel<<<ctx->acc_blocks, ACC_THREADS, sizeof(AccSmem), sc>>>(dc, bd, B.k_out, B.v_out, bd.bctr + 6);
    ctx->total_launches += 2 + (B.key_bits + 7) / 8 + 1;
  }
  CK(cudaEventRecord(ev[11], sc));
  if (split) { CK(cudaStreamWaitEvent(sc, join_cl, 0)); CK(cudaStreamWaitEvent(sc, join_cov, 0)); }
  CK(cudaGetLastError());
  return NPT_OK;
}

cudaError_t pool_get(npt_ctx* ctx, void** out, size_t bytes, size_t* got);
void pool_put(npt_ctx* ctx, void* p, size_t bytes);

// the fast walk's per-batch scratch is dead once the batch's kernels have run
void release_scratch(npt_ctx* ctx, Batch& B) {
  if (B.f_hdr) { pool_put(ctx, B.f_hdr, B.f_hdr_bytes); B.f_hdr = nullptr; }
  if (B.f_ent) { pool_put(ctx, B.f_ent, B.f_ent_bytes); B.f_ent = nullptr; }
  if (B.f_sort) { pool_put(ctx, B.f_sort, B.f_sort_bytes); B.f_sort = nullptr; }
}

int retire_batch(npt_ctx* ctx, size_t bi) {
  Batch& B = ctx->batches[bi];
  if (B.retired) return NPT_OK;
  CK(cudaEventSynchronize(B.done));
  unsigned h[12];
  CK(cudaMemcpy(h, B.bctr, sizeof h, cudaMemcpyDeviceToHost));
  if (!h[1] && h[8] && B.d.fast) {
    // Only the serial walker's reads ran out of overflow area; the bulk (fast walk, its clustering and coverage) is done.
    // The area keeps what the fast walk stored and grows by what the chain can need at most (h[0] counts every request so
    // far, failed ones included); then the chain runs again: store-all pass, clustering and coverage of slow_list.
    const DevCfg& dc = ctx->dc;
    const unsigned old_cap = B.d.break_cap, new_cap = (unsigned)std::min<unsigned long long>(2ull * h[0] + 16, 1ull << 31);
    int* bigger = nullptr;
    size_t bigger_bytes = 0;
    CK(pool_get(ctx, (void**)&bigger, ((size_t)B.n * BRK_INLINE + new_cap) * sizeof(int), &bigger_bytes));
    CK(cudaMemcpyAsync(bigger, B.d.break_arena, ((size_t)B.n * BRK_INLINE + old_cap) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->s_comp));
    CK(cudaMemsetAsync(B.bctr + 3, 0, 2 * sizeof(unsigned), ctx->s_comp));   // read cursors of walk_kernel<1> and <2>
    CK(cudaMemsetAsync(B.bctr + 8, 0, sizeof(unsigned), ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_comp));
    pool_put(ctx, B.d.break_arena, B.arena_bytes);
    B.d.break_arena = bigger; B.d.break_cap = new_cap; B.arena_bytes = bigger_bytes;
    const BatchDev& bd = B.d;
    if (ctx->ref_smem) walk_kernel<2, true><<<ctx->walk_blocks[2], WALK_THREADS, ctx->walk_smem, ctx->s_comp>>>(dc, bd);
    else walk_kernel<2, false><<<ctx->walk_blocks[2], WALK_THREADS, ctx->walk_smem, ctx->s_comp>>>(dc, bd);
    cluster_kernel<2><<<ctx->sm_count, CLUSTER_THREADS, ctx->cluster_smem, ctx->s_comp>>>(dc, bd);
    if (dc.record_depth) {
      if (ctx->ref_smem) walk_kernel<1, true><<<ctx->walk_blocks[1], WALK_THREADS, ctx->walk_smem, ctx->s_comp>>>(dc, bd);
      else walk_kernel<1, false><<<ctx->walk_blocks[1], WALK_THREADS, ctx->walk_smem, ctx->s_comp>>>(dc, bd);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->s_comp));
    ctx->total_launches += 3;
  }
  if (h[1]) {
    // The overflow area for reads with more than BRK_INLINE breaks was too small; cluster_kernel and the coverage pass
    // skipped this batch.  h[0] is the exact need: enlarge and run the batch from the store-all pass on.  Running it
    // late is fine: the tables are order-independent (first-appearance = atomicMin of the global read index).
    const unsigned need = h[0];
    int* bigger = nullptr;
    size_t bigger_bytes = 0;
    CK(pool_get(ctx, (void**)&bigger, ((size_t)B.n * BRK_INLINE + need) * sizeof(int), &bigger_bytes));
    CK(cudaMemcpyAsync(bigger, B.d.break_arena, (size_t)B.n * BRK_INLINE * sizeof(int), cudaMemcpyDeviceToDevice, ctx->s_comp));
    CK(cudaMemsetAsync(B.bctr, 0, 7 * sizeof(unsigned), ctx->s_comp));  // [7] = reads in slow_list stays
    CK(cudaMemsetAsync(B.bctr + 8, 0, sizeof(unsigned), ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_comp));
    pool_put(ctx, B.d.break_arena, B.arena_bytes);
    B.d.break_arena = bigger; B.d.break_cap = need; B.arena_bytes = bigger_bytes;
    // the serial store-all pass now redoes every long list, also those of reads the fast walk had finished (their overflow
    // slots were handed out from the counter that has just been reset)
    if (B.d.fast) { ctx->n_slow += h[7]; B.d.fast = 0; }
    int rc = launch_batch(ctx, B, 1);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->s_comp));
  }
  release_scratch(ctx, B);
  if (B.staging >= 0) { ctx->stg[B.staging].busy = false; ctx->stg[B.staging].batch_index = -1; }
  if (B.d.fast) { ctx->n_slow += h[7]; ctx->slow_share = (float)((double)h[7] / (double)std::max<int64_t>(B.n, 1)); }
  // overflow ints per read this batch needed: the next batches of the chromosome reserve that much (plus a quarter)
  ctx->brk_per_read = std::max(ctx->brk_per_read, (float)((double)h[0] / (double)std::max<int64_t>(B.n, 1)));
  B.retired = true;
  return NPT_OK;
}

cudaError_t pool_get(npt_ctx* ctx, void** out, size_t bytes, size_t* got) {
  int best = -1;
  for (size_t i = 0; i < ctx->pool_free.size(); i++)
    if (ctx->pool_free[i].second >= bytes && (best < 0 || ctx->pool_free[i].second < ctx->pool_free[best].second)) best = (int)i;
  if (best >= 0 && ctx->pool_free[best].second <= 2 * bytes + (1 << 20)) {
    *out = ctx->pool_free[best].first; *got = ctx->pool_free[best].second;
    ctx->pool_free.erase(ctx->pool_free.begin() + best);
    return cudaSuccess;
  }
  *got = bytes;
  return cudaMalloc(out, std::max<size_t>(bytes, 256));
}
void pool_put(npt_ctx* ctx, void* p, size_t bytes) { if (p) ctx->pool_free.emplace_back(p, bytes); }

void free_batches(npt_ctx* ctx) {
  for (auto& B : ctx->batches) {
    pool_put(ctx, B.out_block, B.out_bytes);
    pool_put(ctx, B.d.break_arena, B.arena_bytes);
    pool_put(ctx, B.bctr, B.bctr_bytes);
    release_scratch(ctx, B);
    if (B.f_slow) pool_put(ctx, B.f_slow, B.f_slow_bytes);
    if (B.done) cudaEventDestroy(B.done);
  }
  ctx->batches.clear();
  for (auto& s : ctx->stg) { s.busy = false; s.batch_index = -1; }
  for (auto& e : ctx->walk_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  ctx->walk_events.clear();
  for (auto& e : ctx->stage_events) cudaEventDestroy(e);
  ctx->stage_events.clear();
  for (auto& e : ctx->order_events) cudaEventDestroy(e);
  ctx->order_events.clear();
}

}  // namespace

// ================================================================================================= C ABI
extern "C" {

int npt_abi_version(void) { return NPT_ABI_VERSION; }

const char* npt_last_error(const npt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int npt_create(const npt_config* cfg, npt_ctx** out) {
  npt_ctx* ctx = nullptr;
  if (!cfg || !out) return fail(nullptr, NPT_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->abi_version != NPT_ABI_VERSION) return fail(nullptr, NPT_EINVAL, "abi_version %d != %d", cfg->abi_version, NPT_ABI_VERSION);
  if (cfg->num_sources < 1 || cfg->num_sources > NPT_MAX_SOURCES) return fail(nullptr, NPT_EINVAL, "num_sources out of range");
  if (cfg->bin < 1) return fail(nullptr, NPT_EINVAL, "bin must be >= 1");
  if (cfg->break_thresh < 0) return fail(nullptr, NPT_EINVAL, "break_thresh must be >= 0");
  if (cfg->max_clusters > (1 << 24)) return fail(nullptr, NPT_EINVAL, "max_clusters is limited to 2^24 (cluster rank and first index share one 64-bit sort key)");
  if (cfg->coronavirus)
    for (int s = 0; s < cfg->num_sources; s++)
      if (!cfg->rna[s])
        return fail(nullptr, NPT_EINVAL, "coronavirus mode needs rna[src]=1 (the reference unboxes a null Boolean, IdentityProfile1.java:165-170,268)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, NPT_ENODEVICE, "no CUDA device (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, NPT_ENODEVICE, "device %d of %d", cfg->device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, NPT_ECUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, NPT_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
  ctx = new npt_ctx();
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  auto init = [&]() -> int {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ctx->ev_a));
    CK(cudaEventCreate(&ctx->ev_b));
    CK(cudaEventCreate(&ctx->ev_c0));
    CK(cudaEventCreate(&ctx->ev_c1));
    for (auto& s : ctx->stg) CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    return NPT_OK;
  };
  const int rc = init();
  if (rc) {   // nothing of a half-built context survives; the message moves to where npt_last_error(NULL) finds it
    g_create_error = ctx->err;
    npt_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return NPT_OK;
}

void npt_destroy(npt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  comm_release(ctx);
  free_batches(ctx);
  for (auto& pb : ctx->pool_free) cudaFree(pb.first);
  ctx->pool_free.clear();
  ctx->d_ref.release(); ctx->d_ref_codes.release(); ctx->d_refg.release(); ctx->d_fw_scratch.release(); ctx->d_ostart.release(); ctx->d_oname.release(); ctx->d_exc_first.release(); ctx->d_exc_u[0].release(); ctx->d_exc_u[1].release(); ctx->d_fin_sum_off.release(); ctx->d_fin_sums.release(); ctx->d_pair_keys.release(); ctx->d_pair_keys2.release(); ctx->d_pair_cnt.release(); ctx->d_pair_cnt2.release(); ctx->d_kx.release(); ctx->d_kx_slotmap.release(); ctx->d_kx_ctr.release();
  ctx->d_ostrand.release(); ctx->d_gx.release(); ctx->d_gxhash.release(); ctx->d_cl.release();
  ctx->d_cl_rec.release(); ctx->d_sb.release(); ctx->d_sb_rec.release(); ctx->d_sub_count.release(); ctx->d_sub_break_sum.release();
  ctx->d_sub_sums.release(); ctx->d_sub_min0.release(); ctx->d_pr.release();
  ctx->d_cov.release(); ctx->d_cov_out.release(); ctx->d_small.release(); ctx->d_counters.release(); ctx->d_keys.release();
  ctx->d_keys2.release(); ctx->d_vals.release(); ctx->d_vals2.release(); ctx->d_cl_slot.release(); ctx->d_sb_slot.release();
  ctx->d_cl_rank.release(); ctx->d_cl_slot_rank.release(); ctx->d_sb_grank.release(); ctx->d_cl_sub_off.release(); ctx->d_cub.release(); ctx->d_break_off.release();
  ctx->d_breaks_out.release(); ctx->d_ex_first.release(); ctx->d_ex_flags.release();
  for (auto& b : ctx->d_ex_u) b.release();
  for (auto& b : ctx->d_ex_i) b.release();
  for (auto& s : ctx->stg) { s.buf.release(); if (s.done) cudaEventDestroy(s.done); }
  ctx->r_cluster.release(); ctx->r_sub.release(); ctx->r_start.release(); ctx->r_end.release(); ctx->r_rst.release(); ctx->r_ren.release();
  ctx->r_maps.release(); ctx->r_err.release(); ctx->r_breaks.release(); ctx->r_flags.release(); ctx->r_boff.release(); ctx->r_cov.release();
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_c0) cudaEventDestroy(ctx->ev_c0);
  if (ctx->ev_c1) cudaEventDestroy(ctx->ev_c1);
  if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
  if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
  if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
  delete ctx;
}

int npt_set_read_index_base(npt_ctx* ctx, int64_t base) {
  if (!ctx) return NPT_EINVAL;
  if (base < 0 || base >= (1LL << 40)) return fail(ctx, NPT_EINVAL, "read index base out of range");
  if (ctx->in_chrom && !ctx->batches.empty()) return fail(ctx, NPT_ESTATE, "npt_set_read_index_base after the first npt_submit of a chromosome");
  ctx->idx_user_base = base;
  ctx->idx_base = base;  // also valid when called between npt_begin_chrom and the first npt_submit
  return NPT_OK;
}

int npt_set_next_read_index(npt_ctx* ctx, int64_t index) {
  if (!ctx) return NPT_EINVAL;
  if (index < 0 || index >= (1LL << 40)) return fail(ctx, NPT_EINVAL, "read index out of range");
  ctx->idx_base = index;
  return NPT_OK;
}

int npt_begin_chrom(npt_ctx* ctx, int32_t chrom_index, const uint8_t* ref_codes, int64_t ref_len, const npt_annotation* annot) {
  if (!ctx || !ref_codes || !annot || ref_len < 1) return fail(ctx, NPT_EINVAL, "bad argument");
  const npt_config& cf = ctx->cfg;
  if (ref_len > (1LL << 31) - 64) return fail(ctx, NPT_EINVAL, "reference too long");
  if (annot->kind != NPT_ANNOT_CSV && annot->kind != NPT_ANNOT_EMPTY && annot->kind != NPT_ANNOT_GFF) return fail(ctx, NPT_EINVAL, "unknown annotation kind");
  const bool gff = annot->kind == NPT_ANNOT_GFF;
  if (gff && cf.coronavirus) return fail(ctx, NPT_EINVAL, "a GFF exon table is a host-mode annotation (coronavirus=0); virus mode takes the CSV ORF table");
  if (gff && annot->n > 0 && (!annot->start || !annot->end || !annot->strand || !annot->name_id || !annot->name_hash))
    return fail(ctx, NPT_EINVAL, "GFF annotation needs start, end, strand, name_id and name_hash");
  if (annot->kind == NPT_ANNOT_CSV && cf.coronavirus && annot->n < 2)
    return fail(ctx, NPT_EINVAL, "coronavirus mode needs >= 2 annotation rows (Annotation.isLeader reads start.get(1), Annotation.java:214-216)");
  if (annot->n < 0 || annot->n > (gff ? (1 << 22) : 4096)) return fail(ctx, NPT_EINVAL, "annotation rows out of range");
  if (!cf.coronavirus && annot->kind == NPT_ANNOT_EMPTY)
    for (int s = 0; s < cf.num_sources; s++)
      if (!cf.rna[s])
        return fail(ctx, NPT_EINVAL, "host mode with EmptyAnnotation needs rna[src]=1: the reference passes a null Boolean to "
                    "EmptyAnnotation.nextUpstream(int,int,boolean) and throws (IdentityProfile1.java:293-299, EmptyAnnotation.java:22)");
  if (gff) for (int i = 0; i < annot->n; i++) if (annot->name_id[i] < 0 || annot->name_id[i] >= annot->n) return fail(ctx, NPT_EINVAL, "name_id out of range");
  const int n_orf = gff ? 0 : annot->n;  // ORF rows (shared-memory tables, spliced/unspliced counters): CSV only
  const bool calc_breaks1 = cf.calc_breaks && cf.coronavirus && (int64_t)cf.break_thresh < ref_len;  // ViralTranscriptAnalysisCmd2.java:882
  if ((cf.record_depth || calc_breaks1) && ref_len + 2 >= (1LL << 29))
    return fail(ctx, NPT_EINVAL, "dense coverage / break-point matrix need seqlen < 2^29");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  free_batches(ctx);
  ctx->in_chrom = true; ctx->finalized = false;
  ctx->chrom_index = chrom_index; ctx->G = ref_len; ctx->n_orf = n_orf;
  ctx->idx_base = ctx->idx_user_base;
  ctx->walk_ms = ctx->fin_ms = 0; ctx->walk_launches = ctx->total_launches = 0; ctx->n_slow = 0;

  CK(cudaEventRecord(ctx->ev_c0, ctx->s_comp));
  DevCfg& dc = ctx->dc;
  memset(&dc, 0, sizeof dc);
  dc.T = cf.break_thresh; dc.bin = cf.bin; dc.include_start = cf.include_start; dc.coronavirus = cf.coronavirus;
  dc.start_thresh = cf.start_thresh; dc.end_thresh = cf.end_thresh; dc.enforce_strand = cf.enforce_strand;
  dc.record_depth = cf.record_depth; dc.calc_breaks1 = calc_breaks1; dc.fit = cf.first_is_transcriptome ? 1 : 0; dc.track_sums = cf.track_break_sums;
  dc.S = cf.num_sources; dc.min_fl_exon = cf.coronavirus ? 0 : 10;
  dc.rna_mask = 0;
  for (int s = 0; s < cf.num_sources; s++) if (cf.rna[s]) dc.rna_mask |= 1u << s;
  dc.annot_kind = annot->kind; dc.n_orf = n_orf; dc.G = (int)ref_len; dc.stride = (int)ref_len + 2;

  // reference: 8 bases per word, first base in the top nibble, 4 words of padding; and the same bases on the position grid
  // of the fast walk: word W = positions 8W..8W+7 (1-based).  The codes are uploaded as they are (one byte per base) and
  // packed by a kernel: two host loops over a 250 Mb chromosome cost more than walking its reads.
  const int nwords = (int)((ref_len + 7) / 8) + 4;
  const int gwords = (int)(ref_len >> 3) + 1 + 4;
  CK(ctx->d_ref.ensure(nwords)); CK(ctx->d_refg.ensure(gwords));
  CK(ctx->d_ref_codes.ensure((size_t)ref_len));
  CK(cudaMemcpyAsync(ctx->d_ref_codes.p, ref_codes, (size_t)ref_len, cudaMemcpyHostToDevice, ctx->s_comp));
  pack_reference_kernel<<<ctx->sm_count * 8, 256, 0, ctx->s_comp>>>(ctx->d_ref_codes.p, (long long)ref_len, ctx->d_ref.p, nwords, ctx->d_refg.p, gwords);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->s_comp));   // ref_codes is the caller's again when npt_begin_chrom returns
  dc.ref_words = ctx->d_ref.p; dc.ref_nwords = nwords;
  dc.refg_words = ctx->d_refg.p; dc.refg_nwords = gwords;
  CK(ctx->d_ostart.ensure(n_orf)); CK(ctx->d_oname.ensure(n_orf)); CK(ctx->d_ostrand.ensure(n_orf));
  if (n_orf) {
    CK(cudaMemcpy(ctx->d_ostart.p, annot->start, (size_t)n_orf * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_oname.p, annot->name_id, (size_t)n_orf * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_ostrand.p, annot->strand, (size_t)n_orf, cudaMemcpyHostToDevice));
  }
  if (gff) {
    // exon rows sorted by start (ties: file order) so that a block's candidates are one contiguous range
    // (GFFAnnotation.nextUpstream :80-100 scans every row; the accepted one has |l1 - start| <= 2*tolerance)
    std::vector<int> ord(annot->n);
    for (int i = 0; i < annot->n; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return annot->start[a] < annot->start[b]; });
    std::vector<int4> gx(annot->n);
    for (int i = 0; i < annot->n; i++) { const int o = ord[i]; gx[i] = make_int4(annot->start[o], annot->end[o], (o << 1) | (annot->strand[o] ? 1 : 0), annot->name_id[o]); }
    CK(ctx->d_gx.ensure(annot->n)); CK(ctx->d_gxhash.ensure(annot->n));
    if (annot->n) {
      CK(cudaMemcpy(ctx->d_gx.p, gx.data(), (size_t)annot->n * sizeof(int4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(ctx->d_gxhash.p, annot->name_hash, (size_t)annot->n * 4, cudaMemcpyHostToDevice));
    }
    dc.gx = ctx->d_gx.p; dc.gx_n = annot->n; dc.gx_hash = ctx->d_gxhash.p;
  }
  dc.orf_start = ctx->d_ostart.p; dc.orf_name = ctx->d_oname.p; dc.orf_strand = ctx->d_ostrand.p;

  // tables
  const int max_cl = cf.max_clusters > 0 ? cf.max_clusters : (1 << 16);
  const int max_sb = cf.max_subclusters > 0 ? cf.max_subclusters : (1 << 21);
  const int max_pr = cf.max_break_pairs > 0 ? cf.max_break_pairs : (1 << 20);
  // dense coverage rows (COV_ARRAYS int32 rows of seqlen + 2 per cluster and source): default = what fits 12 GB (10 000 clusters of a 30 kb genome, one source), between 256
  // clusters and max_clusters; the caller may ask for as many as HBM holds (rows are kept zero between chromosomes, so a
  // large capacity costs memory but no time)
  const size_t row_bytes = (size_t)COV_ARRAYS * cf.num_sources * ((size_t)ref_len + 2) * 4;
  const int auto_cov = (int)std::min<size_t>((size_t)max_cl, std::max<size_t>(256, ((size_t)12 << 30) / row_bytes));
  const int max_cov = cf.record_depth ? (cf.max_coverage_clusters > 0 ? cf.max_coverage_clusters : auto_cov) : 0;
  const unsigned ncl = pow2_at_least(2ull * max_cl), nsb = pow2_at_least(2ull * max_sb), npr = calc_breaks1 ? pow2_at_least(2ull * max_pr) : 1;
  CK(ctx->d_cl.ensure(ncl)); CK(ctx->d_sb.ensure(nsb)); CK(ctx->d_pr.ensure(npr));
  dc.cl = ctx->d_cl.p; dc.cl_mask = ncl - 1; dc.cl_cap = max_cl;
  dc.sb = ctx->d_sb.p; dc.sb_mask = nsb - 1; dc.sb_cap = max_sb;
  dc.pr = ctx->d_pr.p; dc.pr_mask = npr - 1; dc.pr_cap = max_pr;
  // record arenas: a thread that loses the race for an empty slot leaks the record it had written, and when a chromosome
  // starts every resident thread can race for the same hot key at once: the slack covers 16 words for each of them
  const unsigned long long slack = std::max<unsigned long long>(1ull << 20, 16ull * ctx->sm_count * 2048);
  dc.tok_cap = (unsigned)std::min<unsigned long long>(24ull * max_cl + slack, 1ull << 30);
  dc.sbkey_cap = (unsigned)std::min<unsigned long long>(12ull * max_sb + slack, 1ull << 30);
  dc.sums_cap = cf.track_break_sums ? dc.sbkey_cap : 0;
  CK(ctx->d_cl_rec.ensure(dc.tok_cap)); CK(ctx->d_sb_rec.ensure(dc.sbkey_cap));
  CK(ctx->d_sub_count.ensure((size_t)nsb * dc.S)); CK(ctx->d_sub_break_sum.ensure(nsb));
  CK(ctx->d_sub_sums.ensure(std::max<unsigned>(dc.sums_cap, 1)));
  dc.cl_rec = ctx->d_cl_rec.p; dc.sb_rec = ctx->d_sb_rec.p; dc.sub_count = ctx->d_sub_count.p; dc.sub_break_sum = ctx->d_sub_break_sum.p;
  dc.sub_sums = ctx->d_sub_sums.p;
  if (dc.fit) {
    CK(ctx->d_sub_min0.ensure(nsb));
    CK(cudaMemsetAsync(ctx->d_sub_min0.p, 0xFF, (size_t)nsb * 8, ctx->s_comp));
  }
  dc.sub_min0 = ctx->d_sub_min0.p;
  CK(cudaMemsetAsync(ctx->d_sub_count.p, 0, (size_t)nsb * dc.S * 4, ctx->s_comp));
  CK(cudaMemsetAsync(ctx->d_sub_break_sum.p, 0, (size_t)nsb * 4, ctx->s_comp));
  CK(cudaMemsetAsync(ctx->d_sub_sums.p, 0, (size_t)std::max<unsigned>(dc.sums_cap, 1) * 8, ctx->s_comp));
  const size_t cov_elems = (size_t)COV_ARRAYS * max_cov * dc.S * dc.stride;
  if (cf.record_depth) {
    const bool fresh = ctx->d_cov.cap < cov_elems;
    cudaError_t e = ctx->d_cov.ensure(cov_elems);
    if (e != cudaSuccess) return fail(ctx, NPT_ENOMEM, "coverage rows: %zu bytes for %d clusters (lower max_coverage_clusters)", cov_elems * 4, max_cov);
    // every row is zero when a chromosome begins: npt_end_chrom clears the rows it used; only a new allocation, or a
    // chromosome that did not reach the end of npt_end_chrom, is cleared as a whole
    if (fresh || ctx->cov_dirty) CK(cudaMemsetAsync(ctx->d_cov.p, 0, ctx->d_cov.cap * 4, ctx->s_comp));
    ctx->cov_dirty = true;
  }
  dc.cov = ctx->d_cov.p; dc.cov_cap = max_cov;
  const int nsmall = dc.S * (1 + 2 * dc.n_orf);
  CK(ctx->d_small.ensure(nsmall));
  CK(cudaMemsetAsync(ctx->d_small.p, 0, (size_t)nsmall * 4, ctx->s_comp));
  dc.small = ctx->d_small.p;
  CK(ctx->d_counters.ensure(CN_COUNT));
  CK(cudaMemsetAsync(ctx->d_counters.p, 0, CN_COUNT * 4, ctx->s_comp));
  dc.counters = ctx->d_counters.p;
  init_tables_kernel<<<ctx->sm_count * 4, 256, 0, ctx->s_comp>>>(dc.cl, ncl, dc.sb, nsb, dc.pr, npr);
  CK(cudaGetLastError());

  // shared memory plan: the packed reference lives on chip when it fits (30 kb genome = 15 KB)
  ctx->ref_smem = (size_t)nwords * 4 <= 96 * 1024;
  if (const char* e = getenv("NPT_REF_SMEM")) if (atoi(e) == 0) ctx->ref_smem = false;  // experiments: reference through L1 instead
  ctx->walk_smem = (ctx->ref_smem ? (size_t)((nwords + 3) & ~3) * 4 : 0) + (size_t)(WALK_THREADS / 32) * 2 * 32 * TILE_STRIDE * 4 + 36 * 4;
  ctx->cluster_smem = (size_t)dc.n_orf * 12 + (size_t)dc.S * (1 + 2 * dc.n_orf) * 4 + 16;
  if (ctx->cluster_smem > ctx->smem_optin) return fail(ctx, NPT_EINVAL, "annotation too large for shared memory (%zu bytes)", ctx->cluster_smem);
  CK(cudaFuncSetAttribute(cluster_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  CK(cudaFuncSetAttribute(cluster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  CK(cudaFuncSetAttribute(cluster_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  {
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cluster_kernel<0>, CLUSTER_THREADS, ctx->cluster_smem));
    ctx->cluster_occ = std::max(1, occ);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_key_kernel, 256, 0));
    ctx->covkey_occ = std::max(1, occ);
  }
  if (const char* e = getenv("NPT_SPLIT")) ctx->split_streams = atoi(e) != 0;
  {
    // persistent grids: resident blocks per SM from the occupancy API (NPT_WALK_BLOCKS_PER_SM overrides, for experiments)
    const char* env = getenv("NPT_WALK_BLOCKS_PER_SM");
    const int forced = env ? atoi(env) : 0;
    int occ[3] = {1, 1, 1};
    if (ctx->ref_smem) {
      CK(cudaFuncSetAttribute(walk_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
      CK(cudaFuncSetAttribute(walk_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
      CK(cudaFuncSetAttribute(walk_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], walk_kernel<0, true>, WALK_THREADS, ctx->walk_smem));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], walk_kernel<1, true>, WALK_THREADS, ctx->walk_smem));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], walk_kernel<2, true>, WALK_THREADS, ctx->walk_smem));
    } else {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], walk_kernel<0, false>, WALK_THREADS, ctx->walk_smem));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], walk_kernel<1, false>, WALK_THREADS, ctx->walk_smem));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], walk_kernel<2, false>, WALK_THREADS, ctx->walk_smem));
    }
    for (int k = 0; k < 3; k++) {
      int bps = std::max(1, occ[k]);
      if (forced > 0) bps = std::min(bps, forced);
      ctx->walk_blocks[k] = ctx->sm_count * bps;
    }
    ctx->walk_blocks[2] = ctx->sm_count;  // the store-all pass only touches the rare long-break-list reads
  }
  {
    // fast walk: tasks are 16 positions, so a break can only hide inside one when break_thresh < 16; with depth recording the
    // reference-grid entries of a chromosome must fit accumulate_kernel's shared tables.  Otherwise: serial walker for all.
    ctx->fast_record = cf.record_depth != 0;
    ctx->fast = cf.break_thresh >= FW_PIECE && ref_len < (1LL << 29) &&
                (!ctx->fast_record || ((dc.stride + 31) / 32 + 1 <= ACC_MAX_ENT && (((unsigned long long)dc.cov_cap * dc.S + 1) << ACC_POS_BITS) < (1ull << 32)));
    if (const char* e = getenv("NPT_FAST")) if (atoi(e) == 0) ctx->fast = false;   // experiments / A-B runs
    if (ctx->fast) {
      const size_t refb = ctx->ref_smem ? (size_t)((dc.refg_nwords + 3) & ~3) * 4 : 0;
      const size_t per_warp = (size_t)fw_warp_words(ctx->fast_record) * 4;
      int warps = (int)std::min<size_t>(FW_WARPS_MAX, (ctx->smem_optin - refb - 64) / per_warp);
      if (warps < 1) ctx->fast = false;
      else {
        ctx->fw_threads = warps * 32;
        ctx->fw_smem = refb + (size_t)warps * per_warp;
        int occ = 1;
        if (ctx->ref_smem) {
          if (ctx->fast_record) { CK(cudaFuncSetAttribute(fast_walk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fast_walk_kernel<true, true>, ctx->fw_threads, ctx->fw_smem)); }
          else { CK(cudaFuncSetAttribute(fast_walk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fast_walk_kernel<true, false>, ctx->fw_threads, ctx->fw_smem)); }
        } else {
          if (ctx->fast_record) { CK(cudaFuncSetAttribute(fast_walk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fast_walk_kernel<false, true>, ctx->fw_threads, ctx->fw_smem)); }
          else { CK(cudaFuncSetAttribute(fast_walk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fast_walk_kernel<false, false>, ctx->fw_threads, ctx->fw_smem)); }
        }
        ctx->fw_blocks = ctx->sm_count * std::max(1, occ);
        if (!ctx->fast_record) {
          CK(ctx->d_fw_scratch.ensure((size_t)ctx->fw_blocks * ctx->fw_threads * FW_BRK_BUF));
          dc.fw_scratch = ctx->d_fw_scratch.p;
        }
        if (ctx->fast_record) {
          CK(cudaFuncSetAttribute(accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem)));
          ctx->acc_blocks = ctx->sm_count;
          ctx->acc_pos_shift = 0;   // coarse start position of a read in the sort key: reference-grid entry >> shift < 2^ACC_POS_BITS
          while ((((dc.stride + 31) / 32 + 1) >> ctx->acc_pos_shift) >= (1 << ACC_POS_BITS)) ctx->acc_pos_shift++;
        }
      }
    }
  }
  CK(cudaStreamSynchronize(ctx->s_comp));
  return NPT_OK;
}

}  // extern "C"

// npt_submit / npt_submit_packed: exactly one of b, pb is set
static int submit_impl(npt_ctx* ctx, const npt_batch* b, const npt_packed_batch* pb) {
  if (!ctx || (!b && !pb)) return fail(ctx, NPT_EINVAL, "null argument");
  if (!ctx->in_chrom || ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_submit outside npt_begin_chrom..npt_end_chrom");
  const int64_t n = b ? b->n : pb->n;
  if (n < 0 || n > (1LL << 31) - 1) return fail(ctx, NPT_EINVAL, "batch size out of range");
  if (n == 0) return NPT_OK;
  if (ctx->idx_base + n >= (1LL << 40)) return fail(ctx, NPT_EINVAL, "more than 2^40 reads");
  CK(cudaSetDevice(ctx->device));
  {
    // Retire what has finished, and keep at most two batches in flight: a batch holds ~1.1 KB of scratch per read (coverage
    // entries, headers, sort buffers) until it is retired, and a long run of device-resident waves (500 M reads as 50
    // submits) would otherwise pile all of it up.  The stream runs in order, so waiting for the oldest costs nothing.
    int in_flight = 0;
    for (size_t i = 0; i < ctx->batches.size(); i++)
      if (!ctx->batches[i].retired) {
        if (cudaEventQuery(ctx->batches[i].done) == cudaSuccess) { int rc = retire_batch(ctx, i); if (rc) return rc; }
        else in_flight++;
      }
    for (size_t i = 0; i < ctx->batches.size() && in_flight >= 2; i++)
      if (!ctx->batches[i].retired) { int rc = retire_batch(ctx, i); if (rc) return rc; in_flight--; }
  }
  Batch B;
  B.n = n;
  BatchDev& d = B.d;
  d.n = n; d.idx_base = ctx->idx_base;
  const uint16_t* host_src = nullptr;
  const int location = b ? b->location : NPT_MEM_HOST;

  if (location == NPT_MEM_HOST) {
    const uint32_t* h_coff = b ? b->cigar_off : pb->cigar_off;
    const uint64_t* h_soff = b ? b->seq_off : pb->seq_off;
    const uint32_t ncig = h_coff[n];
    const uint64_t nseq = h_soff[n];
    if (pb && (pb->n_wide < 0 || pb->n_exc < 0 || (pb->n_wide && (!pb->wide_idx || !pb->wide_word)) || (pb->n_exc && (!pb->exc_byte || !pb->exc_val))))
      return fail(ctx, NPT_EINVAL, "packed batch: wide / exception lists inconsistent");
    // staging layout (256-byte aligned sections)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_pos = 0, o_flag = al(o_pos + (size_t)n * 4), o_src = al(o_flag + (size_t)n * 2), o_coff = al(o_src + (size_t)n * 2);
    const size_t o_cig = al(o_coff + (size_t)(n + 1) * 4), o_soff = al(o_cig + (size_t)ncig * 4), o_seq = al(o_soff + (size_t)(n + 1) * 8);
    const size_t unpacked = al(o_seq + nseq + 16);
    // compact transfer format: the packed arrays land behind the expanded layout and two copy kernels fill it
    const size_t p_cig = unpacked, p_seq = al(p_cig + (size_t)ncig * 2 + 16), p_widx = al(p_seq + (size_t)((nseq + 1) / 2) + 16);
    const size_t p_wword = pb ? al(p_widx + (size_t)pb->n_wide * 4) : p_widx, p_eidx = pb ? al(p_wword + (size_t)pb->n_wide * 4) : p_widx;
    const size_t p_eval = pb ? al(p_eidx + (size_t)pb->n_exc * 8) : p_widx;
    const size_t total = pb ? al(p_eval + (size_t)pb->n_exc + 16) : unpacked;
    int si = -1;
    for (int k = 0; k < 2; k++) if (!ctx->stg[k].busy) { si = k; break; }
    if (si < 0) {  // both in flight: retire the older one
      int older = ctx->stg[0].batch_index < ctx->stg[1].batch_index ? 0 : 1;
      int rc = retire_batch(ctx, (size_t)ctx->stg[older].batch_index);
      if (rc) return rc;
      si = older;
    }
    Staging& S = ctx->stg[si];
    CK(S.buf.ensure(total));
    unsigned char* base = S.buf.p;
    CK(cudaMemcpyAsync(base + o_pos, b ? b->pos : pb->pos, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->s_copy));
    CK(cudaMemcpyAsync(base + o_flag, b ? b->flag : pb->flag, (size_t)n * 2, cudaMemcpyHostToDevice, ctx->s_copy));
    CK(cudaMemcpyAsync(base + o_coff, h_coff, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, ctx->s_copy));
    CK(cudaMemcpyAsync(base + o_soff, h_soff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->s_copy));
    if (b) {
      if (ncig) CK(cudaMemcpyAsync(base + o_cig, b->cigar, (size_t)ncig * 4, cudaMemcpyHostToDevice, ctx->s_copy));
      if (nseq) CK(cudaMemcpyAsync(base + o_seq, b->seq4, (size_t)nseq, cudaMemcpyHostToDevice, ctx->s_copy));
    } else {
      if (ncig) CK(cudaMemcpyAsync(base + p_cig, pb->cigar16, (size_t)ncig * 2, cudaMemcpyHostToDevice, ctx->s_copy));
      if (nseq) CK(cudaMemcpyAsync(base + p_seq, pb->seq2, (size_t)((nseq + 1) / 2), cudaMemcpyHostToDevice, ctx->s_copy));
      if (pb->n_wide) {
        CK(cudaMemcpyAsync(base + p_widx, pb->wide_idx, (size_t)pb->n_wide * 4, cudaMemcpyHostToDevice, ctx->s_copy));
        CK(cudaMemcpyAsync(base + p_wword, pb->wide_word, (size_t)pb->n_wide * 4, cudaMemcpyHostToDevice, ctx->s_copy));
      }
      if (pb->n_exc) {
        CK(cudaMemcpyAsync(base + p_eidx, pb->exc_byte, (size_t)pb->n_exc * 8, cudaMemcpyHostToDevice, ctx->s_copy));
        CK(cudaMemcpyAsync(base + p_eval, pb->exc_val, (size_t)pb->n_exc, cudaMemcpyHostToDevice, ctx->s_copy));
      }
    }
    CK(cudaEventRecord(S.done, ctx->s_copy));
    CK(cudaStreamWaitEvent(ctx->s_comp, S.done, 0));
    if (pb) {   // expand on the compute stream, so that the copy stream is free for the next batch
      const int ug = ctx->sm_count * 8;
      if (ncig) unpack_cigar_kernel<<<ug, 256, 0, ctx->s_comp>>>((const uint16_t*)(base + p_cig), (uint32_t*)(base + o_cig), (long long)ncig);
      if (pb->n_wide) patch_wide_kernel<<<std::max(1, (int)std::min<int64_t>((pb->n_wide + 255) / 256, ug)), 256, 0, ctx->s_comp>>>(
          (const uint32_t*)(base + p_widx), (const uint32_t*)(base + p_wword), (long long)pb->n_wide, (uint32_t*)(base + o_cig), (long long)ncig);
      if (nseq) unpack_seq_kernel<<<ug, 256, 0, ctx->s_comp>>>(base + p_seq, base + o_seq, (long long)nseq);
      if (pb->n_exc) patch_exceptions_kernel<<<std::max(1, (int)std::min<int64_t>((pb->n_exc + 255) / 256, ug)), 256, 0, ctx->s_comp>>>(
          (const unsigned long long*)(base + p_eidx), base + p_eval, (long long)pb->n_exc, base + o_seq, (long long)nseq);
      CK(cudaGetLastError());
      ctx->total_launches += (ncig ? 1 : 0) + (pb->n_wide ? 1 : 0) + (nseq ? 1 : 0) + (pb->n_exc ? 1 : 0);
    }
    d.pos = (const int*)(base + o_pos); d.flag = (const uint16_t*)(base + o_flag);
    host_src = b ? b->src : pb->src;   // copied below into the batch's own output block: fit_fixup_kernel reads it at end of chromosome,
                         // long after this staging buffer has been handed to a later batch
    d.cigar_off = (const uint32_t*)(base + o_coff); d.cigar = (const uint32_t*)(base + o_cig);
    d.seq_off = (const unsigned long long*)(base + o_soff); d.seq4 = base + o_seq; d.seq_bytes = (long long)nseq; d.cigar_words = (long long)ncig;
    S.busy = true; S.batch_index = (int)ctx->batches.size();
    B.staging = si;
  } else if (location == NPT_MEM_DEVICE) {
    if (((uintptr_t)b->seq4 & 15) != 0 || ((uintptr_t)b->cigar & 15) != 0)
      return fail(ctx, NPT_EINVAL, "device seq4 and cigar must be 16-byte aligned (the walk stages them with 16-byte cp.async copies)");
    uint64_t nseq = 0;
    uint32_t ncig = 0;
    CK(cudaMemcpy(&nseq, b->seq_off + n, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ncig, b->cigar_off + n, 4, cudaMemcpyDeviceToHost));
    d.pos = b->pos; d.flag = b->flag; d.src = b->src; d.cigar_off = b->cigar_off; d.cigar = b->cigar;
    d.seq_off = (const unsigned long long*)b->seq_off; d.seq4 = b->seq4; d.seq_bytes = (long long)nseq; d.cigar_words = (long long)ncig;
    B.staging = -1;
  } else {
    return fail(ctx, NPT_EINVAL, "unknown batch location");
  }

  // per-read outputs: 11 int arrays (+ the source indices of a host batch) in one block + break arena
  CK(pool_get(ctx, &B.out_block, (size_t)n * 11 * 4 + (size_t)n * 2 + 16, &B.out_bytes));
  int* ob = (int*)B.out_block;
  if (host_src) {
    uint16_t* keep = (uint16_t*)(ob + 11 * n);
    CK(cudaMemcpyAsync(keep, host_src, (size_t)n * 2, cudaMemcpyHostToDevice, ctx->s_copy));
    CK(cudaEventRecord(ctx->stg[B.staging].done, ctx->s_copy));
    CK(cudaStreamWaitEvent(ctx->s_comp, ctx->stg[B.staging].done, 0));
    d.src = keep;
  }
  d.cluster = ob; d.sub = ob + n; d.start_pos = ob + 2 * n; d.end_pos = ob + 3 * n; d.read_st = ob + 4 * n; d.read_en = ob + 5 * n;
  d.rflags = (unsigned*)(ob + 6 * n); d.maps_sum = ob + 7 * n; d.err_sum = ob + 8 * n; d.nbreaks = ob + 9 * n; d.break_loc = (unsigned*)(ob + 10 * n);
  // overflow area for break lists longer than the inline slot: a quarter of an int per read for virus-shaped reads (a few
  // per cent have a third junction), 24 ints per read in host mode (spliced reads: two breaks per junction), or what
  // earlier batches of this chromosome needed; too small is not an error (the batch is redone with the exact size)
  const double per_read = std::max<double>(ctx->cfg.coronavirus ? 0.25 : 24.0, 1.25 * ctx->brk_per_read);
  d.break_cap = (unsigned)std::min<int64_t>(std::max<int64_t>((int64_t)(n * per_read), 4096), 1LL << 30);
  CK(pool_get(ctx, (void**)&d.break_arena, ((size_t)n * BRK_INLINE + d.break_cap) * sizeof(int), &B.arena_bytes));
  CK(pool_get(ctx, (void**)&B.bctr, 12 * sizeof(unsigned), &B.bctr_bytes));
  CK(cudaMemsetAsync(B.bctr, 0, 12 * sizeof(unsigned), ctx->s_comp));
  d.bctr = B.bctr;
  // fast walk: work list of the serial walker; with depth recording also headers, the entry arena (16 bytes per 32 reference
  // positions a read covers: its share is (bytes of bases / 16) + 8 entries) and the sort buffers
  // 2: the serial walker's chain runs on the second stream — worth it while that chain is a few reads (latency of one warp);
  // with many of them (host mode: 1 % of the reads have more junctions than the fast walk keeps) the two streams only get in
  // each other's way (98.7 -> 106.8 ms on the host-mode bench), so the last retired batch's share decides
  d.fast = ctx->fast ? ((ctx->split_streams && ctx->slow_share < 5e-4f) ? 2 : 1) : 0;
  if (d.fast && ((uint64_t)d.seq_bytes >> 4) + 2 >= (1ull << 32)) d.fast = 0;  // the walk names 16-byte chunks of the bases with 32 bits
  if (d.fast && ctx->fast_record && ((uint64_t)d.seq_bytes >> 4) + 8ull * (uint64_t)n + 8 >= (1ull << 32)) d.fast = 0;  // entry index is 32 bits
  if (d.fast) {
    CK(pool_get(ctx, &B.f_slow, (size_t)n * 4, &B.f_slow_bytes));
    d.slow_list = (unsigned*)B.f_slow;
    if (ctx->fast_record) {
      const size_t n_ent = ((size_t)d.seq_bytes >> 4) + 8 * (size_t)n + 8;
      // headers: uint4[n] + meta u32[n] + rare records u32[n][16], in one block (16-byte aligned sections)
      const size_t hA = (size_t)n * 16, hM = (((size_t)n * 4) + 15) & ~(size_t)15, hR = (size_t)n * FW_RARE_WORDS * 4;
      CK(pool_get(ctx, &B.f_hdr, hA + hM + hR, &B.f_hdr_bytes));
      CK(pool_get(ctx, &B.f_ent, n_ent * 16, &B.f_ent_bytes));
      d.fhdrA = (uint4*)B.f_hdr; d.fmeta = (unsigned*)((unsigned char*)B.f_hdr + hA); d.frare = (unsigned*)((unsigned char*)B.f_hdr + hA + hM);
      d.fent = (uint4*)B.f_ent;
      unsigned long long maxkey = (((unsigned long long)ctx->dc.cov_cap * (unsigned)ctx->dc.S + 1) << ACC_POS_BITS) - 1;
      B.key_bits = 1;
      while ((1ull << B.key_bits) <= maxkey) B.key_bits++;
      size_t tmp = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (int)n, 0, B.key_bits, ctx->s_comp));
      const size_t an = ((size_t)n * 4 + 255) & ~(size_t)255;
      CK(pool_get(ctx, &B.f_sort, 4 * an + tmp + 256, &B.f_sort_bytes));
      unsigned char* sb = (unsigned char*)B.f_sort;
      B.k_in = (unsigned*)sb; B.v_in = (unsigned*)(sb + an); B.k_out = (unsigned*)(sb + 2 * an); B.v_out = (unsigned*)(sb + 3 * an);
      B.cub_tmp = sb + 4 * an; B.cub_bytes = tmp;
    }
  }

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, ctx->s_comp));
  int rc = launch_batch(ctx, B, 0);
  if (rc) return rc;
  ctx->walk_launches += (ctx->dc.record_depth ? 4 : 3) + (d.fast ? 1 : 0);
  CK(cudaEventRecord(e1, ctx->s_comp));
  ctx->walk_events.emplace_back(e0, e1);
  CK(cudaEventCreateWithFlags(&B.done, cudaEventDisableTiming));
  CK(cudaEventRecord(B.done, ctx->s_comp));
  ctx->idx_base += n;
  ctx->batches.push_back(B);
  return NPT_OK;
}

extern "C" {

int npt_submit(npt_ctx* ctx, const npt_batch* b) { return submit_impl(ctx, b, nullptr); }
int npt_submit_packed(npt_ctx* ctx, const npt_packed_batch* pb) { return submit_impl(ctx, nullptr, pb); }

int npt_wait(npt_ctx* ctx) {
  if (!ctx) return NPT_EINVAL;
  CK(cudaSetDevice(ctx->device));
  for (size_t i = 0; i < ctx->batches.size(); i++) { int rc = retire_batch(ctx, i); if (rc) return rc; }
  CK(cudaStreamSynchronize(ctx->s_copy));
  CK(cudaStreamSynchronize(ctx->s_comp));
  return NPT_OK;
}

static int sort_pairs(npt_ctx* ctx, unsigned long long* kin, unsigned long long* kout, int* vin, int* vout, int n) {
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, 64, ctx->s_comp));
  CK(ctx->d_cub.ensure(tmp));
  CK(cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, tmp, kin, kout, vin, vout, n, 0, 64, ctx->s_comp));
  ctx->total_launches += 8;
  return NPT_OK;
}

static int coverage_rows(npt_ctx* ctx, const int* d_row_of_prov, int n_rows) {
  const DevCfg& dc = ctx->dc;
  const int n_prov = std::min(ctx->n_clusters, dc.cov_cap);
  const size_t elems = 4ull * std::max(n_rows, 1) * dc.S * dc.stride;
  CK(ctx->d_cov_out.ensure(elems));
  CK(cudaMemsetAsync(ctx->d_cov_out.p, 0, elems * 4, ctx->s_comp));
  if (n_prov > 0) {
    coverage_rows_kernel<<<4 * n_prov * dc.S, SCAN_THREADS, 0, ctx->s_comp>>>(dc.cov, dc.cov_cap, n_prov, dc.S, dc.stride, d_row_of_prov, n_rows,
                                                                            ctx->d_cov_out.p);
    clear_rows_kernel<<<dim3((unsigned)(((size_t)n_prov * dc.S * dc.stride + 1023) / 1024), COV_ARRAYS), 256, 0, ctx->s_comp>>>(
        dc.cov, (size_t)dc.cov_cap * dc.S * dc.stride, (size_t)n_prov * dc.S * dc.stride);
    ctx->total_launches += 2;
    CK(cudaGetLastError());
  }
  ctx->cov_dirty = false;   // the rows this chromosome used are zero again
  ctx->cov_rows = n_rows;
  return NPT_OK;
}

int npt_end_chrom(npt_ctx* ctx) {
  if (!ctx) return NPT_EINVAL;
  if (!ctx->in_chrom) return fail(ctx, NPT_ESTATE, "npt_end_chrom without npt_begin_chrom");
  if (ctx->finalized) return NPT_OK;
  int rc = npt_wait(ctx);
  if (rc) return rc;
  // walk time = sum over batches of the kernel's event span
  ctx->walk_ms = 0;
  for (auto& e : ctx->walk_events) { float ms = 0; CK(cudaEventElapsedTime(&ms, e.first, e.second)); ctx->walk_ms += ms; }
  for (int k = 0; k < 6; k++) ctx->stage_ms[k] = 0;
  for (size_t i = 0; i + 12 <= ctx->stage_events.size(); i += 12)
    for (int k = 0; k < 6; k++) { float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->stage_events[i + 2 * k], ctx->stage_events[i + 2 * k + 1])); ctx->stage_ms[k] += ms; }
  CK(cudaMemcpy(ctx->h_counters, ctx->d_counters.p, sizeof ctx->h_counters, cudaMemcpyDeviceToHost));
  const unsigned er = ctx->h_counters[CN_ERROR];
  if (er) {
    std::string m = "device table capacity exceeded:";
    if (er & ERR_CL_FULL) m += " clusters(max_clusters)";
    if (er & ERR_SUB_FULL) m += " sub-clusters(max_subclusters)";
    if (er & ERR_PAIR_FULL) m += " break-point pairs(max_break_pairs)";
    if (er & ERR_COV_FULL) m += " coverage rows(max_coverage_clusters)";
    if (er & ERR_ARENA_FULL) m += " key arena";
    if (er & ERR_TOO_MANY_BREAKS) m += " a read has more than 2^20 breaks";
    if (er & ERR_GENESET) m += " a read overlaps more than 32 annotated genes";
    ctx->in_chrom = false;
    if (er & ERR_BAD_SRC) return fail(ctx, NPT_EINVAL, "a record's source index is >= num_sources");
    return fail(ctx, NPT_ECAPACITY, "%s", m.c_str());
  }
  const DevCfg& dc = ctx->dc;
  const int nc = (int)ctx->h_counters[CN_CLUSTERS], ns = (int)ctx->h_counters[CN_SUBS];
  if (ns < nc) return fail(ctx, NPT_ECUDA, "internal: %d sub-clusters for %d clusters", ns, nc);
  ctx->n_clusters = nc; ctx->n_subs = ns;
  CK(cudaEventRecord(ctx->ev_a, ctx->s_comp));
  if (dc.fit && dc.track_sums) {
    for (auto& B : ctx->batches) {
      const int grid = (int)std::min<int64_t>((B.n + 255) / 256, ctx->sm_count * 16);
      fit_fixup_kernel<<<grid, 256, 0, ctx->s_comp>>>(dc, B.d);
      ctx->total_launches++;
    }
    CK(cudaGetLastError());
  }
  const int nmax = std::max(std::max(nc, ns), 1);
  CK(ctx->d_keys.ensure(nmax)); CK(ctx->d_keys2.ensure(nmax)); CK(ctx->d_vals.ensure(nmax)); CK(ctx->d_vals2.ensure(nmax));
  CK(ctx->d_cl_slot.ensure(std::max(nc, 1))); CK(ctx->d_sb_slot.ensure(std::max(ns, 1)));
  CK(ctx->d_cl_rank.ensure(std::max(nc, 1))); CK(ctx->d_cl_sub_off.ensure(nc + 1));
  const unsigned ncl_slots = dc.cl_mask + 1, nsb_slots = dc.sb_mask + 1;
  CK(ctx->d_cl_slot_rank.ensure(ncl_slots)); CK(ctx->d_sb_grank.ensure(nsb_slots));
  if (nc > 0) {
    // sort-and-rank: cluster id = rank of its first read index (CigarClusters.java:83)
    collect_clusters_kernel<<<ctx->sm_count * 2, 256, 0, ctx->s_comp>>>(dc.cl, ncl_slots, dc.cl_rec, nc, ctx->d_keys.p, ctx->d_vals.p, ctx->d_cl_slot.p);
    if ((rc = sort_pairs(ctx, ctx->d_keys.p, ctx->d_keys2.p, ctx->d_vals.p, ctx->d_vals2.p, nc))) return rc;
    ranks_kernel<<<(nc + 255) / 256, 256, 0, ctx->s_comp>>>(ctx->d_vals2.p, nc, ctx->d_cl_rank.p);
    cluster_slot_ranks_kernel<<<(nc + 255) / 256, 256, 0, ctx->s_comp>>>(ctx->d_cl_slot.p, ctx->d_cl_rank.p, nc, ctx->d_cl_slot_rank.p);
    // sub-cluster id = rank of its first read index among the subs of its cluster (CigarCluster.java:660)
    CK(cudaMemsetAsync(ctx->d_cl_sub_off.p, 0, (size_t)(nc + 1) * 4, ctx->s_comp));
    if (ns > 0) {
      unsigned* ctr = ctx->d_counters.p + CN_SUBS;
      CK(cudaMemsetAsync(ctr, 0, 4, ctx->s_comp));
      collect_subs_kernel<<<ctx->sm_count * 2, 256, 0, ctx->s_comp>>>(dc.sb, nsb_slots, dc.sb_rec, dc.cl, dc.cl_rec, ctx->d_cl_rank.p, ns, ctx->d_keys.p,
                                                                       ctx->d_vals.p, ctr);
      if ((rc = sort_pairs(ctx, ctx->d_keys.p, ctx->d_keys2.p, ctx->d_vals.p, ctx->d_vals2.p, ns))) return rc;
      CK(cudaMemcpyAsync(ctx->d_sb_slot.p, ctx->d_vals2.p, (size_t)ns * 4, cudaMemcpyDeviceToDevice, ctx->s_comp));  // sorted slots
      ranks_kernel<<<(ns + 255) / 256, 256, 0, ctx->s_comp>>>(ctx->d_vals2.p, ns, ctx->d_sb_grank.p);  // slot -> global sub rank
      sub_offsets_kernel<<<(ns + 255) / 256, 256, 0, ctx->s_comp>>>(ctx->d_keys2.p, ns, nc, ctx->d_cl_sub_off.p);
    }
    ctx->total_launches += 8;
    CK(cudaGetLastError());
  }
  // per-read: provisional ids -> ranks; break lists -> one compact array in submit order
  int64_t n_total = 0;
  for (auto& B : ctx->batches) n_total += B.n;
  CK(ctx->d_break_off.ensure(n_total + 1));
  int64_t off_r = 0;
  std::vector<long long> batch_break_base(ctx->batches.size() + 1, 0);
  for (size_t bi = 0; bi < ctx->batches.size(); bi++) {
    Batch& B = ctx->batches[bi];
    const int grid = (int)std::min<int64_t>((B.n + 255) / 256, ctx->sm_count * 16);
    if (nc > 0) remap_reads_kernel<<<grid, 256, 0, ctx->s_comp>>>(B.n, B.d.cluster, B.d.sub, ctx->d_cl_slot_rank.p, ctx->d_sb_grank.p, ctx->d_cl_sub_off.p);
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, B.d.nbreaks, ctx->d_break_off.p + off_r, (int)B.n, ctx->s_comp));
    CK(ctx->d_cub.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, tmp, B.d.nbreaks, ctx->d_break_off.p + off_r, (int)B.n, ctx->s_comp));
    long long last_off = 0; int last_n = 0;
    CK(cudaMemcpyAsync(&last_off, ctx->d_break_off.p + off_r + B.n - 1, 8, cudaMemcpyDeviceToHost, ctx->s_comp));
    CK(cudaMemcpyAsync(&last_n, B.d.nbreaks + B.n - 1, 4, cudaMemcpyDeviceToHost, ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_comp));
    batch_break_base[bi + 1] = batch_break_base[bi] + last_off + last_n;
    off_r += B.n;
    ctx->total_launches += 3;
  }
  const long long total_breaks = batch_break_base[ctx->batches.size()];
  CK(ctx->d_breaks_out.ensure(std::max<long long>(total_breaks, 1)));
  off_r = 0;
  for (size_t bi = 0; bi < ctx->batches.size(); bi++) {
    Batch& B = ctx->batches[bi];
    const int grid = (int)std::min<int64_t>((B.n + 255) / 256, ctx->sm_count * 16);
    gather_breaks_kernel<<<grid, 256, 0, ctx->s_comp>>>(B.n, B.d.nbreaks, B.d.break_loc, B.d.break_arena, ctx->d_break_off.p + off_r,
                                                         batch_break_base[bi], ctx->d_breaks_out.p);
    off_r += B.n;
    ctx->total_launches++;
  }
  CK(cudaMemcpyAsync(ctx->d_break_off.p + n_total, &total_breaks, 8, cudaMemcpyHostToDevice, ctx->s_comp));
  CK(cudaGetLastError());
  // break-point pairs -> dense list
  // break-point pairs -> dense list sorted by key = (src, level, start, end): the order of the results, and the same on
  // every rank once the keys were exchanged
  ctx->n_pairs = 0;
  if (dc.calc_breaks1) {
    const int np = (int)ctx->h_counters[CN_PAIRS];
    const size_t np1 = (size_t)std::max(np, 1);
    CK(ctx->d_pair_keys.ensure(np1)); CK(ctx->d_pair_keys2.ensure(np1)); CK(ctx->d_pair_cnt.ensure(np1)); CK(ctx->d_pair_cnt2.ensure(np1));
    unsigned* ctr = ctx->d_counters.p + CN_PAIRS;
    CK(cudaMemsetAsync(ctr, 0, 4, ctx->s_comp));
    compact_pairs_kernel<<<ctx->sm_count * 2, 256, 0, ctx->s_comp>>>(dc.pr, dc.pr_mask + 1, ctx->d_pair_keys.p, ctx->d_pair_cnt.p, ctr, np);
    ctx->total_launches++;
    if (np > 0 && (rc = sort_pairs(ctx, ctx->d_pair_keys.p, ctx->d_pair_keys2.p, ctx->d_pair_cnt.p, ctx->d_pair_cnt2.p, np))) return rc;
    ctx->n_pairs = np;
  }
  // sub-cluster columns in final order (counts, break sums): device copies that fetch reads and a multi-rank caller sums
  {
    const int S = dc.S, nc1 = std::max(nc, 1), ns1 = std::max(ns, 1);
    CK(ctx->d_ex_first.ensure(std::max(nc1, ns1)));
    for (auto& b : ctx->d_ex_u) CK(b.ensure(std::max(nc1, ns1)));
    CK(ctx->d_ex_i[0].ensure((size_t)ns1 * S)); CK(ctx->d_ex_i[1].ensure(ns1));
    CK(ctx->d_ex_flags.ensure(ns1));
    CK(ctx->d_fin_sum_off.ensure(ns1 + 1));
    ctx->total_sums = 0;
    if (ns > 0) {
      export_subs_kernel<<<(ns + 255) / 256, 256, 0, ctx->s_comp>>>(dc.sb, dc.sb_rec, ctx->d_sb_slot.p, ns, S, dc.sub_count, dc.sub_break_sum,
                                                                     ctx->d_ex_first.p, ctx->d_ex_u[0].p, ctx->d_ex_u[1].p, ctx->d_ex_i[0].p, ctx->d_ex_i[1].p,
                                                                     ctx->d_ex_u[2].p, ctx->d_ex_u[3].p, ctx->d_ex_flags.p);
      size_t tmp = 0;
      CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_ex_u[3].p, ctx->d_fin_sum_off.p, ns, ctx->s_comp));
      CK(ctx->d_cub.ensure(tmp));
      CK(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, tmp, ctx->d_ex_u[3].p, ctx->d_fin_sum_off.p, ns, ctx->s_comp));
      unsigned last_off = 0, last_n = 0;
      CK(cudaMemcpyAsync(&last_off, ctx->d_fin_sum_off.p + ns - 1, 4, cudaMemcpyDeviceToHost, ctx->s_comp));
      CK(cudaMemcpyAsync(&last_n, ctx->d_ex_u[3].p + ns - 1, 4, cudaMemcpyDeviceToHost, ctx->s_comp));
      CK(cudaStreamSynchronize(ctx->s_comp));
      ctx->total_sums = dc.track_sums ? (long long)last_off + last_n : 0;
      const unsigned tot = (unsigned)((long long)last_off + last_n);
      CK(cudaMemcpyAsync(ctx->d_fin_sum_off.p + ns, &tot, 4, cudaMemcpyHostToDevice, ctx->s_comp));
      CK(ctx->d_fin_sums.ensure((size_t)std::max<long long>(ctx->total_sums, 1)));
      if (ctx->total_sums)
        gather_sums_kernel<<<(ns + 255) / 256, 256, 0, ctx->s_comp>>>(ns, ctx->d_ex_u[2].p, ctx->d_ex_u[3].p, ctx->d_fin_sum_off.p, dc.sub_sums,
                                                                       ctx->d_fin_sums.p);
      ctx->total_launches += 4;
      CK(cudaGetLastError());
    }
  }
  // coverage: difference rows -> depth, rows in cluster-id order
  if (dc.record_depth) {
    if ((rc = coverage_rows(ctx, ctx->d_cl_rank.p, nc))) return rc;
  }
  CK(cudaEventRecord(ctx->ev_b, ctx->s_comp));
  CK(cudaEventRecord(ctx->ev_c1, ctx->s_comp));
  CK(cudaStreamSynchronize(ctx->s_comp));
  CK(cudaEventElapsedTime(&ctx->fin_ms, ctx->ev_a, ctx->ev_b));
  CK(cudaEventElapsedTime(&ctx->step_ms, ctx->ev_c0, ctx->ev_c1));
  ctx->total_breaks = total_breaks;
  ctx->n_total = n_total;
  ctx->finalized = true;
  ctx->in_chrom = false;
  return NPT_OK;
}

// ------------------------------------------------------------------------------------------------ results

int npt_fetch_results(npt_ctx* ctx, int what, npt_results* out) {
  if (!ctx || !out) return fail(ctx, NPT_EINVAL, "null argument");
  if (!ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_fetch_results before npt_end_chrom");
  CK(cudaSetDevice(ctx->device));
  const DevCfg& dc = ctx->dc;
  const int S = dc.S, nc = ctx->n_clusters, ns = ctx->n_subs;
  npt_results& R = ctx->res;
  memset(&R, 0, sizeof R);
  cudaStream_t st = ctx->s_comp;
  const int64_t n = ctx->n_total;
  R.n_reads = n;
  R.n_clusters = nc; R.n_subs = ns; R.n_orf = dc.n_orf; R.cov_stride = dc.stride;
  R.n_slow_reads = ctx->fast ? ctx->n_slow : ctx->h_counters[CN_SLOW];
  if (what & NPT_FETCH_READS) {
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    if (!ctx->r_cluster.ensure(nn) || !ctx->r_sub.ensure(nn) || !ctx->r_start.ensure(nn) || !ctx->r_end.ensure(nn) || !ctx->r_rst.ensure(nn) ||
        !ctx->r_ren.ensure(nn) || !ctx->r_flags.ensure(nn) || !ctx->r_maps.ensure(nn) || !ctx->r_err.ensure(nn) || !ctx->r_boff.ensure(nn + 1) ||
        !ctx->r_breaks.ensure((size_t)std::max<long long>(ctx->total_breaks, 1)))
      return fail(ctx, NPT_ENOMEM, "pinned result buffers");
    int64_t o = 0;
    for (auto& B : ctx->batches) {
      const size_t bytes = (size_t)B.n * 4;
      CK(cudaMemcpyAsync(ctx->r_cluster.p + o, B.d.cluster, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_sub.p + o, B.d.sub, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_start.p + o, B.d.start_pos, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_end.p + o, B.d.end_pos, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_rst.p + o, B.d.read_st, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_ren.p + o, B.d.read_en, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_flags.p + o, B.d.rflags, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_maps.p + o, B.d.maps_sum, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->r_err.p + o, B.d.err_sum, bytes, cudaMemcpyDeviceToHost, st));
      o += B.n;
    }
    CK(cudaMemcpyAsync(ctx->r_boff.p, ctx->d_break_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (ctx->total_breaks) CK(cudaMemcpyAsync(ctx->r_breaks.p, ctx->d_breaks_out.p, (size_t)ctx->total_breaks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    R.cluster_id = ctx->r_cluster.p; R.sub_id = ctx->r_sub.p; R.start_pos = ctx->r_start.p; R.end_pos = ctx->r_end.p;
    R.read_st = ctx->r_rst.p; R.read_en = ctx->r_ren.p; R.read_flags = ctx->r_flags.p; R.maps_sum = ctx->r_maps.p; R.err_sum = ctx->r_err.p;
    R.break_off = (const uint64_t*)ctx->r_boff.p; R.breaks = ctx->r_breaks.p;
  }
  if (what & NPT_FETCH_TABLES) {
    const int nc1 = std::max(nc, 1), ns1 = std::max(ns, 1);
    CK(ctx->d_exc_first.ensure(nc1)); CK(ctx->d_exc_u[0].ensure(nc1)); CK(ctx->d_exc_u[1].ensure(nc1));
    // clusters
    std::vector<long long> first(nc1);
    std::vector<unsigned> toff(nc1), ntok(nc1);
    if (nc > 0) {
      export_clusters_kernel<<<(nc + 255) / 256, 256, 0, st>>>(dc.cl, dc.cl_rec, ctx->d_cl_slot.p, ctx->d_cl_rank.p, nc, ctx->d_exc_first.p,
                                                              ctx->d_exc_u[0].p, ctx->d_exc_u[1].p);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(first.data(), ctx->d_exc_first.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(toff.data(), ctx->d_exc_u[0].p, (size_t)nc * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ntok.data(), ctx->d_exc_u[1].p, (size_t)nc * 4, cudaMemcpyDeviceToHost, st));
    }
    std::vector<int> tok_arena(std::max<unsigned>(ctx->h_counters[CN_TOK_USED], 1));
    CK(cudaMemcpyAsync(tok_arena.data(), dc.cl_rec, (size_t)ctx->h_counters[CN_TOK_USED] * 4, cudaMemcpyDeviceToHost, st));
    std::vector<int> cl_sub_off(nc + 1, 0);
    CK(cudaMemcpyAsync(cl_sub_off.data(), ctx->d_cl_sub_off.p, (size_t)(nc + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->t_cl_first.assign(first.begin(), first.begin() + nc);
    ctx->t_cl_key_off.assign(nc + 1, 0);
    ctx->t_cl_key_tok.clear();
    for (int r = 0; r < nc; r++) {
      ctx->t_cl_key_off[r] = (uint32_t)ctx->t_cl_key_tok.size();
      ctx->t_cl_key_tok.insert(ctx->t_cl_key_tok.end(), tok_arena.begin() + toff[r], tok_arena.begin() + toff[r] + ntok[r]);
    }
    ctx->t_cl_key_off[nc] = (uint32_t)ctx->t_cl_key_tok.size();
    ctx->t_cl_sub_off.assign(cl_sub_off.begin(), cl_sub_off.end());
    // sub-clusters: columns were put in final order on the device at end of chromosome (and possibly summed across ranks since)
    std::vector<long long> sfirst(ns1);
    std::vector<unsigned> koff(ns1), nkey(ns1);
    std::vector<unsigned char> sflags(ns1);
    ctx->t_sub_count.assign((size_t)ns * S, 0); ctx->t_sub_break_sum.assign(ns, 0);
    ctx->t_sub_sum_off.assign(ns + 1, 0); ctx->t_sub_sums.assign((size_t)ctx->total_sums, 0);
    if (ns > 0) {
      CK(cudaMemcpyAsync(sfirst.data(), ctx->d_ex_first.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(koff.data(), ctx->d_ex_u[0].p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(nkey.data(), ctx->d_ex_u[1].p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->t_sub_count.data(), ctx->d_ex_i[0].p, (size_t)ns * S * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->t_sub_break_sum.data(), ctx->d_ex_i[1].p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(sflags.data(), ctx->d_ex_flags.p, (size_t)ns, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->t_sub_sum_off.data(), ctx->d_fin_sum_off.p, (size_t)(ns + 1) * 4, cudaMemcpyDeviceToHost, st));
      if (ctx->total_sums)
        CK(cudaMemcpyAsync(ctx->t_sub_sums.data(), ctx->d_fin_sums.p, (size_t)ctx->total_sums * 8, cudaMemcpyDeviceToHost, st));
    }
    std::vector<int> key_arena(std::max<unsigned>(ctx->h_counters[CN_SUBKEY_USED], 1));
    CK(cudaMemcpyAsync(key_arena.data(), dc.sb_rec, (size_t)ctx->h_counters[CN_SUBKEY_USED] * 4, cudaMemcpyDeviceToHost, st));
    const int nsmall = S * (1 + 2 * dc.n_orf);
    ctx->t_small.assign(nsmall, 0);
    CK(cudaMemcpyAsync(ctx->t_small.data(), dc.small, (size_t)nsmall * 4, cudaMemcpyDeviceToHost, st));
    const size_t np = (size_t)ctx->n_pairs;
    std::vector<unsigned long long> pkeys(std::max<size_t>(np, 1));
    ctx->t_pair[4].assign(np, 0);
    if (np) {
      CK(cudaMemcpyAsync(pkeys.data(), ctx->d_pair_keys2.p, np * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->t_pair[4].data(), ctx->d_pair_cnt2.p, np * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    ctx->t_sub_first.assign(sfirst.begin(), sfirst.begin() + ns);
    ctx->t_sub_key_off.assign(ns + 1, 0);
    ctx->t_sub_key.clear();
    ctx->t_sub_flags.assign(sflags.begin(), sflags.begin() + ns);
    for (int r = 0; r < ns; r++) {
      ctx->t_sub_key_off[r] = (uint32_t)ctx->t_sub_key.size();
      ctx->t_sub_key.insert(ctx->t_sub_key.end(), key_arena.begin() + koff[r], key_arena.begin() + koff[r] + nkey[r]);
    }
    ctx->t_sub_key_off[ns] = (uint32_t)ctx->t_sub_key.size();
    if (!dc.track_sums) ctx->t_sub_sum_off.assign(ns + 1, 0);
    // cluster readCount[src] = sum over its sub-clusters (CigarCluster.merge :679-682)
    ctx->t_cl_read_count.assign((size_t)nc * S, 0);
    for (int c = 0; c < nc; c++)
      for (int k = cl_sub_off[c]; k < cl_sub_off[c + 1]; k++)
        for (int s = 0; s < S; s++) ctx->t_cl_read_count[(size_t)c * S + s] += ctx->t_sub_count[(size_t)k * S + s];
    // break-point pairs: sorted by key = (src, level, start, end) on the device
    for (int j = 0; j < 4; j++) ctx->t_pair[j].resize(np);
    for (size_t i = 0; i < np; i++) {
      const unsigned long long k = pkeys[i];
      ctx->t_pair[0][i] = (int)((k >> 59) & 0xF); ctx->t_pair[1][i] = (int)((k >> 58) & 1);
      ctx->t_pair[2][i] = (int)((k >> 29) & 0x1FFFFFFF); ctx->t_pair[3][i] = (int)(k & 0x1FFFFFFF);
    }
    R.cl_first_read = ctx->t_cl_first.data(); R.cl_key_off = ctx->t_cl_key_off.data(); R.cl_key_tok = ctx->t_cl_key_tok.data();
    R.cl_read_count = ctx->t_cl_read_count.data(); R.cl_sub_off = ctx->t_cl_sub_off.data();
    R.sub_first_read = ctx->t_sub_first.data(); R.sub_key_off = ctx->t_sub_key_off.data(); R.sub_key = ctx->t_sub_key.data();
    R.sub_count = ctx->t_sub_count.data(); R.sub_break_sum = ctx->t_sub_break_sum.data(); R.sub_sum_off = ctx->t_sub_sum_off.data();
    R.sub_sums = ctx->t_sub_sums.data(); R.sub_flags = ctx->t_sub_flags.data();
    R.total_counts = ctx->t_small.data(); R.spliced = ctx->t_small.data() + S; R.unspliced = ctx->t_small.data() + S + S * dc.n_orf;
    R.n_pairs = ctx->n_pairs;
    R.pair_src = ctx->t_pair[0].data(); R.pair_level = ctx->t_pair[1].data(); R.pair_start = ctx->t_pair[2].data();
    R.pair_end = ctx->t_pair[3].data(); R.pair_count = ctx->t_pair[4].data();
  }
  if ((what & NPT_FETCH_COVERAGE) && dc.record_depth) {
    const size_t per = (size_t)std::max(ctx->cov_rows, 1) * S * dc.stride;
    if (!ctx->r_cov.ensure(4 * per)) return fail(ctx, NPT_ENOMEM, "pinned coverage buffer (%zu bytes)", 4 * per * 4);
    if (ctx->cov_rows) CK(cudaMemcpyAsync(ctx->r_cov.p, ctx->d_cov_out.p, 4 * per * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    R.depth = ctx->r_cov.p; R.errors = ctx->r_cov.p + per; R.map_start = ctx->r_cov.p + 2 * per; R.map_end = ctx->r_cov.p + 3 * per;
  }
  *out = R;
  return NPT_OK;
}

int npt_release_results(npt_ctx* ctx) {
  if (!ctx) return NPT_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->s_comp);
  cudaStreamSynchronize(ctx->s_copy);
  free_batches(ctx);
  ctx->finalized = false;
  return NPT_OK;
}

int npt_device_view_get(npt_ctx* ctx, npt_device_view* out) {
  if (!ctx || !out) return NPT_EINVAL;
  if (!ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_device_view_get before npt_end_chrom");
  out->n_clusters = ctx->n_clusters;
  out->cov_stride = ctx->dc.stride;
  out->coverage = ctx->dc.record_depth ? ctx->d_cov_out.p : nullptr;
  out->coverage_elems = ctx->dc.record_depth ? 4ll * ctx->cov_rows * ctx->dc.S * ctx->dc.stride : 0;
  out->small_counts = ctx->d_small.p;
  out->small_elems = (int64_t)ctx->dc.S * (1 + 2 * ctx->dc.n_orf);
  out->n_subs = ctx->n_subs;
  out->n_pairs = (int32_t)ctx->n_pairs;
  out->sub_count = ctx->n_subs ? ctx->d_ex_i[0].p : nullptr;
  out->sub_count_elems = (int64_t)ctx->n_subs * ctx->dc.S;
  out->sub_break_sum = ctx->n_subs ? ctx->d_ex_i[1].p : nullptr;
  out->sub_break_sum_elems = ctx->n_subs;
  out->sub_sums = ctx->total_sums ? ctx->d_fin_sums.p : nullptr;
  out->sub_sums_elems = ctx->total_sums;
  out->pair_count = ctx->n_pairs ? ctx->d_pair_cnt2.p : nullptr;
  out->pair_count_elems = ctx->n_pairs;
  return NPT_OK;
}

// ------------------------------------------------------------------------------------------------ multi-rank key exchange
int npt_export_keys(npt_ctx* ctx, const int32_t** dev_buf, int64_t* n_words) {
  if (!ctx || !dev_buf || !n_words) return fail(ctx, NPT_EINVAL, "null argument");
  if (!ctx->in_chrom || ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_export_keys belongs between the last npt_submit and npt_end_chrom");
  int rc = npt_wait(ctx);
  if (rc) return rc;
  const DevCfg& dc = ctx->dc;
  unsigned hc[CN_COUNT];
  CK(cudaMemcpy(hc, ctx->d_counters.p, sizeof hc, cudaMemcpyDeviceToHost));
  const int cl_n = (int)std::min<unsigned>(hc[CN_CLUSTERS], (unsigned)dc.cl_cap), sb_n = (int)std::min<unsigned>(hc[CN_SUBS], (unsigned)dc.sb_cap);
  const int pr_n = dc.calc_breaks1 ? (int)std::min<unsigned>(hc[CN_PAIRS], (unsigned)dc.pr_cap) : 0;
  const long long clw = std::min<unsigned>(hc[CN_TOK_USED], dc.tok_cap), sbw = std::min<unsigned>(hc[CN_SUBKEY_USED], dc.sbkey_cap);
  int hdr[KX_HDR] = {0};
  long long o = KX_HDR;
  hdr[0] = KX_MAGIC; hdr[1] = cl_n; hdr[2] = (int)clw; hdr[3] = sb_n; hdr[4] = (int)sbw; hdr[5] = pr_n; hdr[6] = (int)(dc.cl_mask + 1); hdr[7] = dc.S;
  hdr[8] = (int)o; o += 4ll * cl_n;
  hdr[9] = (int)o; o += clw;
  hdr[10] = (int)o; o += 6ll * sb_n;
  hdr[11] = (int)o; o += sbw;
  hdr[12] = (int)o; o += 2ll * pr_n;
  if (o >= (1ll << 31)) return fail(ctx, NPT_ECAPACITY, "key exchange buffer too large");
  hdr[13] = (int)o;
  memcpy(ctx->kx_hdr, hdr, sizeof hdr);
  CK(ctx->d_kx.ensure((size_t)o));
  CK(ctx->d_kx_ctr.ensure(4));
  CK(cudaMemsetAsync(ctx->d_kx_ctr.p, 0, 16, ctx->s_comp));
  CK(cudaMemcpyAsync(ctx->d_kx.p, hdr, sizeof hdr, cudaMemcpyHostToDevice, ctx->s_comp));
  const int g = ctx->sm_count * 2;
  if (cl_n) kx_list_clusters_kernel<<<g, 256, 0, ctx->s_comp>>>(dc.cl, dc.cl_mask + 1, ctx->d_kx.p + hdr[8], ctx->d_kx_ctr.p, cl_n);
  if (clw) CK(cudaMemcpyAsync(ctx->d_kx.p + hdr[9], dc.cl_rec, (size_t)clw * 4, cudaMemcpyDeviceToDevice, ctx->s_comp));
  if (sb_n) kx_list_subs_kernel<<<g, 256, 0, ctx->s_comp>>>(dc.sb, dc.sb_mask + 1, dc.fit ? dc.sub_min0 : nullptr, ctx->d_kx.p + hdr[10], ctx->d_kx_ctr.p + 1, sb_n);
  if (sbw) CK(cudaMemcpyAsync(ctx->d_kx.p + hdr[11], dc.sb_rec, (size_t)sbw * 4, cudaMemcpyDeviceToDevice, ctx->s_comp));
  if (pr_n) kx_list_pairs_kernel<<<g, 256, 0, ctx->s_comp>>>(dc.pr, dc.pr_mask + 1, ctx->d_kx.p + hdr[12], ctx->d_kx_ctr.p + 2, pr_n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->s_comp));
  ctx->total_launches += 3;
  *dev_buf = ctx->d_kx.p;
  *n_words = o;
  return NPT_OK;
}

}  // extern "C"

// the part of npt_import_keys that runs once the buffer's header is known on the host: enqueues the insert kernels on the
// compute stream and does not wait for them
static int import_keys_enqueue(npt_ctx* ctx, const int32_t* dev_buf, int64_t n_words, const int* hdr) {
  const DevCfg& dc = ctx->dc;
  if (hdr[0] != KX_MAGIC || hdr[7] != dc.S) return fail(ctx, NPT_EINVAL, "not a key buffer of a context with the same configuration");
  const long long end = (long long)hdr[12] + 2ll * hdr[5];
  if (hdr[1] < 0 || hdr[3] < 0 || hdr[5] < 0 || hdr[6] < 1 || end > n_words) return fail(ctx, NPT_EINVAL, "inconsistent key buffer header");
  CK(ctx->d_kx_slotmap.ensure((size_t)hdr[6]));
  CK(cudaMemsetAsync(ctx->d_kx_slotmap.p, 0xFF, (size_t)hdr[6] * 4, ctx->s_comp));
  const int g = ctx->sm_count * 2;
  if (hdr[1]) import_clusters_kernel<<<g, 128, 0, ctx->s_comp>>>(dc, dev_buf + hdr[8], hdr[1], dev_buf + hdr[9], ctx->d_kx_slotmap.p);
  if (hdr[3]) import_subs_kernel<<<g, 128, 0, ctx->s_comp>>>(dc, dev_buf + hdr[10], hdr[3], dev_buf + hdr[11], ctx->d_kx_slotmap.p);
  if (hdr[5] && dc.calc_breaks1) import_pairs_kernel<<<g, 128, 0, ctx->s_comp>>>(dc, dev_buf + hdr[12], hdr[5]);
  CK(cudaGetLastError());
  ctx->total_launches += 3;
  return NPT_OK;
}

extern "C" {

int npt_import_keys(npt_ctx* ctx, const int32_t* dev_buf, int64_t n_words) {
  if (!ctx || !dev_buf) return fail(ctx, NPT_EINVAL, "null argument");
  if (!ctx->in_chrom || ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_import_keys belongs between the last npt_submit and npt_end_chrom");
  if (n_words < KX_HDR) return fail(ctx, NPT_EINVAL, "key buffer too short");
  CK(cudaSetDevice(ctx->device));
  int hdr[KX_HDR];
  CK(cudaMemcpy(hdr, dev_buf, sizeof hdr, cudaMemcpyDeviceToHost));
  int rc = import_keys_enqueue(ctx, dev_buf, n_words, hdr);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->s_comp));
  return NPT_OK;
}

void* npt_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void npt_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int npt_last_stage_ms(npt_ctx* ctx, float* out6) {
  if (!ctx || !out6) return NPT_EINVAL;
  for (int k = 0; k < 6; k++) out6[k] = ctx->stage_ms[k];
  return NPT_OK;
}

int npt_last_step_ms(npt_ctx* ctx, float* step_ms) {
  if (!ctx || !step_ms) return NPT_EINVAL;
  *step_ms = ctx->step_ms;
  return NPT_OK;
}

int npt_last_kernel_ms(npt_ctx* ctx, float* walk_ms, float* finalize_ms, int64_t* walk_launches, int64_t* total_launches) {
  if (!ctx) return NPT_EINVAL;
  if (walk_ms) *walk_ms = ctx->walk_ms;
  if (finalize_ms) *finalize_ms = ctx->fin_ms;
  if (walk_launches) *walk_launches = ctx->walk_launches;
  if (total_launches) *total_launches = ctx->total_launches;
  return NPT_OK;
}

}  // extern "C"

#include "comm.inl"
