// fastcov.cuh — device kernels around fastwalk.cuh:
//
//   fast_walk_kernel     one thread per read, rounds of [stage tiles -> produce tasks -> consume tasks -> flush entries]
//   cov_key_kernel       after clustering: sort key (coverage row of the cluster, source) per read + the few
//                        mapStart/mapEnd points of a read (CigarCluster.setStartEnd :360-365 and the break_p events of
//                        CigarCluster.add :476-486) credited with REDs
//   accumulate_kernel    per (cluster, source) group of the key-sorted reads: positional popcount of the reads' entries
//                        with bit-sliced carry-save adders, one thread per (reference-grid entry, read subset), private
//                        16-bit counters in shared memory, one RED per touched position per group at the end.
//
// Why no atomics per event: ncu of round 1 showed the coverage walk bound by the L2 atomic rate (126 REDs per read,
// 4.5e10 RED/s), and shared-memory atomics are no way out on this part (ATOMS ~2 cycles per lane).  A read's coverage is a
// set of per-position flags on the reference grid; summing flags over reads is positional popcount, which plain LOP3
// does 32 positions at a time (2 LOP3 per plane and word).
#pragma once
#include "fastwalk.cuh"

namespace npt {

#ifndef NPT_FW_WARPS
#define NPT_FW_WARPS 16
#endif
constexpr int FW_WARPS_MAX = NPT_FW_WARPS;
constexpr int FW_PW_TILE = 32 * FW_TILE_STRIDE;
__host__ __device__ constexpr int fw_warp_words(bool record) { return FW_PW_TILE + (record ? (FW_RING + FW_RING_PAD) * 32 : 0); }

__device__ __forceinline__ void fw_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void fw_cp16z(uint32_t dst, const void* src, int n) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void fw_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool REF_SMEM, bool RECORD>
__global__ void __launch_bounds__(FW_WARPS_MAX * 32, 1) fast_walk_kernel(const DevCfg c, const BatchDev b) {
  extern __shared__ __align__(16) uint32_t fw_smem[];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const int refw = REF_SMEM ? ((c.refg_nwords + 3) & ~3) : 0;
  uint32_t* s_ref = fw_smem;
  uint32_t* wb = fw_smem + refw + warp * fw_warp_words(RECORD);
  uint32_t* tileT = wb;                                                 // [32 lanes][CIGAR tile | base tile]
  uint32_t* ringT = tileT + FW_PW_TILE;
  if (REF_SMEM)
    for (int i = threadIdx.x; i < c.refg_nwords; i += blockDim.x) s_ref[i] = c.refg_words[i];
  if (RECORD)
    for (int i = lane; i < (FW_RING + FW_RING_PAD) * 32; i += 32) ringT[i] = 0;
  __syncthreads();
  const uint32_t* refg = REF_SMEM ? s_ref : c.refg_words;
  const uint32_t tileS = (uint32_t)__cvta_generic_to_shared(tileT);
  const uint32_t* cigL = tileT + lane * FW_TILE_STRIDE;
  const uint32_t* seqL = cigL + FW_CIG_TILE;
  // the lane's coverage ring (shared memory, lane-interleaved) or, without depth recording, its break-list scratch (global)
  uint32_t* auxL = RECORD ? ringT + lane : c.fw_scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * FW_BRK_BUF;
  const int auxS = RECORD ? 32 : 1;
  unsigned* cursor = b.bctr + 5;
  const unsigned cig_words = (unsigned)b.cigar_words;                   // cigar_off is 32 bits
  const unsigned long long seq_bytes = (unsigned long long)b.seq_bytes;

  FwLane L;
  L.flags = 0; L.tb = -1; L.so = 0; L.readPos = 0; L.ck = 0; L.cs = 0; L.r = 0; L.cend = 0; L.mRem = 0; L.refPos = 0; L.flushE = 0; L.tnib0 = 0; L.st_r = 0;
  for (;;) {
    // ---- (0) lanes without a read take the next ones from the cursor (one atomic per warp and round)
    const bool want = (L.flags & (FWF_ACTIVE | FWF_DONE)) == 0;
    const unsigned wm = __ballot_sync(FULL, want);
    if (wm) {
      unsigned base = 0;
      const int leader = __ffs(wm) - 1;
      if ((int)lane == leader) base = atomicAdd(cursor, (unsigned)__popc(wm));
      base = __shfl_sync(FULL, base, leader);
      if (want) {
        const long long r = (long long)base + __popc(wm & ((1u << lane) - 1u));
        if (r < b.n) fw_start_read(L, b, r); else L.flags = FWF_DONE;
      }
    }
    if (__all_sync(FULL, (L.flags & (FWF_ACTIVE | FWF_DONE)) == FWF_DONE)) break;
    // ---- (1) stage this round's CIGAR words and bases.  Every lane says what it needs (a shuffle away), the warp copies all
    // tiles together with 16-byte cp.async so that neighbouring lanes fetch neighbouring chunks of one tile (a lane copying only
    // its own tile costs one L1 wavefront per lane and copy)
    const bool live = (L.flags & (FWF_ACTIVE | FWF_SLOW)) == FWF_ACTIVE;
    {
      const unsigned ncs = L.ck & ~3u;
      const bool needC = live && ncs != L.cs && L.ck < L.cend;
      if (needC) L.cs = ncs;
      const unsigned myC = needC ? ncs + 1u : 0u;                        // first CIGAR word to stage + 1, 0 = none
      const long long ntb = (long long)((L.so + (unsigned long long)(L.readPos >> 1)) & ~15ull);
      const bool needS = live && ntb != L.tb;
      if (needS) { L.tb = ntb; L.tnib0 = (int)(2 * (ntb - (long long)L.so)); }
      const unsigned myS = needS ? (unsigned)((unsigned long long)ntb >> 4) + 1u : 0u;   // first 16-byte chunk to stage + 1, 0 = none
#pragma unroll
      for (int it = 0; it < FW_CIG_TILE / 4; it++) {
        const unsigned id = it * 32 + lane, owner = id / (FW_CIG_TILE / 4), ch = id - owner * (FW_CIG_TILE / 4);
        const unsigned s1 = __shfl_sync(FULL, myC, owner);
        if (s1) {
          const unsigned w0 = s1 - 1u + 4u * ch;
          const uint32_t dst = tileS + (owner * FW_TILE_STRIDE + 4 * ch) * 4;
          if (w0 + 4u <= cig_words) fw_cp16(dst, b.cigar + w0);
          else fw_cp16z(dst, w0 < cig_words ? (const void*)(b.cigar + w0) : (const void*)b.cigar, w0 < cig_words ? (int)(cig_words - w0) * 4 : 0);
        }
      }
#pragma unroll
      for (int it = 0; it < FW_SEQ_TILE / 4; it++) {
        const unsigned id = it * 32 + lane, owner = id / (FW_SEQ_TILE / 4), ch = id - owner * (FW_SEQ_TILE / 4);
        const unsigned s1 = __shfl_sync(FULL, myS, owner);
        if (s1) {
          const unsigned long long a0 = ((unsigned long long)(s1 - 1u) + ch) << 4;
          const uint32_t dst = tileS + (owner * FW_TILE_STRIDE + FW_CIG_TILE + 4 * ch) * 4;
          if (a0 + 16ull <= seq_bytes) fw_cp16(dst, b.seq4 + a0);
          else fw_cp16z(dst, a0 < seq_bytes ? (const void*)(b.seq4 + a0) : (const void*)b.seq4, a0 < seq_bytes ? (int)(seq_bytes - a0) : 0);
        }
      }
      fw_cp_wait();
      __syncwarp();
    }
    // ---- (2) FW_R steps: [one element that is not M] + [start of the M behind it] + [<= 16 positions of the current M]
    const int ringLimit = (L.flushE << 5) + FW_RING * 8 - 2 * FW_PIECE - 2;
#pragma unroll 1
    for (int it = 0; it < FW_R; it++) fw_step<RECORD>(L, c, b, cigL, seqL, refg, auxL, auxS, live, ringLimit);
    __syncwarp();
    // ---- (3) N elements, completed entries, closed blocks, finished reads
    fw_round_end<RECORD>(L, c, b, cigL, auxL, auxS);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ keys + start/end points
// key = (coverage row of the read's cluster * S + source) << ACC_POS_BITS | coarse start position; reads that are not credited
// by accumulate_kernel (walked by the serial path, bad records, unclustered) get the sentinel row cov_cap * S and sort to the
// end.  The position bits make neighbours in the sorted order cover neighbouring parts of the reference, so that the reads of
// one tile of a wide cluster (genomic fragments anywhere in 20 kb) touch few reference-grid entries.
constexpr int ACC_POS_BITS = 10;
// +1 at *p for every lane that is here with the same address: one RED per distinct address of the converged lanes.  Start,
// end and junction positions of a batch fall on a few dozen addresses (ncu r02b: cov_key_kernel waits on REDs to hot lines).
__device__ __forceinline__ void red_inc_aggregated(int* p) {
#ifdef NPT_NO_RED_AGG
  atomicAdd(p, 1);
#else
  const unsigned conv = __activemask();
  const unsigned grp = __match_any_sync(conv, (unsigned long long)p);
  if ((__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(p, __popc(grp));
#endif
}
__global__ void cov_key_kernel(const DevCfg c, const BatchDev b, unsigned* keys, unsigned* vals, int pos_shift) {
  if (__ldcg(b.bctr + 1)) return;  // break lists incomplete, nothing was clustered: the host re-runs this batch
  const long long arr_stride = (long long)c.cov_cap * c.S * c.stride;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += (long long)gridDim.x * blockDim.x) {
    const unsigned rf = b.rflags[r];
    const int slot = b.cluster[r];
    unsigned key = ((unsigned)c.cov_cap * (unsigned)c.S) << ACC_POS_BITS;
    if (!(rf & 0x30u) && slot >= 0) {
      const int prov = __ldcg(c.cl_rec + (unsigned)(__ldcg(&c.cl[slot].word) & 0xffffffffu));
      const int src = b.src[r];
      if (prov >= c.cov_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_COV_FULL);
      else {
        const unsigned first = __ldg(&b.fhdrA[r].y) & 0xFFFFu;   // first reference-grid entry of the read's first block
        key = (((unsigned)prov * (unsigned)c.S + (unsigned)src) << ACC_POS_BITS) | min(first >> pos_shift, (1u << ACC_POS_BITS) - 1u);
        int* cov = c.cov + ((long long)prov * c.S + src) * c.stride;
        int* mstart = cov + 2 * arr_stride;
        int* mend = cov + 3 * arr_stride;
        // CigarCluster.setStartEnd :360-365 with the (trimmed) start / end of the read
        const int sp = b.start_pos[r], ep = b.end_pos[r];
        if (sp >= 0 && sp < c.stride) red_inc_aggregated(mstart + sp);
        if (ep >= 0 && ep < c.stride) red_inc_aggregated(mend + ep);
        const unsigned meta = __ldg(b.fmeta + r);
        const int njunc = (meta >> 8) & 0xFF, nextra = (meta >> 16) & 0xFF, ndbl = (meta >> 24) & 0xFF;
        for (int d = 0; d < ndbl; d++) {   // positions emitted by two D elements
          const unsigned q = __ldg(b.frare + r * FW_RARE_WORDS + d);
          atomicAdd(cov + 8 * arr_stride + q, 1);
          atomicAdd(cov + arr_stride + q, 1);
        }
        if (njunc | nextra) {
          const uint4 h2 = __ldg((const uint4*)(b.frare + r * FW_RARE_WORDS + 4));
          const uint4 h3 = __ldg((const uint4*)(b.frare + r * FW_RARE_WORDS + 8));
          const uint4 h4 = nextra > 2 ? __ldg((const uint4*)(b.frare + r * FW_RARE_WORDS + 12)) : make_uint4(0u, 0u, 0u, 0u);
          if (njunc > 0) { red_inc_aggregated(mend + h2.x); red_inc_aggregated(mstart + h2.y); }
          if (njunc > 1) { red_inc_aggregated(mend + h2.z); red_inc_aggregated(mstart + h2.w); }
          for (int e = 0; e < nextra; e++) {
            const unsigned a = e == 0 ? h3.x : (e == 1 ? h3.z : (e == 2 ? h4.x : h4.z)), p = e == 0 ? h3.y : (e == 1 ? h3.w : (e == 2 ? h4.y : h4.w));
            if (a & 0x80000000u) {          // a position emitted by two D elements
              const unsigned q = a & 0x7FFFFFFFu;
              atomicAdd(cov + 8 * arr_stride + q, 1);   // depth
              atomicAdd(cov + arr_stride + q, 1);       // errors
            } else {                        // emission more than break_thresh after prev that was not a match
              atomicAdd(mstart + a, 1); atomicAdd(mend + p, 1);
            }
          }
        }
      }
    }
    keys[r] = key; vals[r] = (unsigned)r;
  }
}

// ------------------------------------------------------------------------------------------------ positional popcount
// Per (cluster, source) run of the sorted reads: thread = (reference-grid entry s, read subset q).  A thread adds the entries
// of its reads into 8 bit planes per entry word (carry-save adders: a 3:2 compressor is two LOP3), 252 reads before the planes
// could overflow.  The subsets are then summed plane-wise (bit-sliced ripple adders, still no per-position work) by four
// threads per entry, and only that one total per entry word is expanded into per-position counts and credited with one RED
// per touched position and row.
constexpr int ACC_THREADS = 768;
constexpr int ACC_TILE = 2048;        // key-sorted reads per block tile
constexpr int ACC_MAX_ENT = 8192;     // reference-grid entries (32 positions each) a chromosome may have on this path
constexpr int ACC_PLANES = 8;         // per-thread planes: 255 reads
constexpr int ACC_SUM_PLANES = 12;    // per-entry total of a tile: 2048 reads
constexpr int ACC_EPOCH = 252;        // reads per thread between two reductions (63 steps of four)

struct AccSmem {
  uint32_t planes[ACC_PLANES * 4 * ACC_THREADS];   // [plane][word][thread]
  uint4 hdr[ACC_TILE];                             // the tile's read headers {first entry, block 0, 1, 2}
  uint32_t head[ACC_TILE / 32];
  uint32_t act[ACC_MAX_ENT / 32];
  uint32_t actpre[ACC_MAX_ENT / 32 + 1];
  uint16_t word_of[ACC_MAX_ENT];
};

// four words into the planes at once
__device__ __forceinline__ void acc_add4(uint32_t (&p)[ACC_PLANES], uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t s1 = p[0] ^ a ^ b, c1 = (p[0] & a) | (p[0] & b) | (a & b);
  const uint32_t s2 = s1 ^ c ^ d, c2 = (s1 & c) | (s1 & d) | (c & d);
  p[0] = s2;
  const uint32_t t = p[1] ^ c1 ^ c2;
  uint32_t x = (p[1] & c1) | (p[1] & c2) | (c1 & c2);   // weight 4
  p[1] = t;
#pragma unroll
  for (int k = 2; k < ACC_PLANES; k++) { const uint32_t cy = p[k] & x; p[k] ^= x; x = cy; }
}
// one entry credited directly, position by position (the rare second entry of a read on the same reference-grid entry)
__device__ __noinline__ void acc_credit_entry(const uint4 x, unsigned myw, int* row_dep, int* row_err, int stride) {
  for (int w = 0; w < 4; w++) {
    const uint32_t v = w == 0 ? x.x : (w == 1 ? x.y : (w == 2 ? x.z : x.w));
    for (int i = 0; i < 8; i++) {
      const uint32_t nib = (v >> (28 - 4 * i)) & 0xFu;
      const int pos = (int)myw * 32 + 8 * w + i;
      if (nib && pos < stride) {
        const int cov = nib & 1u, qrk = (nib >> 1) & 1u, mis = (nib >> 3) & 1u;
        if (cov + qrk) atomicAdd(row_dep + pos, cov + qrk);
        if (mis + qrk) atomicAdd(row_err + pos, mis + qrk);
      }
    }
  }
}

// which of a read's entries lies on reference-grid entry `myw`: index into the read's entries, or -1.  `more` = a later block of
// the read holds an entry for the same grid entry too (an N element shorter than 32 positions): the caller adds those one by one
__device__ __forceinline__ int acc_entry_index(const uint4 h, unsigned myw, bool& more) {
  const unsigned c0 = h.y >> 16, c1 = h.z >> 16, c2 = h.w >> 16;
  const unsigned d0 = myw - (h.y & 0xFFFFu), d1 = myw - (h.z & 0xFFFFu), d2 = myw - (h.w & 0xFFFFu);
  const bool in0 = d0 < c0, in1 = d1 < c1, in2 = d2 < c2;
  more = (in0 & (in1 | in2)) | (in1 & in2);
  const unsigned i12 = in1 ? c0 + d1 : c0 + c1 + d2;
  return (in0 | in1 | in2) ? (int)(in0 ? d0 : i12) : -1;
}

__global__ void __launch_bounds__(ACC_THREADS, 1) accumulate_kernel(const DevCfg c, const BatchDev b, const unsigned* keys, const unsigned* vals,
                                                                    unsigned* tile_cursor) {
  extern __shared__ __align__(16) unsigned char acc_raw[];
  if (__ldcg(b.bctr + 1)) return;  // see cov_key_kernel
  AccSmem& sm = *reinterpret_cast<AccSmem*>(acc_raw);
  __shared__ int s_tile, s_wact;
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned invalid = (unsigned)c.cov_cap * (unsigned)c.S;   // row part of the key of reads that are not credited here
  const long long arr_stride = (long long)c.cov_cap * c.S * c.stride;
  const int n_ent = (c.stride + 31) / 32 + 1;
  const int n_actw = (n_ent + 31) / 32;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (;;) {
    if (tid == 0) s_tile = (int)atomicAdd(tile_cursor, 1u);
    __syncthreads();
    const long long n_tiles = (b.n + ACC_TILE - 1) / ACC_TILE;
    // the tiles with many small groups (late-created clusters sort last) cost the most per read: hand them out first
    const long long t0 = (n_tiles - 1 - (long long)s_tile) * ACC_TILE;
    __syncthreads();
    if (t0 < 0) break;
    const int tn = (int)min((long long)ACC_TILE, b.n - t0);
    if ((__ldg(keys + t0) >> ACC_POS_BITS) == invalid) continue;   // sorted: this tile holds only reads that are not credited here
    // ---- the tile: read headers into shared memory, head flags of the runs of equal (row, source)
    for (int i = tid; i < ACC_TILE; i += ACC_THREADS) {
      bool head = false;
      if (i < tn) {
        const unsigned k = __ldg(keys + t0 + i) >> ACC_POS_BITS;
        sm.hdr[i] = k == invalid ? zero4 : __ldg(b.fhdrA + __ldg(vals + t0 + i));
        head = (i == 0) || (k != (__ldg(keys + t0 + i - 1) >> ACC_POS_BITS));
      }
      const unsigned hb = __ballot_sync(0xffffffffu, head);
      if (lane == 0) sm.head[i >> 5] = hb;
    }
    __syncthreads();
    int i0 = 0;
    while (i0 < tn) {
      // ---- next run [i0, i1)
      int i1 = tn;
      {
        int w = (i0 + 1) >> 5;
        unsigned bits = (i0 + 1) < ACC_TILE ? (sm.head[w] & (0xffffffffu << ((i0 + 1) & 31))) : 0u;
        const int nw = (tn + 31) >> 5;
        while (!bits && ++w < nw) bits = sm.head[w];
        if (bits && w < nw) i1 = min(tn, w * 32 + __ffs(bits) - 1);
      }
      const unsigned key = __ldg(keys + t0 + i0) >> ACC_POS_BITS;
      if (key == invalid) break;
      const int rl = i1 - i0;
      const uint4* hdr = sm.hdr + i0;
      int* row_err = c.cov + arr_stride + (long long)key * c.stride;
      int* row_dep = c.cov + 8 * arr_stride + (long long)key * c.stride;
      // ---- (A) reference-grid entries touched by the run
      for (int i = tid; i < n_actw; i += ACC_THREADS) sm.act[i] = 0;
      __syncthreads();
      for (int i = tid; i < rl; i += ACC_THREADS) {
        const uint4 h0 = hdr[i];
#pragma unroll
        for (int k = 0; k < FW_MAX_BLOCKS; k++) {
          const unsigned bk = k == 0 ? h0.y : (k == 1 ? h0.z : h0.w);
          const unsigned f = bk & 0xFFFFu, cn = bk >> 16;
          if (cn) {
            const unsigned lo = f, hi = f + cn - 1;   // entries [lo, hi]
            for (unsigned w = lo >> 5; w <= (hi >> 5) && w < (unsigned)n_actw; w++) {
              unsigned m = 0xffffffffu;
              if (w == (lo >> 5)) m &= 0xffffffffu << (lo & 31);
              if (w == (hi >> 5)) m &= 0xffffffffu >> (31 - (hi & 31));
              if ((sm.act[w] & m) != m) atomicOr(&sm.act[w], m);
            }
          }
        }
      }
      __syncthreads();
      if (tid < 32) {   // exclusive prefix of the popcounts (n_actw <= 256)
        int carry = 0;
        for (int base = 0; base < n_actw; base += 32) {
          const int w = base + lane;
          const int pc = w < n_actw ? __popc(sm.act[w]) : 0;
          int inc = pc;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
          if (w < n_actw) sm.actpre[w] = carry + inc - pc;
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_wact = carry;
      }
      __syncthreads();
      for (int w = tid; w < n_actw; w += ACC_THREADS) {
        unsigned bits = sm.act[w];
        int slot = sm.actpre[w];
        while (bits) { const int bpos = __ffs(bits) - 1; bits &= bits - 1; sm.word_of[slot++] = (uint16_t)(w * 32 + bpos); }
      }
      __syncthreads();
      const int wact = s_wact;
      // ---- (B) passes of up to ACC_THREADS entries; thread = (entry s, read subset q)
      for (int slot0 = 0; slot0 < wact; slot0 += ACC_THREADS) {
        const int W = min(ACC_THREADS, wact - slot0);
        const int Wpad = (W + 31) & ~31;
        // subsets: the main loop costs ~ rl / Q per thread, the reduction ~ Q per entry word: balanced near sqrt(rl)
        int Q = 1;
        while (Q * Q * 2 < rl && (Q + 1) * Wpad <= ACC_THREADS) Q++;
        const int q = tid / Wpad, s = tid - q * Wpad;
        const bool live = q < Q && s < W;
        const unsigned myw = (s < W) ? sm.word_of[slot0 + s] : 0u;
        for (int e0 = 0; e0 < rl; e0 += ACC_EPOCH * Q) {
          const int e1 = min(rl, e0 + ACC_EPOCH * Q);
          uint32_t P0[ACC_PLANES], P1[ACC_PLANES], P2[ACC_PLANES], P3[ACC_PLANES];
#pragma unroll
          for (int k = 0; k < ACC_PLANES; k++) { P0[k] = 0; P1[k] = 0; P2[k] = 0; P3[k] = 0; }
          if (live) {
            // four reads per step: headers come from shared memory (one broadcast load each), so the four entry loads are in flight together
            for (int i = e0 + q; i < e1; i += 4 * Q) {
              uint4 e[4];
              bool extra = false;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                e[j] = zero4;
                const int ii = min(i + j * Q, e1 - 1);
                const bool ok = i + j * Q < e1;
                const uint4 h = hdr[ii];
                bool more;
                const int idx = acc_entry_index(h, myw, more);
                if (ok & (idx >= 0)) e[j] = __ldg(b.fent + (size_t)h.x + idx);
                extra |= ok & more;
              }
              acc_add4(P0, e[0].x, e[1].x, e[2].x, e[3].x);
              acc_add4(P1, e[0].y, e[1].y, e[2].y, e[3].y);
              acc_add4(P2, e[0].z, e[1].z, e[2].z, e[3].z);
              acc_add4(P3, e[0].w, e[1].w, e[2].w, e[3].w);
              if (extra) {   // rare: the later blocks' entries for the same grid entry are credited directly
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                  const int ii = i + j * Q;
                  if (ii >= e1) break;
                  const uint4 h = hdr[ii];
                  const unsigned c0 = h.y >> 16, c1 = h.z >> 16, c2 = h.w >> 16;
                  const unsigned d0 = myw - (h.y & 0xFFFFu), d1 = myw - (h.z & 0xFFFFu), d2 = myw - (h.w & 0xFFFFu);
                  const bool in0 = d0 < c0, in1 = d1 < c1, in2 = d2 < c2;
                  if (in0 && in1) acc_credit_entry(__ldg(b.fent + (size_t)h.x + c0 + d1), myw, row_dep, row_err, c.stride);
                  if ((in0 || in1) && in2) acc_credit_entry(__ldg(b.fent + (size_t)h.x + c0 + c1 + d2), myw, row_dep, row_err, c.stride);
                }
              }
            }
#pragma unroll
            for (int k = 0; k < ACC_PLANES; k++) {
              sm.planes[(k * 4 + 0) * ACC_THREADS + tid] = P0[k]; sm.planes[(k * 4 + 1) * ACC_THREADS + tid] = P1[k];
              sm.planes[(k * 4 + 2) * ACC_THREADS + tid] = P2[k]; sm.planes[(k * 4 + 3) * ACC_THREADS + tid] = P3[k];
            }
          }
          __syncthreads();
          // ---- (C) per entry word: plane-wise sum of the subsets, then per-position counts and the REDs.
          // Thread (s, w = q) for q < 4; with fewer than four subsets the remaining words are taken in further rounds.
          for (int w = q; w < 4; w += Q) {
            if (!live) break;
            uint32_t T[ACC_SUM_PLANES];
#pragma unroll
            for (int k = 0; k < ACC_SUM_PLANES; k++) T[k] = 0;
            for (int qq = 0; qq < Q; qq++) {
              uint32_t cy = 0;
#pragma unroll
              for (int k = 0; k < ACC_PLANES; k++) {   // full adders: T[k] + plane k of subset qq + carry
                const uint32_t x = sm.planes[(k * 4 + w) * ACC_THREADS + qq * Wpad + s];
                const uint32_t sum = T[k] ^ x ^ cy;
                cy = (T[k] & x) | (T[k] & cy) | (x & cy);
                T[k] = sum;
              }
#pragma unroll
              for (int k = ACC_PLANES; k < ACC_SUM_PLANES; k++) { const uint32_t t2 = T[k] & cy; T[k] ^= cy; cy = t2; }
            }
            // position 8w + i of the entry = nibble i of the word counted from the top: flag bit fb of it is bit 28 - 4i + fb
#pragma unroll
            for (int i = 0; i < 8; i++) {
              int cov = 0, qrk = 0, mis = 0;
#pragma unroll
              for (int k = 0; k < ACC_SUM_PLANES; k++) {
                const uint32_t nib = (T[k] >> (28 - 4 * i)) & 0xFu;
                cov += (int)(nib & 1u) << k; qrk += (int)((nib >> 1) & 1u) << k; mis += (int)((nib >> 3) & 1u) << k;
              }
              const int pos = (int)myw * 32 + 8 * w + i;
              if (pos < c.stride) {
                if (cov + qrk) atomicAdd(row_dep + pos, cov + qrk);
                if (mis + qrk) atomicAdd(row_err + pos, mis + qrk);
              }
            }
          }
          __syncthreads();
        }
      }
      i0 = i1;
    }
    __syncthreads();
  }
}

}  // namespace npt
