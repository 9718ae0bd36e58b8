// kernels.cuh — the hot path as three small kernels, one THREAD per read with lane-level dynamic scheduling.
//
//   walk_kernel<0>   CIGAR walk + 4-bit base compare, 32 bases per loop step (four 8-base XORs) from lane-private tiles the
//                    owning lane stages with cp.async -> break list, read offsets, error sums
//                    (IdentityProfile1.processCigar + CigarCluster.add, J/cluster/IdentityProfile1.java:482-585,
//                    J/cluster/CigarCluster.java:442-498, incl. the D/=/X "advance, then emit" quirk)
//   cluster_kernel   break list -> key tokens / break-point pairs / ORF counters / type -> cluster and sub-cluster hash
//                    tables (IdentityProfile1.processRefPositions :149-356, CigarClusters.matchCluster :75-102,
//                    CigarCluster.merge :633-698).  Records are written before the CAS that publishes them: no waiting.
//   walk_kernel<1>   second walk (the cluster must be known before a position can be credited) that sends the read's
//                    depth runs as difference-array events (+1 at run start, -1 after run end), its short deletions as one
//                    histogram point each and its error / mapStart / mapEnd points straight to the rows of its cluster with
//                    RED.ADD; at end of chromosome the histograms are folded in and a segmented scan turns difference rows
//                    into depth.  Bound by the L2 atomic rate (126 REDs per read), not by the walk (profiles/NOTES.md).
//   walk_kernel<2>   store-all pass for the few reads whose break list does not fit the inline slot.
//
// Why thread-per-read: the first design of this file was warp-per-read (32 ops per step, warp scans for offsets, ballot
// compaction, shared-memory event staging).  ncu (profiles/r01a..r01c) showed it instruction- and I-cache-bound at 1 % of
// the HBM roofline: ~7 000 warp-instructions per read, `no_instruction` the top stall, DRAM 1 % busy.  A read is a strictly
// sequential state machine with ~8-base granularity; giving each lane its own read removes every scan, shuffle and ballot
// (7 000 -> 1 700 warp-instructions per read today) and lets each kernel fit the instruction cache.  Lanes fetch the next read from
// an atomic cursor as soon as they finish one, so a warp stays busy whatever the read lengths are.
#pragma once
#include "npt_device.cuh"

namespace npt {

// coverage updates of the walk; the experiment switch keeps every address computation and loop but issues no RED
// (profiles/NOTES.md: how much of walk_kernel<1> is the atomics themselves)
#ifdef NPT_EXPERIMENT_NORED
#define COV_RED(addr, v) do { int* a__ = (addr); if (a__ == nullptr) atomicAdd(a__, (v)); } while (0)
#else
#define COV_RED(addr, v) atomicAdd((addr), (v))
#endif
// the few updates per read that land on a handful of addresses (run boundaries at the leader / junction / 3' end, read
// start and end): second experiment switch
#ifdef NPT_EXPERIMENT_NOHOTRED
#define COV_RED_HOT(addr, v) do { int* a__ = (addr); if (a__ == nullptr) atomicAdd(a__, (v)); } while (0)
#else
#define COV_RED_HOT(addr, v) COV_RED(addr, v)
#endif

constexpr int BRK_INLINE = 6;     // breaks stored in the per-read inline slot
constexpr int WALK_THREADS = 256;
constexpr int GRAB = 1;           // reads a lane takes from the cursor at a time
constexpr int MAX_BREAKS = 1 << 20;
constexpr int TILE_W = 20;        // words per lane tile: five 16-byte chunks fetched by the owning lane with cp.async
constexpr int STEP_WINS = 4;      // 8-base windows compared per loop iteration
constexpr int STEP_BASES = 8 * STEP_WINS;
constexpr int DEL_HIST = 4;        // deletion lengths recorded in the per-length histogram rows (coverage arrays 4..7)
constexpr int COV_ARRAYS = 4 + DEL_HIST + 1;  // + absolute depth credited by accumulate_kernel (fastcov.cuh)
constexpr int TILE_STRIDE = 20;   // 80 bytes: keeps every lane tile 16-byte aligned for cp.async

template <typename T>
__device__ __forceinline__ T ldcg(const T* p) { return __ldcg(p); }

// 8 reference bases starting at 0-based index `nib` (words hold the first base in bits 31..28)
__device__ __forceinline__ uint32_t ref_win8(const uint32_t* w, uint32_t nib) {
  const uint32_t k = nib >> 3, o = (nib & 7) << 2;
  return __funnelshift_l(w[k + 1], w[k], o);
}
// 16-byte global -> shared copies of the owning lane (LDGSTS, L2 only); n < 16 zero-fills the rest without reading it
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, int n) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// ------------------------------------------------------------------------------------------------ the walk
// MODE 0: first pass (breaks into the inline slot, counts, read offsets).  MODE 1: coverage pass (REDs into the cluster's
// rows).  MODE 2: store-all pass for the reads whose break list did not fit the inline slot.
template <int MODE, bool REF_SMEM>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(const DevCfg c, const BatchDev b) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  if (MODE == 1 && ldcg(b.bctr + 1)) return;  // break lists incomplete (overflow area too small): the host re-runs this batch
  if (MODE == 1 && b.fast == 2 && ldcg(b.bctr + 8)) return;  // ... of the serial walker's reads only: the host re-runs their chain
  // behind the fast walk only the handed-over reads are walked here: blocks that would find the list empty leave at once
  // (the grid is sized for a whole batch; loading the reference into shared memory costs more than walking a few reads)
  if (b.fast && (unsigned long long)blockIdx.x * WALK_THREADS >= (unsigned long long)ldcg(b.bctr + 7)) return;
  // shared memory: [packed reference (if it fits), padded to 16 bytes] [per-warp CIGAR tiles 32 x 20] [per-warp base tiles 32 x 20]
  uint32_t* s_ref = s_dyn;
  const int ref_words = REF_SMEM ? ((c.ref_nwords + 3) & ~3) : 0;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* cigTile = s_dyn + ref_words + warp * (2 * 32 * TILE_STRIDE);
  uint32_t* seqTile = cigTile + 32 * TILE_STRIDE;
  uint32_t* s_vm = s_dyn + ref_words + (WALK_THREADS / 32) * (2 * 32 * TILE_STRIDE);  // valid-base masks by step length
  uint32_t* myCig = cigTile + lane * TILE_STRIDE;
  uint32_t* mySeq = seqTile + lane * TILE_STRIDE;
  if (REF_SMEM)
    for (int i = threadIdx.x; i < c.ref_nwords; i += blockDim.x) s_ref[i] = c.ref_words[i];
  if (threadIdx.x <= STEP_BASES) {  // s_vm[n]: bit 4j+q set iff base 8q + 7 - j < n (the layout of the step masks below)
    uint32_t m = 0;
    for (int i = 0; i < (int)threadIdx.x; i++) m |= 1u << (((7 - (i & 7)) << 2) + (i >> 3));
    s_vm[threadIdx.x] = m;
  }
  __syncthreads();
  const uint32_t* refw = REF_SMEM ? s_ref : c.ref_words;
  const uint32_t* seqw = (const uint32_t*)b.seq4;
  const long long seq_words = (b.seq_bytes + 3) >> 2;  // tiles never read past the arrays: the last chunks are clamped
  const long long cig_words = (long long)b.cigar_words;
  const uint32_t myCigS = (uint32_t)__cvta_generic_to_shared(myCig), mySeqS = (uint32_t)__cvta_generic_to_shared(mySeq);
  const int T = c.T, G = c.G;
  const bool per_base = T < STEP_BASES;  // a break could hide inside one step: always take the exact per-base path
  unsigned* cursor = b.bctr + 2 + MODE;  // one read cursor per pass
  const long long arr_stride = (long long)c.cov_cap * c.S * c.stride;

  long long r = 0;
  bool active = false;
  // state of the read being walked
  uint32_t nib0 = 0;
  int alnStart = 0, refPos = 0, readPos = 0, prev = 0, nb = 0, maps = 0, errs = 0;
  // prev is kept lazily: while pmat != 0 the last match is the last set base of pmat, a step mask at position ppos0; it is
  // decoded only when a comparison against T could come out true (junctions), which needs the exact value
  uint32_t pmat = 0; int ppos0 = 0;
  int mStart0 = 0, mRpos = 0, mN = 0, mOff = 0;  // M op being windowed: windows remain while mOff < mN
  int st_r = 0, maxEnd = 0, eBs = 0, eRs = 0;
  int nbases = 0;           // MODE 0: bases this record holds (2 per byte: the pad nibble of an odd-length read counts)
  bool stFound = false;
  int pendingEnd = -1;      // MODE 1: exclusive end of the last depth run, not yet written
  int* cov = nullptr;       // MODE 1: row [arr 0][cluster][src][0]
  int* bdst = nullptr;      // MODE 0/2: where this read's breaks go
  int slow_windows = 0;

  auto last_match = [](uint32_t m) -> int {  // base index of the last set base of a step mask, -1 if none
    int l = -1;
#pragma unroll
    for (int q = 0; q < STEP_WINS; q++) { const uint32_t mq = (m >> q) & 0x11111111u; if (mq) l = 8 * q + 7 - ((__ffs(mq) - 1) >> 2); }
    return l;
  };
  auto sync_prev = [&]() { if (pmat) { prev = ppos0 + last_match(pmat); pmat = 0; } };
  // CigarCluster.add for one emitted position (exact, used only near junctions or when T < 8)
  auto add = [&](int p, bool match) {
    const bool break_p = prev < 0 || p - prev > T;
    if (MODE == 1 && break_p && prev > 0) { COV_RED(cov + 2 * arr_stride + p, 1); COV_RED(cov + 3 * arr_stride + prev, 1); }
    if (match) {
      if (break_p) {
        if (prev > 0) { if (MODE == 2 || (MODE == 0 && nb < BRK_INLINE)) bdst[nb] = prev; nb++; }
        if (MODE == 2 || (MODE == 0 && nb < BRK_INLINE)) bdst[nb] = p;
        nb++;
      }
      prev = p;
    }
  };
  // depth run [base, base+n): difference events, cancelling "end of previous run == start of this run"
  auto depth_run = [&](int base, int n) {
    if (MODE != 1 || n <= 0) return;
    if (pendingEnd != base) {
      if (pendingEnd >= 0) COV_RED_HOT(cov + pendingEnd, -1);
      COV_RED_HOT(cov + base, 1);
    }
    pendingEnd = base + n;
  };

  // Fixed loop body, re-converged with __syncwarp between the phases:
  //   (A0) a lane whose read is finished writes its results and takes the next read from the cursor; (Rc) stages its CIGAR;
  //   (A1) a lane whose M op has no windows left consumes at most one element that is not M, then (A2) at most one M
  //        element (straight-line predicated code);
  //   (R)  lanes that ran out of staged CIGAR / base words refill their own 80-byte tile with five 16-byte cp.async copies;
  //   (B)  every lane that is inside an M op compares up to 32 bases (four 8-base windows) from its tile.
  // Lanes never read the streams with plain loads: a lane-private 4-byte global load drags a 32-byte sector through L1 that
  // is evicted before its neighbours are used (ncu r01h: 13 KB of L2->L1 traffic per 1.7 KB read, 40 warps slower than 24).
  // The tiles were first filled by the whole warp (coalesced 64-byte rows, two lanes served per step): correct, but the
  // ballots, shuffles and clamps of that scheme were ~30 % of all issued instructions (ncu r01n).
  // Lanes that ran out of reads stay in the loop (idle) until the whole warp is done so that the
  // full-mask __syncwarp is legal; without it the compiler lets every lane run on its own (ncu r01e: 7 of 32 lanes active).
  bool done = false;
  uint32_t cs = 0;           // index of the first CIGAR word in my tile (multiple of 4, valid for [cs, cs+20)); offsets are uint32 in the ABI
  uint32_t ckg = 0, cendg = 0;  // CIGAR cursor / end of this read
  uint32_t ss = 0x80000000u; // first base word in my tile (multiple of 4), relative to rp16 (valid for [ss, ss+20))
  long long rp16 = 0;        // word index of the 16-byte aligned chunk that holds this read's first base
  uint32_t dmask = 0; int dpos = 0;  // MODE 1: error positions of a D/X op, flushed in phase (B)
  for (;;) {
    __syncwarp();
    // ---- (A0) read transitions: finish the current read and/or take the next one from the cursor
    if (!done && mOff >= mN && (!active || ckg == cendg)) {
      if (active) {
        // ---- end of read: CigarCluster.addEnd :466-469, read offsets of alignment start/end
        const int alnEnd = refPos;  // alnStart + reference span - 1
        if (MODE == 2 || (MODE == 0 && nb < BRK_INLINE)) bdst[nb] = alnEnd;
        nb++;
        if (MODE == 0) {
          b.nbreaks[r] = nb;
          b.break_loc[r] = 0xFFFFFFFFu;
          b.read_st[r] = st_r;
          b.read_en[r] = (alnEnd > 0 && maxEnd == alnEnd) ? (alnEnd < eBs ? 0 : alnEnd - eBs + eRs) : 0;
          b.maps_sum[r] = c.record_depth ? maps : 0;
          b.err_sum[r] = c.record_depth ? errs : 0;
          b.rflags[r] = b.fast ? 0x20u : 0u;
          if (nb > MAX_BREAKS) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_TOO_MANY_BREAKS);
        }
        if (MODE == 1) {
          if (pendingEnd >= 0) COV_RED_HOT(cov + pendingEnd, -1);
          // CigarCluster.setStartEnd :360-365 with the (trimmed) start/end of the read; positions past seqlen+1 only come
          // from alignments that run off the reference and are dropped by the dense rows
          const int sp = b.start_pos[r], ep = b.end_pos[r];
          if (sp >= 0 && sp < c.stride) COV_RED_HOT(cov + 2 * arr_stride + sp, 1);
          if (ep >= 0 && ep < c.stride) COV_RED_HOT(cov + 3 * arr_stride + ep, 1);
        }
        active = false;
      }
      {
        bool ok = true;
        r = (long long)atomicAdd(cursor, 1u);
        if (b.fast) {   // only the reads the fast walk handed over (fastwalk.cuh)
          if (r >= (long long)ldcg(b.bctr + 7)) { done = true; ok = false; r = 0; }
          else r = (long long)b.slow_list[r];
        } else if (r >= b.n) { done = true; ok = false; }
        if (ok && MODE == 2 && b.nbreaks[r] <= BRK_INLINE) ok = false;
        if (ok && MODE == 1) {
          const int slot = b.cluster[r];
          if (slot < 0) ok = false;  // bad record
          else {
            const int prov = ldcg(c.cl_rec + (unsigned)(ldcg(&c.cl[slot].word) & 0xffffffffu));
            if (prov >= c.cov_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_COV_FULL); ok = false; }
            else { cov = c.cov + ((long long)prov * c.S + __ldg(b.src + r)) * c.stride; pendingEnd = -1; }
          }
        }
        if (ok) {
          alnStart = __ldg(b.pos + r);
          ckg = __ldg(b.cigar_off + r); cendg = __ldg(b.cigar_off + r + 1);
          cs = ckg + 64u;               // nothing staged yet (ckg - cs wraps to a huge value)
          const unsigned long long so = __ldg(b.seq_off + r);
          if (MODE == 0) nbases = (int)min(2ull * (__ldg(b.seq_off + r + 1) - so), 0x7FFFFFFFull);
          rp16 = (long long)(so >> 4) << 2;
          ss = 0x80000000u;             // base tile belongs to another read
          nib0 = (uint32_t)(so & 15) * 2;
          if (MODE == 0) bdst = b.break_arena + r * BRK_INLINE;
          if (MODE == 2) {
            const unsigned n_all = (unsigned)b.nbreaks[r];
            const unsigned loc = atomicAdd(b.bctr, n_all);
            // beside a fast walk the bulk of the batch is clustered on another stream at this moment: only this chain is redone
            if (loc + n_all > b.break_cap) { atomicOr(b.bctr + (b.fast == 2 ? 8 : 1), 1u); b.break_loc[r] = 0xFFFFFFFEu; ok = false; }
            else { b.break_loc[r] = loc; bdst = b.break_arena + b.n * BRK_INLINE + loc; }
          }
        }
        if (ok && alnStart <= 0) {  // not a mapped record
          if (MODE == 0) { b.nbreaks[r] = 0; b.rflags[r] = b.fast ? 0x30u : 0x10u; b.read_st[r] = 0; b.read_en[r] = 0; b.maps_sum[r] = 0; b.err_sum[r] = 0; b.break_loc[r] = 0xFFFFFFFFu; }
          ok = false;
        }
        if (ok) {
          refPos = alnStart - 1; readPos = 0; maps = 0; errs = 0;
          stFound = false; st_r = 0; maxEnd = INT_MIN; eBs = 0; eRs = 0;
          if (MODE != 1) bdst[0] = alnStart;  // CigarCluster.addStart :442-446
          nb = 1;
          prev = alnStart; pmat = 0;
          active = true;
        }
      }
    }
    // ---- (Rc) CIGAR tile of the lanes that ran out of staged elements (a tile is valid by index)
    if (active && ckg < cendg && ckg - cs >= (uint32_t)TILE_W) {
      cs = ckg & ~3u;
      const uint32_t* src = b.cigar + cs;
      if ((long long)cs + TILE_W <= cig_words) {
#pragma unroll
        for (int j = 0; j < TILE_W / 4; j++) cp16(myCigS + 16 * j, src + 4 * j);
      } else {
#pragma unroll
        for (int j = 0; j < TILE_W / 4; j++) {
          const long long left = (cig_words - ((long long)cs + 4 * j)) * 4;
          cp16z(myCigS + 16 * j, left > 0 ? (const void*)(src + 4 * j) : (const void*)b.cigar, (int)max(0LL, min(16LL, left)));
        }
      }
      cp_wait_all();
    }
    __syncwarp();
    // ---- (A1) at most one element that is not M, then (A2) at most one M element: the usual pattern is one I/D element
    // between two M elements, so a lane normally enters (B) with a fresh window every iteration.  Both are straight-line
    // predicated code (lanes handle different op types side by side).
    bool haveW = active && mOff >= mN && ckg < cendg;
    uint32_t w = haveW ? myCig[(int)(ckg - cs)] : 0u;
    if (haveW && (w & 15u) != 0u) {
      ckg++;
      const int len = (int)(w >> 4), op = (int)(w & 15);
      if (op > 8) {  // the reference throws (IdentityProfile1.java:577-578): record not clustered
        if (MODE == 0) { b.nbreaks[r] = 0; b.rflags[r] = b.fast ? 0x30u : 0x10u; b.read_st[r] = 0; b.read_en[r] = 0; b.maps_sum[r] = 0; b.err_sum[r] = 0; b.break_loc[r] = 0xFFFFFFFFu; }
        active = false;
      } else {
        const bool isDX = (op == 2) | (op == 8), isEQ = op == 7;
        const int ref0 = refPos, read0 = readPos;
        refPos += ((0x18C >> op) & 1) ? len : 0;    // D N = X consume reference
        readPos += ((0x192 >> op) & 1) ? len : 0;   // I S = X consume read
        if (isEQ | (op == 8)) {
          // htsjdk AlignmentBlock(readStart, refStart, len) bookkeeping for getReadPositionAtReferencePosition
          const int bStart = ref0 + 1, bEnd = ref0 + len, rStart = read0 + 1;
          if (!stFound && bEnd >= alnStart) { st_r = alnStart < bStart ? 0 : alnStart - bStart + rStart; stFound = true; }
          if (bEnd > maxEnd) { maxEnd = bEnd; eBs = bStart; eRs = rStart; }
        }
        // D / = / X advance first and emit afterwards (quirk :513-522, :552-576)
        const int nE = (isDX | isEQ) ? max(0, min(len, G - refPos)) : 0;
        const int base = refPos + 1;
        maps += nE;
        errs += isDX ? nE : 0;
        // Short deletions (the bulk of the coverage events of a noisy read) are recorded as ONE point in a per-length
        // histogram instead of 4 depth-difference events + len error points: when the previous depth run ends right before
        // the deletion the run is simply carried across it, and coverage_rows_kernel removes the deleted positions and adds
        // the "advance, then emit" positions (depth and error) from the histogram at end of chromosome.
        // emitted positions more than T past the last match are mapStart/mapEnd events (and breaks for '=')
        bool far = false;
        if (nE > 0 && (isEQ || base + nE - 1 - (pmat ? ppos0 : prev) > T)) { sync_prev(); far = prev > 0 && base + nE - 1 - prev > T; }
        if (MODE == 1 && op == 2 && nE == len && len >= 1 && len <= DEL_HIST && pendingEnd == ref0 + 1 && !far) {
          COV_RED(cov + (long long)(3 + len) * arr_stride + ref0 + 1, 1);
          pendingEnd = base;
        } else if (MODE == 1 && nE > 0) {
          depth_run(base, nE);
          if (isDX) {
            if (nE <= 8) { dmask = 0x11111111u & (0xFFFFFFFFu << ((8 - nE) << 2)); dpos = base; }  // flushed after (B)
            else for (int i = 0; i < nE; i++) COV_RED(cov + arr_stride + base + i, 1);
            if (far)
              for (int i = max(0, prev + T + 1 - base); i < nE; i++) { COV_RED(cov + 2 * arr_stride + base + i, 1); COV_RED(cov + 3 * arr_stride + prev, 1); }
          }
        }
        if (isEQ && nE > 0) {
          if (per_base || base + nE - 1 - prev > T) for (int i = 0; i < nE; i++) add(base + i, true);
          else prev = base + nE - 1;
        }
      }
      haveW = active && ckg < cendg && ckg - cs < (uint32_t)TILE_W;
      w = haveW ? myCig[(int)(ckg - cs)] : 0u;
    }
    if (haveW && (w & 15u) == 0u) {
      // M emits from its own start
      ckg++;
      const int len = (int)(w >> 4);
      const int ref0 = refPos, read0 = readPos;
      refPos += len; readPos += len;
      const int bStart = ref0 + 1, bEnd = ref0 + len, rStart = read0 + 1;
      if (!stFound && bEnd >= alnStart) { st_r = alnStart < bStart ? 0 : alnStart - bStart + rStart; stFound = true; }
      if (bEnd > maxEnd) { maxEnd = bEnd; eBs = bStart; eRs = rStart; }
      const int nE = max(0, min(len, G - ref0));
      if (MODE == 0 && nE > 0 && (long long)read0 + nE > nbases) {
        // the reference's readSeq.getBase runs past the end of the read and throws (IdentityProfile1.java:531-551): not clustered
        b.nbreaks[r] = 0; b.rflags[r] = b.fast ? 0x30u : 0x10u; b.read_st[r] = 0; b.read_en[r] = 0; b.maps_sum[r] = 0; b.err_sum[r] = 0; b.break_loc[r] = 0xFFFFFFFFu;
        active = false;
      }
      maps += nE;
      mStart0 = ref0; mRpos = read0; mOff = 0; mN = active ? nE : 0;
      if (MODE == 1 && nE > 0) depth_run(ref0 + 1, nE);
    }
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    // ---- (R) the refill point of running reads.  CIGAR tile: the elements are consumed by the next iteration's phase A (a
    // tile is valid by index, so filling it early is harmless; (Rc) then only serves reads that have just started).  Base tile: the windows at read offset mRpos+mOff need words
    // kg..kg+STEP_WINS (relative to rp16).  Each lane copies for itself; nothing is shared between lanes.
    const bool hasWin = !done && mOff < mN;
    const uint32_t nibw = nib0 + (uint32_t)(mRpos + mOff);
    const uint32_t kg = nibw >> 3;  // word index relative to rp16
    {
      const bool needC = active && ckg < cendg && ckg - cs >= (uint32_t)TILE_W;
      const bool needS = hasWin && kg - ss > (uint32_t)(TILE_W - 1 - STEP_WINS);
      if (needC) {
        cs = ckg & ~3u;
        const uint32_t* src = b.cigar + cs;
        if ((long long)cs + TILE_W <= cig_words) {
#pragma unroll
          for (int j = 0; j < TILE_W / 4; j++) cp16(myCigS + 16 * j, src + 4 * j);
        } else {
#pragma unroll
          for (int j = 0; j < TILE_W / 4; j++) {
            const long long left = (cig_words - ((long long)cs + 4 * j)) * 4;
            cp16z(myCigS + 16 * j, left > 0 ? (const void*)(src + 4 * j) : (const void*)b.cigar, (int)max(0LL, min(16LL, left)));
          }
        }
      }
      if (needS) {
        ss = kg & ~3u;
        const long long w0 = rp16 + ss;
        const uint32_t* src = seqw + w0;
        if (w0 + TILE_W <= seq_words) {
#pragma unroll
          for (int j = 0; j < TILE_W / 4; j++) cp16(mySeqS + 16 * j, src + 4 * j);
        } else {
#pragma unroll
          for (int j = 0; j < TILE_W / 4; j++) {
            const long long left = (seq_words - (w0 + 4 * j)) * 4;
            cp16z(mySeqS + 16 * j, left > 0 ? (const void*)(src + 4 * j) : (const void*)seqw, (int)max(0LL, min(16LL, left)));
          }
        }
      }
      if (needC || needS) cp_wait_all();
    }
    __syncwarp();
    uint32_t mis = 0;  // mismatches of this step: bit 4j+q <-> base 8q + 7 - j
    int pos0 = 0;
    if (hasWin) {
      // ---- (B) up to STEP_BASES bases (8-base windows) of the current M op
      const int cnt = min(STEP_BASES, mN - mOff);
      pos0 = mStart0 + 1 + mOff;
      const int ki = (int)(kg - ss);
      const uint32_t sh = (nibw & 7) << 2;
      const uint32_t rn = (uint32_t)(mStart0 + mOff), rk = rn >> 3, rs = (rn & 7) << 2;
      uint32_t w[STEP_WINS + 1], g[STEP_WINS + 1];
#pragma unroll
      for (int q = 0; q <= STEP_WINS; q++) { w[q] = __byte_perm(mySeq[ki + q], 0, 0x0123); g[q] = refw[rk + q]; }
      uint32_t nzc = 0;  // bases that differ (valid or not), same layout as mis
#pragma unroll
      for (int q = 0; q < STEP_WINS; q++) {
        const uint32_t x = __funnelshift_l(w[q + 1], w[q], sh) ^ __funnelshift_l(g[q + 1], g[q], rs);
        nzc |= ((x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u) << q;
      }
      const uint32_t vm = s_vm[cnt];
      mis = nzc & vm;
      const uint32_t mat = vm & ~nzc;  // matches
      mOff += STEP_BASES;
      errs += __popc(mis);
      // Fast path when every position of the step is within T of the last match and that match does not lie AHEAD of the
      // step (then no position can be a break, whatever the matches inside).  The test uses a lower bound of prev; only
      // if it fails is prev decoded.  After an '=' op prev can lie ahead of the step (advance-then-emit quirk): matches
      // then pull prev back and a gap can open inside the step.
      if (per_base || pos0 + cnt - 1 - (pmat ? ppos0 : prev) > T || (pmat == 0 && prev > pos0)) {
        sync_prev();
        const int l = last_match(mat);
        if (per_base || prev > pos0) {
          // exact per-base CigarCluster.add (T < STEP_BASES, or prev ahead of the step)
          slow_windows++;
#pragma unroll 1
          for (int i = 0; i < cnt; i++) add(pos0 + i, (mat >> (((7 - (i & 7)) << 2) + (i >> 3))) & 1u);
        } else if (pos0 + cnt - 1 - prev > T) {
          // The step right after a junction (or after a long run without a match), in closed form.  Positions i >= i0 are
          // more than T past the last match; every one of them up to and including the first match f of the step is a
          // break_p position (mapStart[p]++, mapEnd[prev]++ each, CigarCluster.java:476-486), the first match itself
          // closes the break (prev, p).  After that match every position of the step is within STEP_BASES-1 <= T of one.
          slow_windows++;
          int f = cnt;
#pragma unroll
          for (int q = STEP_WINS - 1; q >= 0; q--) { const uint32_t m = (mat >> q) & 0x11111111u; if (m) f = 8 * q + (__clz(m) >> 2); }
          const int i0 = prev < 0 ? 0 : max(0, prev + T + 1 - pos0);
          if (f >= i0) {
            if (MODE == 1 && prev > 0) {
              const int last = min(f, cnt - 1);
              for (int i = i0; i <= last; i++) COV_RED(cov + 2 * arr_stride + pos0 + i, 1);
              COV_RED(cov + 3 * arr_stride + prev, last - i0 + 1);
            }
            if (f < cnt) {
              if (prev > 0) { if (MODE == 2 || (MODE == 0 && nb < BRK_INLINE)) bdst[nb] = prev; nb++; }
              if (MODE == 2 || (MODE == 0 && nb < BRK_INLINE)) bdst[nb] = pos0 + f;
              nb++;
            }
          }
          if (l >= 0) prev = pos0 + l;
        } else if (l >= 0) {
          prev = pos0 + l;
        }
      } else if (mat) {
        pmat = mat; ppos0 = pos0;  // every position is within T of the last match before it
      }
    }
    if (MODE == 1) {
      // error positions: mismatches of this step and/or the D/X op consumed in phase (A)
      int* err_row = cov + arr_stride;
      while (mis) {
        const int bit = 31 - __clz(mis);
        mis &= ~(1u << bit);
        COV_RED(err_row + pos0 + ((bit & 3) << 3) + 7 - (bit >> 2), 1);
      }
      while (dmask) {
        const int bit = 31 - __clz(dmask);
        dmask &= ~(1u << bit);
        COV_RED(err_row + dpos + ((28 - bit) >> 2), 1);
      }
    }
  }
  if (MODE == 0 && slow_windows && !b.fast) atomicAdd(c.counters + CN_SLOW, (unsigned)slow_windows);
}

// ------------------------------------------------------------------------------------------------ firstIsTranscriptome
// Reads of sources != 0 contribute to Count.true_breaks / break_sum only while their sub-cluster has seen no source-0 read
// (CigarCluster.java:85).  Runs at end of chromosome, before slots are rewritten to ids.
__global__ void fit_fixup_kernel(const DevCfg c, const BatchDev b) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += (long long)gridDim.x * blockDim.x) {
    const int sb_slot = b.sub[r];
    if (sb_slot < 0 || b.src[r] == 0) continue;
    const unsigned long long gidx = (unsigned long long)(b.idx_base + r);
    if (gidx >= ldcg(c.sub_min0 + sb_slot)) continue;
    const int nb = b.nbreaks[r];
    const unsigned loc = b.break_loc[r];
    const int* br = (loc == 0xFFFFFFFFu) ? b.break_arena + r * BRK_INLINE : b.break_arena + b.n * BRK_INLINE + loc;
    atomicAdd(c.sub_break_sum + sb_slot, 1);
    const unsigned off = (unsigned)(ldcg(&c.sb[sb_slot].word) & 0xffffffffu);
    const unsigned soff = (unsigned)ldcg(c.sb_rec + off + 2);
    if (ldcg(c.sb_rec + off + 3) == nb)
      for (int i = 0; i < nb; i++) atomicAdd(c.sub_sums + soff + i, (unsigned long long)(long long)br[i]);
  }
}

// ------------------------------------------------------------------------------------------------ annotation lookups
// Annotation.nextUpstream :87-97 (highest row i with leftBreak + tol >= start[i]) / EmptyAnnotation :22-24
__device__ __forceinline__ int tok_up(const DevCfg& c, const int* ostart, const int* oname, const int* ostrand, int p, bool fwd) {
  if (c.annot_kind == 1) return p / c.bin;
  for (int i = c.n_orf - 1; i >= 0; i--)
    if ((!c.enforce_strand || (int)fwd == ostrand[i]) && p + c.bin >= ostart[i]) return oname[i];
  return -1;  // NPT_TOK_ST
}
// Annotation.nextDownstream :62-72 (lowest row i with rightBreak - tol <= start[i])
__device__ __forceinline__ int tok_down(const DevCfg& c, const int* ostart, const int* oname, const int* ostrand, int p, bool fwd) {
  if (c.annot_kind == 1) return p / c.bin;
  for (int i = 0; i < c.n_orf; i++)
    if ((!c.enforce_strand || (int)fwd == ostrand[i]) && p - c.bin <= ostart[i]) return oname[i];
  return -2;  // NPT_TOK_END
}

// GFFAnnotation.nextUpstream(l1, r1, chrom_index, forward) :80-100 with contains :72-76: among the exon rows that contain
// both block ends within the tolerance (and lie on the read's strand when it is known) the one with the smallest
// |l1-start| + |r1-end|, lowest file row on ties (the TreeMap keeps the last put of the descending scan), accepted iff that
// score <= 2*tolerance.  An accepted row has |l1-start| <= 2*tol, and `contains` needs start <= l1+tol, so only rows with
// start in [l1-2tol, l1+tol] can be the answer: a binary search in the start-sorted table plus a short scan.
// Returns the row's gene name id, -1 for null.
__device__ int gff_upstream(const DevCfg& c, int l1, int r1, bool fwd_known, bool fwd) {
  if (l1 < 0) return -1;
  const int tol = c.bin;
  int lo = 0, hi = c.gx_n;
  const int smin = l1 - 2 * tol;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&c.gx[mid].x) < smin) lo = mid + 1; else hi = mid; }
  int best = INT_MAX, best_row = INT_MAX, best_name = -1;
  for (int i = lo; i < c.gx_n; i++) {
    const int4 e = __ldg(c.gx + i);
    if (e.x > l1 + tol) break;
    if (fwd_known && (int)fwd != (e.z & 1)) continue;
    if (!(r1 - tol <= e.y && r1 + tol >= e.x && l1 - tol <= e.y)) continue;
    const int score = abs(l1 - e.x) + abs(r1 - e.y), row = e.z >> 1;
    if (score < best || (score == best && row < best_row)) { best = score; best_row = row; best_name = e.w; }
  }
  return best <= 2 * tol ? best_name : -1;
}

// break-point matrix cell += 1   (BreakPoints.addBreakPoint, J/cluster/BreakPoints.java:102-108)
__device__ __forceinline__ unsigned long long pair_key(int src, int level, int st, int en) {
  return (1ULL << 63) | ((unsigned long long)src << 59) | ((unsigned long long)level << 58) | ((unsigned long long)(unsigned)st << 29) |
         (unsigned long long)(unsigned)en;
}
// find-or-insert of a cell without counting (keys imported from another rank)
__device__ void pair_touch(const DevCfg& c, unsigned long long key) {
  unsigned slot = (unsigned)mix64(key) & c.pr_mask;
  for (unsigned probe = 0; probe <= c.pr_mask; probe++, slot = (slot + 1) & c.pr_mask) {
    PairEntry* e = c.pr + slot;
    unsigned long long k = ldcg(&e->key);
    if (k == 0) {
      k = atomicCAS(&e->key, 0ULL, key);
      if (k == 0) {
        const unsigned np = atomicAdd(c.counters + CN_PAIRS, 1u);
        if ((int)np >= c.pr_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
        return;
      }
    }
    if (k == key) return;
  }
  atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
}
__device__ void pair_add(const DevCfg& c, int src, int level, int st, int en) {
  const unsigned long long key = pair_key(src, level, st, en);
  unsigned slot = (unsigned)mix64(key) & c.pr_mask;
  for (unsigned probe = 0; probe <= c.pr_mask; probe++, slot = (slot + 1) & c.pr_mask) {
    PairEntry* e = c.pr + slot;
    unsigned long long k = ldcg(&e->key);
    if (k == 0) {
      k = atomicCAS(&e->key, 0ULL, key);
      if (k == 0) {
        const unsigned np = atomicAdd(c.counters + CN_PAIRS, 1u);
        if ((int)np >= c.pr_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
        k = key;
      }
    }
    if (k == key) {
      // opportunistic warp aggregation: the canonical junctions are hit by a large share of the reads
      const unsigned grp = __match_any_sync(__activemask(), (unsigned long long)(uintptr_t)e);
      if ((__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&e->count, __popc(grp));
      return;
    }
  }
  atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
}

// ------------------------------------------------------------------------------------------------ table inserts
// Cluster table: find or insert a token list (token(k), k < ntok).  Returns the slot, -1 if the table or arena is full.
template <typename TokenFn>
__device__ int cluster_find_or_insert(const DevCfg& c, int ntok, TokenFn token) {
  unsigned long long h = 0x1234ULL ^ ((unsigned long long)ntok << 48);
  for (int k = 0; k < ntok; k++) h = mix64(h + (unsigned)token(k));
  const unsigned h32 = (unsigned)(h >> 32) | 1u;
  int cl_slot = -1;
  unsigned slot = (unsigned)h & c.cl_mask;
  unsigned my_off = 0xFFFFFFFFu;
  for (unsigned probe = 0; probe <= c.cl_mask; probe++, slot = (slot + 1) & c.cl_mask) {
    unsigned long long w = ldcg(&c.cl[slot].word);
    if (w == 0) {
      if (my_off == 0xFFFFFFFFu) {  // write the record first, publish it with the CAS
        my_off = atomicAdd(c.counters + CN_TOK_USED, (unsigned)(CL_REC_HDR + ntok));
        if (my_off + CL_REC_HDR + ntok > c.tok_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_ARENA_FULL); break; }
        c.cl_rec[my_off] = -1; c.cl_rec[my_off + 1] = ntok;
        for (int k = 0; k < ntok; k++) c.cl_rec[my_off + CL_REC_HDR + k] = token(k);
        __threadfence();
      }
      w = atomicCAS(&c.cl[slot].word, 0ULL, ((unsigned long long)h32 << 32) | my_off);
      if (w == 0) {
        const unsigned prov = atomicAdd(c.counters + CN_CLUSTERS, 1u);
        if ((int)prov >= c.cl_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_CL_FULL);
        c.cl_rec[my_off] = (int)prov;  // read by the coverage pass (a later kernel) and at end of chromosome
        cl_slot = (int)slot;
        break;
      }
    }
    if ((unsigned)(w >> 32) == h32) {  // verify the full key
      const unsigned off = (unsigned)(w & 0xffffffffu);
      bool eq = ldcg(c.cl_rec + off + 1) == ntok;
      for (int k = 0; eq && k < ntok; k++) eq = ldcg(c.cl_rec + off + CL_REC_HDR + k) == token(k);
      if (eq) { cl_slot = (int)slot; break; }
    }
  }
  if (cl_slot < 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_CL_FULL);
  return cl_slot;
}
// Sub-cluster table: key = (cluster slot, key(i), i < nkey) — CigarCluster.cloneBreaks :620-631.  nsum = entries reserved in the
// int64 break-sum arena when the record is created.  Returns the slot, -1 if full.
template <typename KeyFn>
__device__ int sub_find_or_insert(const DevCfg& c, int cl_slot, int nkey, int nsum, KeyFn key) {
  unsigned long long hs = 0x77ULL + (unsigned long long)(unsigned)cl_slot * 0x9E3779B97F4A7C15ULL + ((unsigned long long)nkey << 48);
  for (int i = 0; i < nkey; i++) hs = mix64(hs + (unsigned)key(i));
  const unsigned hs32 = (unsigned)(hs >> 32) | 1u;
  unsigned slot = (unsigned)hs & c.sb_mask;
  unsigned my_off = 0xFFFFFFFFu;
  for (unsigned probe = 0; probe <= c.sb_mask; probe++, slot = (slot + 1) & c.sb_mask) {
    unsigned long long w = ldcg(&c.sb[slot].word);
    if (w == 0) {
      if (my_off == 0xFFFFFFFFu) {
        my_off = atomicAdd(c.counters + CN_SUBKEY_USED, (unsigned)(SB_REC_HDR + nkey));
        const unsigned soff = nsum ? atomicAdd(c.counters + CN_SUMS_USED, (unsigned)nsum) : 0u;
        if (my_off + SB_REC_HDR + nkey > c.sbkey_cap || soff + nsum > c.sums_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_ARENA_FULL); return -1; }
        c.sb_rec[my_off] = cl_slot; c.sb_rec[my_off + 1] = nkey; c.sb_rec[my_off + 2] = (int)soff; c.sb_rec[my_off + 3] = nsum;
        for (int i = 0; i < nkey; i++) c.sb_rec[my_off + SB_REC_HDR + i] = key(i);
        __threadfence();
      }
      w = atomicCAS(&c.sb[slot].word, 0ULL, ((unsigned long long)hs32 << 32) | my_off);
      if (w == 0) {
        const unsigned ns = atomicAdd(c.counters + CN_SUBS, 1u);
        if ((int)ns >= c.sb_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_SUB_FULL);
        return (int)slot;
      }
    }
    if ((unsigned)(w >> 32) == hs32) {
      const unsigned off = (unsigned)(w & 0xffffffffu);
      bool eq = ldcg(c.sb_rec + off) == cl_slot && ldcg(c.sb_rec + off + 1) == nkey;
      for (int i = 0; eq && i < nkey; i++) eq = ldcg(c.sb_rec + off + SB_REC_HDR + i) == key(i);
      if (eq) return (int)slot;
    }
  }
  return -1;
}

// ------------------------------------------------------------------------------------------------ clustering
constexpr int CLUSTER_THREADS = 128;

// One thread per read.  Key tokens are functions of the break list, so they are regenerated on the fly for hashing,
// comparing and storing instead of being held in registers.
#ifndef NPT_CLUSTER_MIN_BLOCKS
#define NPT_CLUSTER_MIN_BLOCKS 1
#endif
// which = 0: every read of the batch; 1: the reads the fast walk finished (the serial walker may still be writing the
// others: only their SLOWPATH bit is looked at); 2: the reads of slow_list.  1 and 2 run concurrently on two streams.
template <int which>
__global__ void __launch_bounds__(CLUSTER_THREADS, NPT_CLUSTER_MIN_BLOCKS) cluster_kernel(const DevCfg c, const BatchDev b) {
  extern __shared__ __align__(16) int s_int[];
  if (ldcg(b.bctr + 1)) return;  // break lists incomplete (overflow area too small): the host re-runs this batch
  if (which == 2 && ldcg(b.bctr + 8)) return;
  int* s_ostart = s_int;
  int* s_oname = s_ostart + c.n_orf;
  int* s_ostrand = s_oname + c.n_orf;
  int* s_cnt = s_ostrand + c.n_orf;  // total_counts[S], spliced[S][n_orf], unspliced[S][n_orf] (block-privatised)
  const int ncnt = c.S * (1 + 2 * c.n_orf);
  for (int i = threadIdx.x; i < c.n_orf; i += blockDim.x) { s_ostart[i] = c.orf_start[i]; s_oname[i] = c.orf_name[i]; s_ostrand[i] = c.orf_strand[i]; }
  for (int i = threadIdx.x; i < ncnt; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();

  const long long n_mine = which == 2 ? (long long)ldcg(b.bctr + 7) : b.n;
  for (long long k0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; k0 < n_mine; k0 += (long long)gridDim.x * blockDim.x) {
    const long long r = which == 2 ? (long long)b.slow_list[k0] : k0;
    const unsigned rf0 = which == 1 ? ldcg(b.rflags + r) : b.rflags[r];
    const unsigned slowbit = rf0 & 0x20u;  // diagnostic: walked by the serial path
    if (which == 1 && slowbit) continue;
    if (rf0 & 0x10u) {  // bad record: not clustered
      b.cluster[r] = -1; b.sub[r] = -1; b.start_pos[r] = 0; b.end_pos[r] = 0;
      continue;
    }
    int nb = b.nbreaks[r];
    unsigned loc = b.break_loc[r];
    int* br = (loc == 0xFFFFFFFFu) ? b.break_arena + r * BRK_INLINE : b.break_arena + b.n * BRK_INLINE + loc;
    const int src = b.src[r];
    if (src >= c.S) {  // a source index outside the configuration: the reference would throw; reported as NPT_EINVAL
      atomicOr(c.counters + CN_ERROR, (unsigned)ERR_BAD_SRC);
      b.cluster[r] = -1; b.sub[r] = -1; b.start_pos[r] = 0; b.end_pos[r] = 0; b.rflags[r] = 0x10u | slowbit;
      continue;
    }
    const bool neg = (b.flag[r] & 0x10) != 0;
    // ---- small first/last exon removal (IdentityProfile1.java:205-228), host mode only: drops the first and/or last pair
    if (c.min_fl_exon > 0 && nb > 2) {
      const int m = c.min_fl_exon;
      const int diff0 = br[1] - br[0], diff1 = br[nb - 1] - br[nb - 2];
      int a = 0, e = nb;
      if (nb == 4 && diff0 < m && diff1 < m) { if (diff1 < diff0) e -= 2; else a += 2; }
      else { if (diff1 < m) e -= 2; if (diff0 < m) a += 2; }
      if (a) {
        if (loc == 0xFFFFFFFFu) { for (int i = a; i < e; i++) br[i - a] = br[i]; }
        else { loc += (unsigned)a; b.break_loc[r] = loc; br += a; }
      }
      nb = e - a;
      b.nbreaks[r] = nb;
    }
    const int startPos = br[0], endPos = br[nb - 1];
    const bool fwd_known = (c.rna_mask >> src) & 1;
    const bool fwd = !neg;  // meaningful when fwd_known; in coronavirus mode callers guarantee rna[src]
    bool leader = false;
    if (c.coronavirus && nb > 1) leader = (c.annot_kind == 1) ? (br[1] < 100) : (br[1] < s_ostart[1] - c.bin);

    // ---- GFF mode (:300-323): one exon lookup per block; the key is the set of genes hit, in the iteration order of the
    // reference's HashSet<String>, or — when no block hits an exon — the binned coordinate of one block end
    int gtok[GFF_MAX_GENES + 1];
    int gff_n = 0;
    if (c.annot_kind == 2) {
      unsigned gh[GFF_MAX_GENES];
      int ng = 0, coord_first = -1, coord_last = -1;
      bool tree_bin = false;
      unsigned cap = 16;  // HashMap table size: doubles past 0.75 load, or when a chain already holds 8 nodes (treeifyBin < 64)
      const bool fwdOrNull = !fwd_known || fwd;
      for (int i = 0; i + 1 < nb; i += 2) {
        const int name = gff_upstream(c, br[i], br[i + 1], fwd_known, fwd);
        if (name >= 0) {
          bool seen = false;
          for (int k = 0; k < ng; k++) seen |= gtok[1 + k] == name;
          if (!seen) {
            if (ng == GFF_MAX_GENES) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_GENESET); break; }
            const unsigned h = __ldg(c.gx_hash + name);
            int chain = 0;
            for (int k = 0; k < ng; k++) chain += ((gh[k] ^ h) & (cap - 1)) == 0;
            gtok[1 + ng] = name; gh[ng] = h; ng++;
            if (chain >= 8) { if (cap < 64) cap <<= 1; else tree_bin = true; }
            if ((unsigned)ng > cap * 3 / 4) cap <<= 1;
          }
        } else {
          const int cv = (fwdOrNull ? br[i + 1] : br[i]) / c.bin;  // GFFAnnotation.getCoord :132-135 (positions are >= 1)
          if (coord_first < 0) coord_first = cv;
          coord_last = cv;
        }
      }
      if (tree_bin) {  // a real tree bin (> 10 fully colliding names in one read): iteration order not restated, record dropped
        b.cluster[r] = -1; b.sub[r] = -1; b.start_pos[r] = 0; b.end_pos[r] = 0; b.rflags[r] = 0x10u | slowbit;
        continue;
      }
      if (ng > 0) {
        // iteration order: by bucket, insertion order inside a bucket (stable insertion sort on the bucket index)
        for (int i = 1; i < ng; i++) {
          const int t = gtok[1 + i]; const unsigned h = gh[i];
          int j = i;
          while (j > 0 && (gh[j - 1] & (cap - 1)) > (h & (cap - 1))) { gtok[1 + j] = gtok[j]; gh[j] = gh[j - 1]; j--; }
          gtok[1 + j] = t; gh[j] = h;
        }
        gtok[0] = -6;  // NPT_TOK_GENESET
        gff_n = 1 + ng;
      } else {
        gtok[0] = fwdOrNull ? coord_last : coord_first;
        gff_n = 1;
      }
    }
    // ---- key tokens (IdentityProfile1.java:266-299, 325-327) as a function of the token index
    const bool ends = c.include_start || nb == 2;
    const int nbody = c.annot_kind == 2 ? gff_n : (c.coronavirus ? (ends ? nb : nb - 1) : 1);
    const int ntok = nbody + (c.enforce_strand ? 1 : 0);
    auto token = [&](int k) -> int {
      if (k >= nbody) return fwd ? -3 : -4;                                          // '+' / '-'
      if (c.annot_kind == 2) return gtok[k];
      if (!c.coronavirus) return tok_up(c, s_ostart, s_oname, s_ostrand, (fwd_known && !fwd) ? br[0] : br[nb - 1], fwd);
      if (!ends && k == 0) return -5;                                                // NPT_TOK_NOENDS marker
      const bool isEnd = ends && (k == 0 || k == nb - 1);
      const bool up = isEnd || (k & 1);
      return up ? tok_up(c, s_ostart, s_oname, s_ostrand, br[k], fwd) : tok_down(c, s_ostart, s_oname, s_ostrand, br[k], fwd);
    };

    // ---- break-point pairs (:271-288): level 0 for the first internal gap > T of the read, 1 for the others
    if (c.coronavirus) {
      bool first = true;
      for (int i = 1; i < nb - 1; i += 2) {
        if (br[i + 1] - br[i] > c.T) {
          leader = true;
          if (c.calc_breaks1) pair_add(c, src, first ? 0 : 1, br[i], br[i + 1]);
          first = false;
        }
      }
    }
    // ---- type (Annotation.getTypeInd :233-236)
    int type_ind = 0;
    if (c.coronavirus) type_ind = (startPos <= c.start_thresh) ? (endPos >= c.G - c.end_thresh ? 0 : 1) : (endPos >= c.G - c.end_thresh ? 2 : 3);
    // ---- totalCounts (CigarClusters.java:81) and ORF spliced/unspliced counters (:349-356)
    atomicAdd(&s_cnt[src], 1);
    if (c.coronavirus) {
      const int st1 = br[nb - 2];
      const bool spliced = startPos <= c.start_thresh;
      for (int i = c.n_orf - 1; i >= 0; i--) {
        if (s_ostart[i] < st1 - 10) break;
        if (endPos > s_ostart[i] + 100) atomicAdd(&s_cnt[c.S + (spliced ? 0 : c.S * c.n_orf) + src * c.n_orf + i], 1);
      }
    }

    // ---- cluster table: find or insert the token list
    const unsigned long long gidx = (unsigned long long)(b.idx_base + r);
    const int cl_slot = cluster_find_or_insert(c, ntok, token);
    // ---- sub-cluster table: key = (cluster slot, binned breaks) — CigarCluster.cloneBreaks :620-631
    int sb_slot = -1;
    if (cl_slot >= 0) {
      if (gidx < ldcg(&c.cl[cl_slot].min_idx)) atomicMin(&c.cl[cl_slot].min_idx, gidx);
      int a = 0, e = nb;
      if (!c.include_start && fwd_known) { if (fwd) a = 1; else e = nb - 1; }
      const int nkey = e - a;
      const int nsum = c.track_sums ? nb : 0;
      sb_slot = sub_find_or_insert(c, cl_slot, nkey, nsum, [&](int i) { return br[a + i] / c.bin; });
      if (sb_slot < 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_SUB_FULL);
      else {
        const unsigned long long idxf = (gidx << 2) | (unsigned long long)(((fwd_known && fwd) ? 1 : 0) | (neg ? 2 : 0));
        if (idxf < ldcg(&c.sb[sb_slot].min_idx)) atomicMin(&c.sb[sb_slot].min_idx, idxf);
        // Count.addBreaks :84-96 (exact integer sums).  With firstIsTranscriptome only source-0 reads add here; a read of
        // another source adds iff no source-0 read of its sub-cluster came before it (count[0]==0 at that time), which is
        // decided once the first source-0 index is final (fit_fixup_kernel at end of chromosome).
        if (c.fit && src == 0 && gidx < ldcg(c.sub_min0 + sb_slot)) atomicMin(c.sub_min0 + sb_slot, gidx);
        const bool addb = c.track_sums && (!c.fit || src == 0);
        // Warp-aggregated updates: a few sub-clusters own most reads (one owns ~40 % of the dRNA workload), so per-read
        // atomics on their counters serialise in L2 (ncu r01m: cluster_kernel 6 % issue-active).  Lanes of the warp that
        // hit the same (sub-cluster, source) elect a leader that adds the group's totals once.
        const unsigned conv = __activemask();
        const unsigned grp = __match_any_sync(conv, ((unsigned long long)(unsigned)sb_slot << 8) | ((unsigned long long)src << 1) | (addb ? 1ull : 0ull));
        const bool leader_lane = (__ffs(grp) - 1) == (int)(threadIdx.x & 31);
        const int gn = __popc(grp);
        if (leader_lane) atomicAdd(c.sub_count + (size_t)sb_slot * c.S + src, gn);  // Count.increment / new Count
        if (addb) {
          if (leader_lane) atomicAdd(c.sub_break_sum + sb_slot, gn);
          const unsigned off = (unsigned)(ldcg(&c.sb[sb_slot].word) & 0xffffffffu);
          const unsigned soff = (unsigned)ldcg(c.sb_rec + off + 2);
          if (ldcg(c.sb_rec + off + 3) == nb)   // same sub-cluster => same number of breaks for every lane of the group
            for (int i = 0; i < nb; i++) {
              const unsigned v = (unsigned)br[i];
              const unsigned lo = __reduce_add_sync(grp, v & 0xFFFFu), hi = __reduce_add_sync(grp, v >> 16);
              if (leader_lane) atomicAdd(c.sub_sums + soff + i, ((unsigned long long)hi << 16) + lo);
            }
        }
      }
    }
    b.cluster[r] = cl_slot; b.sub[r] = sb_slot; b.start_pos[r] = startPos; b.end_pos[r] = endPos;
    b.rflags[r] = (unsigned)type_ind | (leader ? 4u : 0u) | (neg ? 8u : 0u) | slowbit;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncnt; i += blockDim.x) if (s_cnt[i]) atomicAdd(c.small + i, s_cnt[i]);
}

// ------------------------------------------------------------------------------------------------ multi-rank key exchange
// Packed key buffer of one rank (int32 words), produced by npt_export_keys and consumed by npt_import_keys on the other ranks:
//   [0..16) header: magic, cl_n, cl_rec_words, sb_n, sb_rec_words, pr_n, cl_slots, S, then the word offsets of the five sections
//   cl_list  [cl_n][4]  slot, record offset, first index lo, hi
//   cl_rec   the rank's cluster record arena (records [id, ntok, tokens...])
//   sb_list  [sb_n][6]  record offset, (first index << 2 | flags) lo, hi, first source-0 index lo, hi, 0
//   sb_rec   the rank's sub-cluster record arena (records [cluster slot, nkey, -, nsum, keys...])
//   pr_keys  [pr_n][2]  break-point pair key lo, hi
// Importing inserts the KEYS with their first-appearance indices and no counts: afterwards every rank's tables hold the
// same key set and minima, so the sort-and-rank at end of chromosome yields the same global numbering everywhere and the
// counters (in that order) can be summed with one all-reduce.
constexpr int KX_MAGIC = 0x4b58, KX_HDR = 16;

__global__ void import_clusters_kernel(const DevCfg c, const int* list, int n, const int* rec, int* slotmap) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int rslot = list[4 * i], off = list[4 * i + 1];
    const unsigned long long first = (unsigned)list[4 * i + 2] | ((unsigned long long)(unsigned)list[4 * i + 3] << 32);
    const int ntok = rec[off + 1];
    const int slot = cluster_find_or_insert(c, ntok, [&](int k) { return rec[off + CL_REC_HDR + k]; });
    slotmap[rslot] = slot;
    if (slot >= 0 && first < ldcg(&c.cl[slot].min_idx)) atomicMin(&c.cl[slot].min_idx, first);
  }
}
__global__ void import_subs_kernel(const DevCfg c, const int* list, int n, const int* rec, const int* slotmap) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int off = list[6 * i];
    const unsigned long long first = (unsigned)list[6 * i + 1] | ((unsigned long long)(unsigned)list[6 * i + 2] << 32);
    const unsigned long long first0 = (unsigned)list[6 * i + 3] | ((unsigned long long)(unsigned)list[6 * i + 4] << 32);
    const int cl_slot = slotmap[rec[off]];
    if (cl_slot < 0) continue;  // its cluster did not fit: already flagged
    const int nkey = rec[off + 1], nsum = rec[off + 3];
    const int slot = sub_find_or_insert(c, cl_slot, nkey, nsum, [&](int k) { return rec[off + SB_REC_HDR + k]; });
    if (slot < 0) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_SUB_FULL); continue; }
    if (first < ldcg(&c.sb[slot].min_idx)) atomicMin(&c.sb[slot].min_idx, first);
    if (c.fit && first0 < ldcg(c.sub_min0 + slot)) atomicMin(c.sub_min0 + slot, first0);
  }
}
__global__ void import_pairs_kernel(const DevCfg c, const int* keys, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    pair_touch(c, (unsigned)keys[2 * i] | ((unsigned long long)(unsigned)keys[2 * i + 1] << 32));
}

}  // namespace npt
