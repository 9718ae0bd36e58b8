// walk.cuh — the hot kernel: one warp per read.
//
//   phase 1  CIGAR decode (32 ops per step, 128-byte coalesced loads), warp scans for ref/read offsets,
//            4-bit base compare 8 bases per 32-bit word, break emission by ballot compaction into shared memory
//            (restates IdentityProfile1.processCigar + CigarCluster.add, J/cluster/IdentityProfile1.java:482-585,
//            J/cluster/CigarCluster.java:442-498, including the D/=/X "advance, then emit" quirk)
//   phase 2  break list -> key tokens / break-point pairs / ORF counters / type
//            (IdentityProfile1.processRefPositions :149-356, Annotation.nextUpstream/nextDownstream :62-97)
//   phase 3  deterministic hash tables: cluster (token list) and sub-cluster (binned breaks); provisional dense ids +
//            atomicMin(first read index); ranks are assigned later by sort-and-rank (CigarClusters.matchCluster :75-102,
//            CigarCluster.merge :633-698)
//   phase 4  coverage: the read's depth runs were staged in shared memory as difference-array events (+1 at run start,
//            -1 after run end) and are flushed with one RED per event into the cluster's row; errors / mapStart / mapEnd
//            are point events.  A segmented scan at end of chromosome turns the difference rows into depth.
//
// Reads whose structure the windowed fast path cannot decide exactly (a run of mismatches longer than breakThresh inside
// one op, more events than the staging buffer holds, breakThresh < 8) are re-walked by an exact per-base serial walker on
// lane 0 of the same warp — still on the device; there is no host fallback.
#pragma once
#include "npt_device.cuh"

namespace npt {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int MAXB = 256;       // breaks per read held in shared memory
constexpr int MAXEV = 448;      // staged run/start/end events per read
constexpr int MAXMM = 256;      // staged error windows per read (position of an 8-base window + its mismatch mask)
constexpr int BRK_INLINE = 6;   // breaks stored in the per-read inline slot

constexpr unsigned EV_DEPTH_INC = 0, EV_DEPTH_DEC = 1, EV_MSTART = 3, EV_MEND = 4;
constexpr unsigned EV_POS_MASK = (1u << 29) - 1;

struct WarpScratch {
  int br[MAXB];
  int tok[MAXB + 2];
  // descriptors of the M ops of the current 32-op chunk, indexed by their rank among the chunk's M ops
  int m_start0[32], m_rpos0[32], m_nE[32], m_wstart[32], m_fm[32], m_lm[32];
  unsigned ev[MAXEV];        // (type << 29) | position
  unsigned mm_pos[MAXMM];    // 1-based position of the first base of an 8-base window
  unsigned mm_mask[MAXMM];   // error mask of that window, bit 28-4j <-> base j
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ unsigned lanemask_le() { unsigned m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

template <typename T>
__device__ __forceinline__ T ldcg(const T* p) { return __ldcg(p); }

__device__ __forceinline__ int warp_incl_scan(int v, unsigned lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane >= d) v += t;
  }
  return v;
}

// 8 reference bases starting at 0-based index `nib` (words hold the first base in bits 31..28)
__device__ __forceinline__ uint32_t ref_win8(const uint32_t* w, uint32_t nib) {
  uint32_t k = nib >> 3, o = (nib & 7) << 2;
  return __funnelshift_l(w[k + 1], w[k], o);
}
// 8 read bases from the BAM-packed byte stream (little-endian word loads, byte-swapped to base order).
// Reads one word past the window: batches carry >= 8 readable bytes after seq4 (npt.h).
__device__ __forceinline__ uint32_t read_win8(const uint32_t* w, uint32_t nib) {
  uint32_t k = nib >> 3, o = (nib & 7) << 2;
  uint32_t a = __byte_perm(__ldg(w + k), 0, 0x0123);
  uint32_t b = __byte_perm(__ldg(w + k + 1), 0, 0x0123);
  return __funnelshift_l(b, a, o);
}
__device__ __forceinline__ int read_base(const uint32_t* w, uint32_t nib) { return (int)(read_win8(w, nib) >> 28); }
__device__ __forceinline__ int ref_base(const uint32_t* w, uint32_t nib) { return (int)((w[nib >> 3] >> (28 - ((nib & 7) << 2))) & 0xF); }

// mismatch / match masks of an 8-base window (bit 28-4j <-> base j), cnt valid bases
__device__ __forceinline__ void cmp8(uint32_t rw, uint32_t gw, int cnt, uint32_t& mis, uint32_t& mat) {
  uint32_t x = rw ^ gw;
  uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
  uint32_t vm = 0x11111111u & (0xFFFFFFFFu << ((8 - cnt) << 2));
  mis = nz & vm;
  mat = vm & ~nz;
}

// Where fast_walk sends coverage events.  SINK 0: nowhere (depth off, or first pass of a long read);
// SINK 1: shared-memory staging (flushed once the cluster is known); SINK 2: straight to the cluster's rows.
struct CovSink { int* cov; long long arr_stride; };

// error positions of one window -> REDs on the cluster's error row
__device__ __forceinline__ void red_error_window(int* err_row, unsigned pos0, uint32_t mask) {
  while (mask) {
    const int bit = 31 - __clz(mask);
    mask &= ~(1u << bit);
    atomicAdd(err_row + pos0 + ((28 - bit) >> 2), 1);
  }
}

struct ReadCtx {
  int alnStart, alnEnd, src, neg;
  uint32_t c0, c1;
  const uint32_t* wread;
  uint32_t nib0, lastw;
  int st_r, end_r;
  int maps_sum, err_sum;
  int nb;
  bool bad;
};

// ------------------------------------------------------------------------------------------------ serial exact walker
// Lane 0 only.  Literal restatement of processCigar + CigarCluster.add.  When cov != nullptr the depth/error/start/end
// events go straight to the cluster's rows (cov points at [arr=0][prov][src][0]; arr_stride ints between arrays).
__device__ __noinline__ void serial_walk(const DevCfg& c, const uint32_t* refw, const BatchDev& b, ReadCtx& rc, WarpScratch* ws, int* cov,
                            long long arr_stride) {
  const bool store = cov == nullptr;  // the coverage pass must not disturb the (possibly trimmed) break list
  int nb = 0;
  int prev = -1;
  int readPos = 0, refPos = rc.alnStart - 1;
  if (store) ws->br[nb] = rc.alnStart;
  nb++;
  prev = rc.alnStart;
  int maps = 0, errs = 0;
  const int T = c.T, G = c.G;
  auto add = [&](int pos, bool match) {
    bool break_p = prev < 0 || pos - prev > T;
    if (c.record_depth) {
      maps++;
      if (!match) errs++;
      if (cov) {
        atomicAdd(cov + pos, 1);
        atomicAdd(cov + pos + 1, -1);
        if (!match) atomicAdd(cov + arr_stride + pos, 1);
        if (break_p && prev > 0) { atomicAdd(cov + 2 * arr_stride + pos, 1); atomicAdd(cov + 3 * arr_stride + prev, 1); }
      }
    }
    if (match) {
      if (break_p) {
        if (prev > 0) { if (store && nb < MAXB) ws->br[nb] = prev; nb++; }
        if (store && nb < MAXB) ws->br[nb] = pos;
        nb++;
      }
      prev = pos;
    }
  };
  for (uint32_t k = rc.c0; k < rc.c1; k++) {
    uint32_t w = __ldg(b.cigar + k);
    int len = (int)(w >> 4), op = (int)(w & 15);
    switch (op) {
      case 5: case 6: break;
      case 4: readPos += len; break;
      case 3: refPos += len; break;
      case 2:
        refPos += len;
        for (int i = 0; i < len && refPos + i < G; i++) add(refPos + i + 1, false);
        break;
      case 1: readPos += len; break;
      case 0:
        for (int i = 0; i < len && refPos + i < G; i++) {
          bool m = ref_base(refw, (uint32_t)(refPos + i)) == read_base(rc.wread, rc.nib0 + (uint32_t)(readPos + i));
          add(refPos + i + 1, m);
        }
        readPos += len; refPos += len;
        break;
      case 7:
        readPos += len; refPos += len;
        for (int i = 0; i < len && refPos + i < G; i++) add(refPos + i + 1, true);
        break;
      case 8:
        readPos += len; refPos += len;
        for (int i = 0; i < len && refPos + i < G; i++) add(refPos + i + 1, false);
        break;
      default: rc.bad = true; return;
    }
  }
  rc.alnEnd = refPos;  // alnStart + refSpan - 1
  if (store && nb < MAXB) ws->br[nb] = rc.alnEnd;
  nb++;
  rc.nb = nb;
  rc.maps_sum = maps; rc.err_sum = errs;
}

// htsjdk getReadPositionAtReferencePosition over alignment blocks (lane 0)
__device__ __noinline__ int serial_read_pos(const BatchDev& b, const ReadCtx& rc, int pos) {
  if (pos <= 0) return 0;
  int readBase = 1, refBase = rc.alnStart;
  for (uint32_t k = rc.c0; k < rc.c1; k++) {
    uint32_t w = __ldg(b.cigar + k);
    int len = (int)(w >> 4), op = (int)(w & 15);
    if (op == 4 || op == 1) readBase += len;
    else if (op == 3 || op == 2) refBase += len;
    else if (op == 0 || op == 7 || op == 8) {
      int bend = refBase + len - 1;
      if (bend >= pos) return pos < refBase ? 0 : pos - refBase + readBase;
      readBase += len; refBase += len;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ fast warp walker
// Per 32-op chunk: (1) lane = op: decode, warp scans give every op its reference/read offset; M ops publish a descriptor
// to shared memory.  (2) lane = 8-base window: the chunk's M bases are cut into windows dealt round-robin to the lanes, so
// the compare runs with (almost) all lanes busy whatever the op lengths are.  (3) lane = op again: depth runs as
// difference events.  All staging offsets are warp-uniform (ballot + popc), no shared atomics.
//
// Two tiers per chunk.  Tier 1 (hot): when every window holds at least one match, consecutive matches are at most 15
// apart inside an op, so with breakThresh >= 16 a break / mapStart / mapEnd event needs an op whose first possible match
// lies > T past the previous op's last possible match — a purely geometric test (ballot).  Chunks that pass it skip the
// per-op first/last-match machinery altogether.  Tier 2 (chunk_exact, not inlined): chunks with a candidate junction, a
// window without a match, an '=' op or a small breakThresh get the exact segmented first/last-match computation.
struct WalkOut { int nev, nmm, ovf; };
struct WalkState { int prev, nb, nev, ovf; bool slow; };

// sink 0: no coverage events; 1: shared-memory staging; 2: straight to the cluster's rows (also: do not touch ws->br)
__device__ __noinline__ void chunk_exact(const DevCfg& c, const uint32_t* refw, const ReadCtx& rc, WarpScratch* ws, const unsigned lane,
                                         const int sink, const CovSink sk, const int op, const int nE, const int base, const bool isM,
                                         const int mrank, const int W, const int wStart, WalkState& st) {
  const int T = c.T;
  const unsigned lt = lanemask_lt(), le = lanemask_le();
  if (isM) { ws->m_fm[mrank] = -1; ws->m_lm[mrank] = -1; }
  __syncwarp();
  bool slow = false;
  int headsBefore = 0;
  for (int r0 = 0; r0 < W; r0 += 32) {
    const int v = r0 + (int)lane;
    const bool wv = v < W;
    const unsigned hbit = (isM && wStart >= r0 && wStart < r0 + 32) ? 1u << (wStart - r0) : 0u;
    const unsigned heads = __reduce_or_sync(0xffffffffu, hbit);
    const int mr = max(0, headsBefore + __popc(heads & le) - 1);
    headsBefore += __popc(heads);
    uint32_t mis = 0, mat = 0;
    int j8 = 0;
    if (wv) {
      const int d_start0 = ws->m_start0[mr], d_rpos0 = ws->m_rpos0[mr], d_nE = ws->m_nE[mr];
      j8 = (v - ws->m_wstart[mr]) << 3;
      const int cnt = min(8, d_nE - j8);
      cmp8(read_win8(rc.wread, rc.nib0 + (uint32_t)(d_rpos0 + j8)), ref_win8(refw, (uint32_t)(d_start0 + j8)), cnt, mis, mat);
    }
    // per-op first / last match through segments of equal op
    const bool hasMat = mat != 0;
    const int f = j8 + (__clz(mat | 1u) >> 2);
    const int l = j8 + 7 - ((__ffs(mat | 0x80000000u) - 1) >> 2);
    const unsigned matMask = __ballot_sync(0xffffffffu, hasMat);
    const unsigned segHeads = heads | 1u;
    const int a = 31 - __clz(segHeads & le);
    const unsigned above = segHeads & ~le;
    const unsigned upto = above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu;
    const unsigned segMask = upto & ~((1u << a) - 1u);
    const unsigned sm = matMask & segMask;
    const int fmSeg = __shfl_sync(0xffffffffu, f, sm ? __ffs(sm) - 1 : (int)lane);
    const int lmSeg = __shfl_sync(0xffffffffu, l, sm ? 31 - __clz(sm) : (int)lane);
    const unsigned pm = sm & lt;
    const int lprev = __shfl_sync(0xffffffffu, l, pm ? 31 - __clz(pm) : (int)lane);
    if (hasMat) {
      const int pl = pm ? lprev : ws->m_lm[mr];  // last match of this op before this window (earlier rounds via shared)
      if (pl >= 0 && f - pl > T) slow = true;     // a gap > T inside one op: serial walker
    }
    __syncwarp();
    if (wv && (int)lane == a && sm) {
      if (ws->m_fm[mr] < 0) ws->m_fm[mr] = fmSeg;
      ws->m_lm[mr] = lmSeg;
    }
    __syncwarp();
  }
  int fmI = -1, lmI = -1;
  if (isM) { fmI = ws->m_fm[mrank]; lmI = ws->m_lm[mrank]; }
  if (op == 7 && nE > 0) { fmI = 0; lmI = nE - 1; }
  const bool has = fmI >= 0;
  const int fm = base + fmI, lm = base + lmI;
  // last matched position before this op (prev_position of CigarCluster.add at the op's first base)
  const unsigned hasMask = __ballot_sync(0xffffffffu, has);
  const unsigned lowerH = hasMask & lt;
  const int lmPrev = __shfl_sync(0xffffffffu, lm, lowerH ? 31 - __clz(lowerH) : (int)lane);
  const int prev_j = lowerH ? lmPrev : st.prev;
  // breaks: first match of the op further than T from the previous match
  const bool brk = has && (fm - prev_j > T);
  const unsigned bm = __ballot_sync(0xffffffffu, brk);
  if (bm) {
    const int o = st.nb + 2 * __popc(bm & lt);
    if (sink != 2 && brk && o + 1 < MAXB) { ws->br[o] = prev_j; ws->br[o + 1] = fm; }
    st.nb += 2 * __popc(bm);
  }
  if (hasMask) st.prev = __shfl_sync(0xffffffffu, lm, 31 - __clz(hasMask));
  if (__any_sync(0xffffffffu, slow)) st.slow = true;
  if (sink != 0) {
    // mapStart / mapEnd (CigarCluster.add :479-484): every emitted position further than T from the last match before it.
    // lead: positions up to and including the op's first match, against prev_j; trail: positions after the op's last match.
    for (int pass = 0; pass < 2; pass++) {
      int lo = 0, n = 0, pj = 0;
      if (nE > 0) {
        if (pass == 0) {
          const int hi = has ? fmI : nE - 1;
          lo = max(0, T + prev_j - base + 1);
          if (lo <= hi) n = hi - lo + 1;
          pj = prev_j;
        } else if (has) {
          lo = lmI + T + 1;
          if (lo <= nE - 1) n = nE - lo;
          pj = lm;
        }
      }
      unsigned msMask = __ballot_sync(0xffffffffu, n > 0);
      while (msMask) {
        const int Lm = __ffs(msMask) - 1;
        msMask &= msMask - 1;
        const int p0 = __shfl_sync(0xffffffffu, base + lo, Lm), nn = __shfl_sync(0xffffffffu, n, Lm);
        const int pp = __shfl_sync(0xffffffffu, pj, Lm);
        if (sink == 1) {
          if (st.nev + 2 * nn > MAXEV) st.ovf = 1;
          else for (int i = lane; i < nn; i += 32) { ws->ev[st.nev + 2 * i] = (EV_MSTART << 29) | (unsigned)(p0 + i); ws->ev[st.nev + 2 * i + 1] = (EV_MEND << 29) | (unsigned)pp; }
          st.nev += 2 * nn;
        } else {
          for (int i = lane; i < nn; i += 32) { atomicAdd(sk.cov + 2 * sk.arr_stride + p0 + i, 1); atomicAdd(sk.cov + 3 * sk.arr_stride + pp, 1); }
        }
      }
    }
  }
}

// Returns false when the read must be re-walked serially.  Breaks land in ws->br, coverage events in ws->ev / ws->mm_*.
__device__ __noinline__ bool fast_walk(const DevCfg& c, const uint32_t* refw, const BatchDev& b, ReadCtx& rc, WarpScratch* ws,
                                       const unsigned lane, const int sink, const CovSink sk, WalkOut& wo) {
  const int T = c.T, G = c.G;
  const bool depth = sink != 0;
  const bool tier1 = T >= 16;
  const unsigned lt = lanemask_lt(), le = lanemask_le();
  int refBase = rc.alnStart - 1, readBase = 0;
  WalkState st{rc.alnStart, 1, 0, 0, false};
  if (sink != 2 && lane == 0) ws->br[0] = rc.alnStart;
  int maps = 0, errs = 0;
  bool stFound = false; int st_r = 0;
  int kBs = 0, kRs = 0, kEnd = INT_MIN, lastBlkLane = -1;  // per-lane copy of the last chunk that held an alignment block
  int pendingEnd = -1;  // depth run end (exclusive) of the last emitting op, not yet written as an event
  int nmm = 0;

  for (uint32_t k0 = rc.c0; k0 < rc.c1; k0 += 32) {
    const uint32_t k = k0 + lane;
    const bool valid = k < rc.c1;
    const uint32_t w = valid ? __ldg(b.cigar + k) : 5u;
    const int len = (int)(w >> 4), op = (int)(w & 15);
    const bool isBlk = valid && (op == 0 || op == 7 || op == 8);
    // bad op -> the reference throws; zero-length M/=/X -> let the serial walker decide the alignment-block corner cases
    const unsigned odd = __ballot_sync(0xffffffffu, op > 8 || (isBlk && len == 0));
    if (odd) { if (__any_sync(0xffffffffu, op > 8)) { rc.bad = true; return true; } return false; }
    const int refAdv = ((0x18D >> op) & 1) ? len : 0;   // M D N = X
    const int readAdv = ((0x193 >> op) & 1) ? len : 0;  // M I S = X
    const int ri = warp_incl_scan(refAdv, lane), qi = warp_incl_scan(readAdv, lane);
    const int refPos0 = refBase + ri - refAdv, readPos0 = readBase + qi - readAdv;
    refBase += __shfl_sync(0xffffffffu, ri, 31);
    readBase += __shfl_sync(0xffffffffu, qi, 31);
    const bool emits = (0x185 >> op) & 1;  // M D = X
    const int start0 = (op == 0) ? refPos0 : refPos0 + len;  // 0-based index of the first emitted reference base
    const int nE = emits ? max(0, min(len, G - start0)) : 0;
    const int base = start0 + 1;
    const bool em = nE > 0;

    // ---- (2) balanced compare of the chunk's M bases
    const bool isM = (op == 0) && em;
    const int nW = isM ? (nE + 7) >> 3 : 0;
    const int wEnd = warp_incl_scan(nW, lane);
    const int W = __shfl_sync(0xffffffffu, wEnd, 31);
    const int wStart = wEnd - nW;
    const unsigned mMask = __ballot_sync(0xffffffffu, isM);
    const int mrank = __popc(mMask & lt);
    if (isM) { ws->m_start0[mrank] = start0; ws->m_rpos0[mrank] = readPos0; ws->m_nE[mrank] = nE; ws->m_wstart[mrank] = wStart; }
    __syncwarp();
    int headsBefore = 0;
    bool noMatch = false;
    uint32_t mat = 0; int j8 = 0, d_start0 = 0;  // of this lane's window in the last round (for the exact prev carry)
    for (int r0 = 0; r0 < W; r0 += 32) {
      const int v = r0 + (int)lane;
      const unsigned hbit = (isM && wStart >= r0 && wStart < r0 + 32) ? 1u << (wStart - r0) : 0u;
      const unsigned heads = __reduce_or_sync(0xffffffffu, hbit);
      const int mr = max(0, headsBefore + __popc(heads & le) - 1);
      headsBefore += __popc(heads);
      uint32_t mis = 0;
      mat = 0;
      if (v < W) {
        d_start0 = ws->m_start0[mr];
        const int d_rpos0 = ws->m_rpos0[mr], d_nE = ws->m_nE[mr];
        j8 = (v - ws->m_wstart[mr]) << 3;
        const int cnt = min(8, d_nE - j8);
        cmp8(read_win8(rc.wread, rc.nib0 + (uint32_t)(d_rpos0 + j8)), ref_win8(refw, (uint32_t)(d_start0 + j8)), cnt, mis, mat);
        noMatch |= mat == 0;
      }
      errs += __popc(mis);
      if (depth) {
        const unsigned hm = __ballot_sync(0xffffffffu, mis != 0);
        if (hm) {
          if (sink == 1) {
            const int o = nmm + __popc(hm & lt);
            if (mis && o < MAXMM) { ws->mm_pos[o] = (unsigned)(d_start0 + 1 + j8); ws->mm_mask[o] = mis; }
            nmm += __popc(hm);
          } else if (mis) {
            red_error_window(sk.cov + sk.arr_stride, (unsigned)(d_start0 + 1 + j8), mis);
          }
        }
      }
    }

    // ---- tier decision: can this chunk hold a break or a mapStart/mapEnd event?
    const unsigned lowerM = mMask & lt;
    const int E_j = base + nE;
    const int eprev = __shfl_sync(0xffffffffu, E_j, lowerM ? 31 - __clz(lowerM) : (int)lane);
    const int lb = lowerM ? eprev - 8 : st.prev;                       // the previous match is at or after this position
    const int ub = isM ? base + min(7, nE - 1) : base + nE - 1;        // the next match / far non-match is at or before this one
    const bool cand = em && (ub - lb > T);
    if (!tier1 || __any_sync(0xffffffffu, cand || noMatch || (op == 7 && valid))) {
      chunk_exact(c, refw, rc, ws, lane, sink, sk, op, nE, base, isM, mrank, W, wStart, st);
    } else if (W > 0) {
      // every window matched somewhere: the chunk's last match is the last match of its last window
      const int lpos = d_start0 + 1 + j8 + 7 - ((__ffs(mat | 0x80000000u) - 1) >> 2);
      st.prev = __shfl_sync(0xffffffffu, lpos, (W - 1) & 31);
    }

    // ---- (3) per-op stage: depth runs, D/X error windows, alignment-block bookkeeping
    const bool dx = (op == 2 || op == 8) && em;  // every emitted position of D / X is an error
    if (dx) errs += nE;
    maps += nE;
    if (depth) {
      // depth runs as difference events, cancelling "end of previous run == start of this run"
      const unsigned emMask = __ballot_sync(0xffffffffu, em);
      const unsigned lowerE = emMask & lt;
      const int pe_s = __shfl_sync(0xffffffffu, E_j, lowerE ? 31 - __clz(lowerE) : (int)lane);
      const int prevE = lowerE ? pe_s : pendingEnd;
      const bool emitStart = em && prevE != base;
      if (emMask) pendingEnd = __shfl_sync(0xffffffffu, E_j, 31 - __clz(emMask));
      const unsigned sMask = __ballot_sync(0xffffffffu, emitStart);
      if (sMask) {
        if (sink == 1) {
          // two slots per run start: [-1 at the previous run's end] [+1 at this run's start]; the very first run of a read has
          // no previous end and writes a +0 no-op (EV_DEPTH_INC at position 0 is never read: rows start at position 1 ... see flush)
          const int o = st.nev + 2 * __popc(sMask & lt);
          if (emitStart && o + 1 < MAXEV) {
            ws->ev[o] = prevE >= 0 ? ((EV_DEPTH_DEC << 29) | (unsigned)prevE) : (7u << 29);
            ws->ev[o + 1] = (EV_DEPTH_INC << 29) | (unsigned)base;
          }
          st.nev += 2 * __popc(sMask);
        } else if (emitStart) {
          if (prevE >= 0) atomicAdd(sk.cov + prevE, -1);
          atomicAdd(sk.cov + base, 1);
        }
      }
      const unsigned dxMask = __ballot_sync(0xffffffffu, dx);
      if (dxMask) {
        if (sink == 1) {
          if (__any_sync(0xffffffffu, dx && nE > 8)) st.ovf = 1;  // long deletions are rare: direct pass
          const int o = nmm + __popc(dxMask & lt);
          if (dx && nE <= 8 && o < MAXMM) { ws->mm_pos[o] = (unsigned)base; ws->mm_mask[o] = 0x11111111u & (0xFFFFFFFFu << ((8 - nE) << 2)); }
          nmm += __popc(dxMask);
        } else if (dx) {
          for (int i = 0; i < nE; i++) atomicAdd(sk.cov + sk.arr_stride + base + i, 1);
        }
      }
    }
    // read offsets of alignment start / end (htsjdk alignment blocks = M,=,X elements; all have len > 0 here)
    const unsigned blkMask = __ballot_sync(0xffffffffu, isBlk);
    if (blkMask) {
      const int bStart = refPos0 + 1, rStart = readPos0 + 1;
      if (!stFound) {
        const int l0 = __ffs(blkMask) - 1;
        const int bs = __shfl_sync(0xffffffffu, bStart, l0), rs = __shfl_sync(0xffffffffu, rStart, l0);
        st_r = (bs == rc.alnStart) ? rs : 0;  // first block: its end >= alnStart always; 0 if the read starts with D/N
        stFound = true;
      }
      kBs = bStart; kRs = rStart; kEnd = refPos0 + len; lastBlkLane = 31 - __clz(blkMask);
    }
  }
  __syncwarp();
  if (st.slow) return false;
  if (st.nb + 1 > MAXB) return false;  // let the serial path report it
  rc.alnEnd = refBase;
  if (sink != 2 && lane == 0) ws->br[st.nb] = rc.alnEnd;
  st.nb++;
  if (depth && pendingEnd >= 0) {
    if (sink == 1) { if (lane == 0 && st.nev < MAXEV) ws->ev[st.nev] = (EV_DEPTH_DEC << 29) | (unsigned)pendingEnd; st.nev++; }
    else if (lane == 0) atomicAdd(sk.cov + pendingEnd, -1);
  }
  if (sink == 1 && (st.nev > MAXEV - 2 || nmm > MAXMM)) st.ovf = 1;
  __syncwarp();
  wo.nev = st.nev; wo.nmm = nmm; wo.ovf = st.ovf;
  rc.nb = st.nb;
  rc.maps_sum = __reduce_add_sync(0xffffffffu, maps);
  rc.err_sum = __reduce_add_sync(0xffffffffu, errs);
  rc.st_r = st_r;
  int end_r = 0;
  if (lastBlkLane >= 0) {
    const int bs = __shfl_sync(0xffffffffu, kBs, lastBlkLane), rs = __shfl_sync(0xffffffffu, kRs, lastBlkLane);
    const int be = __shfl_sync(0xffffffffu, kEnd, lastBlkLane);
    if (rc.alnEnd > 0 && be == rc.alnEnd) end_r = rc.alnEnd - bs + rs;  // the last block reaches alnEnd unless the read ends with D/N
  }
  rc.end_r = end_r;
  return true;
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

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// order-sensitive 64-bit hash of a list held in shared memory, seeded
__device__ __forceinline__ unsigned long long list_hash(const int* v, int n, unsigned long long seed, unsigned lane) {
  unsigned lo = 0, hi = 0;
  for (int i = lane; i < n; i += 32) {
    unsigned long long h = mix64(((unsigned long long)(unsigned)v[i] << 32 | (unsigned)i) + seed * 0x9E3779B97F4A7C15ULL);
    lo += (unsigned)h; hi += (unsigned)(h >> 32);
  }
  lo = __reduce_add_sync(0xffffffffu, lo);
  hi = __reduce_add_sync(0xffffffffu, hi);
  unsigned long long h = mix64(((unsigned long long)hi << 32 | lo) ^ ((unsigned long long)n << 48) ^ seed);
  return h ? h : 1ULL;
}

// ------------------------------------------------------------------------------------------------ hash tables
// Cluster table: find or insert the token list ws->tok[0..ntok).  Warp-uniform; lane 0 owns the atomics.
// Returns the provisional dense id, or -1 (capacity error recorded in counters[CN_ERROR]).
__device__ int cluster_find_or_insert(const DevCfg& c, const int* tok, int ntok, unsigned long long h, unsigned long long idx, unsigned lane) {
  unsigned slot = (unsigned)h & c.cl_mask;
  for (unsigned probe = 0; probe <= c.cl_mask; probe++, slot = (slot + 1) & c.cl_mask) {
    ClEntry* e = c.cl + slot;
    int res = 2;  // 0 created, 1 same tag, 2 other tag
    if (lane == 0) {
      unsigned long long tag = ldcg(&e->tag);
      if (tag == 0) { tag = atomicCAS(&e->tag, 0ULL, h); if (tag == 0) res = 0; }
      if (res != 0 && tag == h) res = 1;
    }
    res = __shfl_sync(0xffffffffu, res, 0);
    if (res == 2) continue;
    if (res == 0) {
      int prov = -2; unsigned off = 0;
      if (lane == 0) {
        prov = (int)atomicAdd(c.counters + CN_CLUSTERS, 1u);
        off = atomicAdd(c.counters + CN_TOK_USED, (unsigned)ntok);
        if (prov >= c.cl_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_CL_FULL); prov = -2; }
        else if (off + (unsigned)ntok > c.tok_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_ARENA_FULL); prov = -2; }
      }
      prov = __shfl_sync(0xffffffffu, prov, 0);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (prov >= 0) {
        for (int i = lane; i < ntok; i += 32) c.cl_tok[off + i] = tok[i];
        __threadfence();
      }
      __syncwarp();
      if (lane == 0) {
        e->tok_off = off; e->ntok = (unsigned)ntok;
        atomicMin(&e->min_idx, idx);
        __threadfence();
        *(volatile int*)&e->prov = prov;
      }
      return prov >= 0 ? prov : -1;
    }
    // same tag: wait for the creator to publish, then verify the full key
    int prov = -1; unsigned off = 0, n = 0;
    if (lane == 0) {
      while ((prov = *(volatile int*)&e->prov) == -1) __nanosleep(40);
      __threadfence();
      off = ldcg(&e->tok_off); n = ldcg(&e->ntok);
    }
    prov = __shfl_sync(0xffffffffu, prov, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    if (prov < 0) return -1;
    bool eq = (int)n == ntok;
    if (eq) for (int i = lane; i < ntok; i += 32) eq = eq && (ldcg(c.cl_tok + off + i) == tok[i]);
    if (!__all_sync(0xffffffffu, eq)) continue;
    if (lane == 0 && idx < ldcg(&e->min_idx)) atomicMin(&e->min_idx, idx);
    return prov;
  }
  if (lane == 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_CL_FULL);
  return -1;
}

// Sub-cluster table: key = (cluster prov, binned list key[0..nkey)).  idxf = (global index << 2) | first-read flags.
__device__ int sub_find_or_insert(const DevCfg& c, int cl_prov, const int* key, int nkey, unsigned long long h, unsigned long long idxf,
                                  int nsum, unsigned lane) {
  unsigned slot = (unsigned)h & c.sb_mask;
  for (unsigned probe = 0; probe <= c.sb_mask; probe++, slot = (slot + 1) & c.sb_mask) {
    SubEntry* e = c.sb + slot;
    int res = 2;
    if (lane == 0) {
      unsigned long long tag = ldcg(&e->tag);
      if (tag == 0) { tag = atomicCAS(&e->tag, 0ULL, h); if (tag == 0) res = 0; }
      if (res != 0 && tag == h) res = 1;
    }
    res = __shfl_sync(0xffffffffu, res, 0);
    if (res == 2) continue;
    if (res == 0) {
      int prov = -2; unsigned off = 0, soff = 0;
      if (lane == 0) {
        prov = (int)atomicAdd(c.counters + CN_SUBS, 1u);
        off = atomicAdd(c.counters + CN_SUBKEY_USED, (unsigned)nkey);
        soff = atomicAdd(c.counters + CN_SUMS_USED, (unsigned)nsum);
        if (prov >= c.sb_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_SUB_FULL); prov = -2; }
        else if (off + (unsigned)nkey > c.sbkey_cap || soff + (unsigned)nsum > c.sums_cap) { atomicOr(c.counters + CN_ERROR, (unsigned)ERR_ARENA_FULL); prov = -2; }
      }
      prov = __shfl_sync(0xffffffffu, prov, 0);
      off = __shfl_sync(0xffffffffu, off, 0);
      soff = __shfl_sync(0xffffffffu, soff, 0);
      if (prov >= 0) {
        for (int i = lane; i < nkey; i += 32) c.sb_key[off + i] = key[i];
        if (lane == 0) { c.sub_sum_off[prov] = soff; c.sub_nsum[prov] = (unsigned)nsum; }
        __threadfence();
      }
      __syncwarp();
      if (lane == 0) {
        e->key_off = off; e->nkey = (unsigned)nkey; e->cl_prov = cl_prov;
        atomicMin(&e->min_idx, idxf);
        __threadfence();
        *(volatile int*)&e->prov = prov;
      }
      return prov >= 0 ? prov : -1;
    }
    int prov = -1; unsigned off = 0, n = 0; int cp = -1;
    if (lane == 0) {
      while ((prov = *(volatile int*)&e->prov) == -1) __nanosleep(40);
      __threadfence();
      off = ldcg(&e->key_off); n = ldcg(&e->nkey); cp = ldcg(&e->cl_prov);
    }
    prov = __shfl_sync(0xffffffffu, prov, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    cp = __shfl_sync(0xffffffffu, cp, 0);
    if (prov < 0) return -1;
    bool eq = (int)n == nkey && cp == cl_prov;
    if (eq) for (int i = lane; i < nkey; i += 32) eq = eq && (ldcg(c.sb_key + off + i) == key[i]);
    if (!__all_sync(0xffffffffu, eq)) continue;
    if (lane == 0 && idxf < ldcg(&e->min_idx)) atomicMin(&e->min_idx, idxf);
    return prov;
  }
  if (lane == 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_SUB_FULL);
  return -1;
}

// break-point matrix cell += 1   (BreakPoints.addBreakPoint, J/cluster/BreakPoints.java:102-108)
__device__ void pair_add(const DevCfg& c, int src, int level, int st, int en) {
  const unsigned long long key = (1ULL << 63) | ((unsigned long long)src << 59) | ((unsigned long long)level << 58) |
                                 ((unsigned long long)(unsigned)st << 29) | (unsigned long long)(unsigned)en;
  unsigned slot = (unsigned)mix64(key) & c.pr_mask;
  for (unsigned probe = 0; probe <= c.pr_mask; probe++, slot = (slot + 1) & c.pr_mask) {
    PairEntry* e = c.pr + slot;
    unsigned long long k = ldcg(&e->key);
    if (k == 0) {
      k = atomicCAS(&e->key, 0ULL, key);
      if (k == 0) {
        unsigned np = atomicAdd(c.counters + CN_PAIRS, 1u);
        if ((int)np >= c.pr_cap) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
        k = key;
      }
    }
    if (k == key) { atomicAdd(&e->count, 1); return; }
  }
  atomicOr(c.counters + CN_ERROR, (unsigned)ERR_PAIR_FULL);
}

// ------------------------------------------------------------------------------------------------ the kernel
// mode 0: full pass.  mode 1: re-emit breaks only at exact offsets (b.break_loc holds the exclusive scan; no side effects).
template <bool REF_SMEM>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 3) walk_kernel(const DevCfg c, const BatchDev b, const int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_ref = (uint32_t*)smem_raw;
  const int refw_n = REF_SMEM ? c.ref_nwords : 0;
  int* s_ostart = (int*)(s_ref + refw_n);
  int* s_oname = s_ostart + c.n_orf;
  int* s_ostrand = s_oname + c.n_orf;
  int* s_cnt = s_ostrand + c.n_orf;  // total_counts[S], spliced[S][n_orf], unspliced[S][n_orf]
  const int ncnt = c.S * (1 + 2 * c.n_orf);
  WarpScratch* s_ws = (WarpScratch*)(((uintptr_t)(s_cnt + ncnt) + 15) & ~(uintptr_t)15);

  for (int i = threadIdx.x; i < refw_n; i += blockDim.x) s_ref[i] = c.ref_words[i];
  for (int i = threadIdx.x; i < c.n_orf; i += blockDim.x) { s_ostart[i] = c.orf_start[i]; s_oname[i] = c.orf_name[i]; s_ostrand[i] = c.orf_strand[i]; }
  for (int i = threadIdx.x; i < ncnt; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t* refw = REF_SMEM ? s_ref : c.ref_words;

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  WarpScratch* ws = s_ws + warp;
  const long long nwarps = (long long)gridDim.x * WARPS_PER_BLOCK;
  const long long total_words = (b.seq_bytes + 3) >> 2;

  for (long long r = (long long)blockIdx.x * WARPS_PER_BLOCK + warp; r < b.n; r += nwarps) {
    if (mode == 1 && b.nbreaks[r] <= BRK_INLINE) continue;
    ReadCtx rc;
    rc.alnStart = __ldg(b.pos + r);
    rc.src = __ldg(b.src + r);
    rc.neg = (__ldg(b.flag + r) & 0x10) ? 1 : 0;
    rc.c0 = __ldg(b.cigar_off + r); rc.c1 = __ldg(b.cigar_off + r + 1);
    const unsigned long long so = __ldg(b.seq_off + r);
    rc.wread = (const uint32_t*)b.seq4 + (so >> 2);
    rc.nib0 = (uint32_t)(so & 3) * 2;
    const long long lw = total_words - 1 - (long long)(so >> 2);
    rc.lastw = (uint32_t)max(0LL, min(lw, 0xFFFFFFFFLL));
    rc.bad = false; rc.nb = 0; rc.maps_sum = 0; rc.err_sum = 0; rc.st_r = 0; rc.end_r = 0; rc.alnEnd = 0;
    __syncwarp();

    // coverage plan: short reads stage their events in shared memory during the walk; reads that could overflow the
    // staging buffer are walked without events first and a second time, straight into the cluster's rows, once the
    // cluster is known (their CIGAR and bases are L1/L2-resident by then).
    const CovSink nosink{nullptr, 0};
    WalkOut wo{0, 0, 0};
    bool direct_cov = c.record_depth && (rc.c1 - rc.c0) > 520u;
    if (rc.alnStart <= 0) rc.bad = true;  // not a mapped record
    bool serial = c.force_serial;
    if (!serial && !rc.bad) {
      if (c.record_depth && !direct_cov) {
        serial = !fast_walk(c, refw, b, rc, ws, lane, 1, nosink, wo);
        if (!serial && wo.ovf) direct_cov = true;
      } else {
        serial = !fast_walk(c, refw, b, rc, ws, lane, 0, nosink, wo);
      }
    }
    if (serial && !rc.bad) {
      if (lane == 0) {
        serial_walk(c, refw, b, rc, ws, nullptr, 0);
        if (!rc.bad) { rc.st_r = serial_read_pos(b, rc, rc.alnStart); rc.end_r = serial_read_pos(b, rc, rc.alnEnd); }
      }
      rc.bad = __shfl_sync(0xffffffffu, (int)rc.bad, 0) != 0;
      rc.nb = __shfl_sync(0xffffffffu, rc.nb, 0);
      rc.alnEnd = __shfl_sync(0xffffffffu, rc.alnEnd, 0);
      rc.maps_sum = __shfl_sync(0xffffffffu, rc.maps_sum, 0);
      rc.err_sum = __shfl_sync(0xffffffffu, rc.err_sum, 0);
      rc.st_r = __shfl_sync(0xffffffffu, rc.st_r, 0);
      rc.end_r = __shfl_sync(0xffffffffu, rc.end_r, 0);
      if (!rc.bad && rc.nb > MAXB) {
        if (lane == 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_TOO_MANY_BREAKS);
        rc.bad = true;
      }
    }
    __syncwarp();

    if (rc.bad) {
      if (lane == 0 && mode == 0) {
        b.cluster[r] = -1; b.sub[r] = -1; b.start_pos[r] = 0; b.end_pos[r] = 0; b.read_st[r] = 0; b.read_en[r] = 0;
        b.rflags[r] = 0x10u; b.maps_sum[r] = 0; b.err_sum[r] = 0; b.nbreaks[r] = 0; b.break_loc[r] = 0xFFFFFFFFu;
      }
      continue;
    }

    int nb = rc.nb;
    int* br = ws->br;
    // ---- small first/last exon removal (IdentityProfile1.java:205-228), host mode only
    if (c.min_fl_exon > 0 && nb > 2) {
      if (lane == 0) {
        const int m = c.min_fl_exon;
        const int diff0 = br[1] - br[0], diff1 = br[nb - 1] - br[nb - 2];
        int a = 0, e = nb;
        if (nb == 4 && diff0 < m && diff1 < m) { if (diff1 < diff0) e -= 2; else a += 2; }
        else { if (diff1 < m) e -= 2; if (diff0 < m) a += 2; }
        if (a) for (int i = a; i < e; i++) br[i - a] = br[i];
        nb = e - a;
      }
      nb = __shfl_sync(0xffffffffu, nb, 0);
      __syncwarp();
    }

    if (mode == 1) {  // re-emit of the long break lists (the first pass ran out of overflow area)
      if (nb > BRK_INLINE) {
        unsigned loc = 0;
        if (lane == 0) { loc = atomicAdd(b.bctr, (unsigned)nb); b.break_loc[r] = loc; }
        loc = __shfl_sync(0xffffffffu, loc, 0);
        if (loc + (unsigned)nb <= b.break_cap) for (int i = lane; i < nb; i += 32) b.break_arena[b.n * BRK_INLINE + loc + i] = br[i];
      }
      continue;
    }

    const int startPos = br[0], endPos = br[nb - 1];
    const bool fwd_known = (c.rna_mask >> rc.src) & 1;
    const bool fwd = !rc.neg;  // meaningful when fwd_known; in coronavirus mode callers guarantee rna[src]
    bool leader = false;
    if (c.coronavirus && nb > 1) leader = (c.annot_kind == 1) ? (br[1] < 100) : (br[1] < s_ostart[1] - c.bin);

    // ---- key tokens (IdentityProfile1.java:266-299, 325-327)
    int* tok = ws->tok;
    int ntok;
    if (c.coronavirus) {
      const bool ends = c.include_start || nb == 2;
      for (int k = lane; k < nb; k += 32) {
        const bool isEnd = (k == 0 || k == nb - 1);
        if (isEnd && !ends) { if (k == 0) tok[0] = -5; continue; }  // NPT_TOK_NOENDS marker
        const bool up = isEnd || (k & 1);
        tok[k] = up ? tok_up(c, s_ostart, s_oname, s_ostrand, br[k], fwd) : tok_down(c, s_ostart, s_oname, s_ostrand, br[k], fwd);
      }
      ntok = ends ? nb : nb - 1;
    } else {
      if (lane == 0) tok[0] = tok_up(c, s_ostart, s_oname, s_ostrand, (fwd_known && !fwd) ? br[0] : br[nb - 1], fwd);
      ntok = 1;
    }
    if (c.enforce_strand) { if (lane == 0) tok[ntok] = fwd ? -3 : -4; ntok++; }
    __syncwarp();

    // ---- break-point pairs (:271-288): level 0 for the first internal gap > T of the read, 1 for the others
    if (c.coronavirus) {
      bool firstDone = false;
      const int npair = (nb - 2) / 2;
      for (int q0 = 0; q0 < npair; q0 += 32) {
        const int q = q0 + lane;
        bool big = false; int a = 0, e2 = 0;
        if (q < npair) { a = br[2 * q + 1]; e2 = br[2 * q + 2]; big = (e2 - a) > c.T; }
        const unsigned m = __ballot_sync(0xffffffffu, big);
        if (m) {
          leader = true;
          if (c.calc_breaks1 && big) {
            const int level = (!firstDone && (int)lane == __ffs(m) - 1) ? 0 : 1;
            pair_add(c, rc.src, level, a, e2);
          }
          firstDone = true;
        }
        __syncwarp();
      }
    }

    // ---- type (Annotation.getTypeInd :233-236)
    int type_ind = 0;
    if (c.coronavirus) type_ind = (startPos <= c.start_thresh) ? (endPos >= c.G - c.end_thresh ? 0 : 1) : (endPos >= c.G - c.end_thresh ? 2 : 3);

    // ---- ORF spliced/unspliced counters (:349-356), block-privatised in shared memory
    if (lane == 0) atomicAdd(&s_cnt[rc.src], 1);  // CigarClusters.totalCounts (CigarClusters.java:81)
    if (c.coronavirus) {
      const int st1 = br[nb - 2];
      const bool spliced = startPos <= c.start_thresh;
      bool stop = false;
      for (int hi = c.n_orf; hi > 0 && !stop; hi -= 32) {
        const int i = hi - 1 - (int)lane;  // lane 0 = highest row
        const bool below = i >= 0 && s_ostart[i] < st1 - 10;
        const unsigned bmask = __ballot_sync(0xffffffffu, below);
        const int firstBelow = bmask ? __ffs(bmask) - 1 : 32;  // lanes before it are counted
        if (i >= 0 && (int)lane < firstBelow && endPos > s_ostart[i] + 100)
          atomicAdd(&s_cnt[c.S + (spliced ? 0 : c.S * c.n_orf) + rc.src * c.n_orf + i], 1);
        stop = bmask != 0;
      }
    }

    // ---- cluster + sub-cluster
    const unsigned long long gidx = (unsigned long long)(b.idx_base + r);
    const unsigned long long hcl = list_hash(tok, ntok, 0x1234ULL, lane);
    const int cl_prov = cluster_find_or_insert(c, tok, ntok, hcl, gidx, lane);
    int sub_prov = -1;
    if (cl_prov >= 0) {
      // CigarCluster.cloneBreaks :620-631 — binned breaks, reuse tok[] as the key buffer
      int a = 0, e = nb;
      if (!c.include_start && fwd_known) { if (fwd) a = 1; else e = nb - 1; }
      __syncwarp();
      for (int i = a + lane; i < e; i += 32) tok[i - a] = br[i] / c.bin;
      __syncwarp();
      const int nkey = e - a;
      const unsigned long long hsb = list_hash(tok, nkey, 0x77ULL + (unsigned long long)(unsigned)cl_prov * 0x10001ULL, lane);
      const unsigned long long idxf = (gidx << 2) | (unsigned long long)((fwd ? 1 : 0) | (rc.neg ? 2 : 0));
      sub_prov = sub_find_or_insert(c, cl_prov, tok, nkey, hsb, idxf, c.track_sums ? nb : 0, lane);
      if (sub_prov >= 0) {
        if (lane == 0) {
          atomicAdd(c.sub_count + (size_t)sub_prov * c.S + rc.src, 1);
          if (c.track_sums) atomicAdd(c.sub_break_sum + sub_prov, 1);
        }
        if (c.track_sums) {
          unsigned soff = 0, ns = 0;
          if (lane == 0) { soff = ldcg(c.sub_sum_off + sub_prov); ns = ldcg(c.sub_nsum + sub_prov); }
          soff = __shfl_sync(0xffffffffu, soff, 0); ns = __shfl_sync(0xffffffffu, ns, 0);
          if ((int)ns == nb)
            for (int i = lane; i < nb; i += 32) atomicAdd(c.sub_sums + soff + i, (unsigned long long)(long long)br[i]);
        }
      }
    }

    // ---- coverage rows of the cluster
    if (c.record_depth && cl_prov >= 0) {
      if (cl_prov >= c.cov_cap) {
        if (lane == 0) atomicOr(c.counters + CN_ERROR, (unsigned)ERR_COV_FULL);
      } else {
        const long long arr_stride = (long long)c.cov_cap * c.S * c.stride;
        int* cov = c.cov + ((long long)cl_prov * c.S + rc.src) * c.stride;
        if (serial) {
          if (lane == 0) { ReadCtx rc2 = rc; serial_walk(c, refw, b, rc2, ws, cov, arr_stride); }
          __syncwarp();
        } else if (direct_cov) {
          ReadCtx rc2 = rc;
          WalkOut wo2{0, 0, 0};
          fast_walk(c, refw, b, rc2, ws, lane, 2, CovSink{cov, arr_stride}, wo2);
        } else {
          for (int i = lane; i < wo.nev; i += 32) {
            const unsigned ev = ws->ev[i];
            const unsigned t = ev >> 29, p = ev & EV_POS_MASK;
            if (t == 7u) continue;  // filler slot of a read's first depth run
            const int arr = t <= EV_DEPTH_DEC ? 0 : (int)t - 1;
            atomicAdd(cov + arr * arr_stride + p, t == EV_DEPTH_DEC ? -1 : 1);
          }
          for (int i = lane; i < wo.nmm; i += 32) red_error_window(cov + arr_stride, ws->mm_pos[i], ws->mm_mask[i]);
        }
        // CigarCluster.setStartEnd :360-365 (uses the trimmed startPos/endPos)
        // (positions beyond seqlen+1 only arise from alignments running past the reference end; dense rows drop them)
        if (lane == 0 && startPos >= 0 && startPos < c.stride) atomicAdd(cov + 2 * arr_stride + startPos, 1);
        if (lane == 1 && endPos >= 0 && endPos < c.stride) atomicAdd(cov + 3 * arr_stride + endPos, 1);
      }
    }

    // ---- per-read outputs
    unsigned loc = 0xFFFFFFFFu;
    if (nb > BRK_INLINE) {
      if (lane == 0) {
        loc = atomicAdd(b.bctr, (unsigned)nb);
        if (loc + (unsigned)nb > b.break_cap) { atomicOr(b.bctr + 1, 1u); loc = 0xFFFFFFFEu; }
      }
      loc = __shfl_sync(0xffffffffu, loc, 0);
      if (loc < 0xFFFFFFFEu) for (int i = lane; i < nb; i += 32) b.break_arena[b.n * BRK_INLINE + loc + i] = br[i];
    } else if ((int)lane < nb) {
      b.break_arena[r * BRK_INLINE + lane] = br[lane];
    }
    if (lane == 0) {
      b.cluster[r] = cl_prov; b.sub[r] = sub_prov; b.start_pos[r] = startPos; b.end_pos[r] = endPos;
      b.read_st[r] = rc.st_r; b.read_en[r] = rc.end_r;
      b.rflags[r] = (unsigned)type_ind | (leader ? 4u : 0u) | (rc.neg ? 8u : 0u) | (serial ? 0x20u : 0u);
      b.maps_sum[r] = c.record_depth ? rc.maps_sum : 0; b.err_sum[r] = c.record_depth ? rc.err_sum : 0; b.nbreaks[r] = nb; b.break_loc[r] = loc;
      if (serial) atomicAdd(c.counters + CN_SLOW, 1u);
    }
    __syncwarp();
  }

  __syncthreads();
  if (mode == 0)
    for (int i = threadIdx.x; i < ncnt; i += blockDim.x) if (s_cnt[i]) atomicAdd(c.small + i, s_cnt[i]);
}

}  // namespace npt
