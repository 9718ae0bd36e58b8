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
constexpr int MAXEV = 832;      // coverage events per read staged in shared memory
constexpr int LONGOP = 64;      // M ops longer than this are compared by the whole warp
constexpr int BRK_INLINE = 6;   // breaks stored in the per-read inline slot

constexpr unsigned EV_DEPTH_INC = 0, EV_DEPTH_DEC = 1, EV_ERR = 2, EV_MSTART = 3, EV_MEND = 4;
constexpr unsigned EV_POS_MASK = (1u << 29) - 1;

struct WarpScratch {
  int nev;
  int ev_ovf;
  int br[MAXB];
  int tok[MAXB + 2];
  unsigned ev[MAXEV];
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

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
// 8 read bases from the BAM-packed byte stream (little-endian word loads, byte-swapped to base order)
__device__ __forceinline__ uint32_t read_win8(const uint32_t* w, uint32_t nib, uint32_t lastw) {
  uint32_t k = nib >> 3, o = (nib & 7) << 2;
  uint32_t a = __byte_perm(__ldg(w + min(k, lastw)), 0, 0x0123);
  uint32_t b = __byte_perm(__ldg(w + min(k + 1, lastw)), 0, 0x0123);
  return __funnelshift_l(b, a, o);
}
__device__ __forceinline__ int read_base(const uint32_t* w, uint32_t nib, uint32_t lastw) { return (int)(read_win8(w, nib, lastw) >> 28); }
__device__ __forceinline__ int ref_base(const uint32_t* w, uint32_t nib) { return (int)((w[nib >> 3] >> (28 - ((nib & 7) << 2))) & 0xF); }

// mismatch / match masks of an 8-base window (bit 28-4j <-> base j), cnt valid bases
__device__ __forceinline__ void cmp8(uint32_t rw, uint32_t gw, int cnt, uint32_t& mis, uint32_t& mat) {
  uint32_t x = rw ^ gw;
  uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
  uint32_t vm = 0x11111111u & (0xFFFFFFFFu << ((8 - cnt) << 2));
  mis = nz & vm;
  mat = vm & ~nz;
}

__device__ __forceinline__ int ev_reserve(WarpScratch* ws, int n) {
  int o = atomicAdd(&ws->nev, n);
  if (o + n > MAXEV) { ws->ev_ovf = 1; return -1; }
  return o;
}

// Where fast_walk sends coverage events.  SINK 0: nowhere (depth off, or first pass of a long read);
// SINK 1: shared-memory staging (flushed once the cluster is known); SINK 2: straight to the cluster's rows.
struct CovSink { int* cov; long long arr_stride; };
template <int SINK>
__device__ __forceinline__ void emit_mismatches(WarpScratch* ws, const CovSink& sk, uint32_t mis, int pos0) {
  if (SINK == 1) {
    int o = ev_reserve(ws, __popc(mis));
    if (o >= 0) {
      uint32_t m = mis;
      while (m) { int bit = 31 - __clz(m); m &= ~(1u << bit); ws->ev[o++] = (EV_ERR << 29) | (unsigned)(pos0 + ((28 - bit) >> 2)); }
    }
  } else if (SINK == 2) {
    uint32_t m = mis;
    while (m) { int bit = 31 - __clz(m); m &= ~(1u << bit); atomicAdd(sk.cov + sk.arr_stride + pos0 + ((28 - bit) >> 2), 1); }
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
__device__ void serial_walk(const DevCfg& c, const uint32_t* refw, const BatchDev& b, ReadCtx& rc, WarpScratch* ws, int* cov,
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
          bool m = ref_base(refw, (uint32_t)(refPos + i)) == read_base(rc.wread, rc.nib0 + (uint32_t)(readPos + i), rc.lastw);
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
__device__ int serial_read_pos(const BatchDev& b, const ReadCtx& rc, int pos) {
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
// Returns false when the read must be re-walked serially.  Breaks land in ws->br, coverage events in ws->ev.
template <bool REF_SMEM, int SINK>
__device__ bool fast_walk(const DevCfg& c, const uint32_t* refw, const BatchDev& b, ReadCtx& rc, WarpScratch* ws, unsigned lane,
                          const CovSink sk) {
  const int T = c.T, G = c.G;
  constexpr bool depth = SINK != 0;
  constexpr bool store_br = SINK != 2;  // the direct coverage pass must not disturb the (possibly trimmed) break list
  const unsigned lt = lanemask_lt();
  int refBase = rc.alnStart - 1, readBase = 0;
  int prev = rc.alnStart;
  int nb = 1;
  if (store_br && lane == 0) ws->br[0] = rc.alnStart;
  int maps = 0, errs = 0;
  bool slow = false;
  bool stFound = false; int st_r = 0;
  int endMax = INT_MIN, endBs = 0, endRs = 0;
  int pendingEnd = -1;  // depth run end (exclusive) of the last emitting op, not yet written as an event

  for (uint32_t k0 = rc.c0; k0 < rc.c1; k0 += 32) {
    const uint32_t k = k0 + lane;
    const bool valid = k < rc.c1;
    const uint32_t w = valid ? __ldg(b.cigar + k) : 5u;
    const int len = (int)(w >> 4), op = (int)(w & 15);
    if (__any_sync(0xffffffffu, op > 8)) { rc.bad = true; return true; }
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

    int fmI = -1, lmI = -1, nmis = 0;
    const bool isM = (op == 0) && nE > 0;
    if (isM && nE <= LONGOP) {
      for (int j8 = 0; j8 < nE; j8 += 8) {
        const int cnt = min(8, nE - j8);
        const uint32_t rw = read_win8(rc.wread, rc.nib0 + (uint32_t)(readPos0 + j8), rc.lastw);
        const uint32_t gw = ref_win8(refw, (uint32_t)(start0 + j8));
        uint32_t mis, mat;
        cmp8(rw, gw, cnt, mis, mat);
        nmis += __popc(mis);
        if (mat) {
          const int f = j8 + (__clz(mat) >> 2);
          if (fmI < 0) fmI = f; else if (f - lmI > T) slow = true;
          lmI = j8 + 7 - ((__ffs(mat) - 1) >> 2);
        }
        if (depth && mis) emit_mismatches<SINK>(ws, sk, mis, base + j8);
      }
    }
    __syncwarp();
    unsigned longMask = __ballot_sync(0xffffffffu, isM && nE > LONGOP);
    while (longMask) {
      const int L = __ffs(longMask) - 1;
      longMask &= longMask - 1;
      const int startL = __shfl_sync(0xffffffffu, start0, L), rposL = __shfl_sync(0xffffffffu, readPos0, L);
      const int nEL = __shfl_sync(0xffffffffu, nE, L);
      int fmL = -1, lmL = -1, nmisL = 0;
      for (int jb = 0; jb < nEL; jb += 256) {
        const int j8 = jb + (int)lane * 8;
        const int cnt = max(0, min(8, nEL - j8));
        uint32_t mis = 0, mat = 0;
        if (cnt > 0) {
          const uint32_t rw = read_win8(rc.wread, rc.nib0 + (uint32_t)(rposL + j8), rc.lastw);
          const uint32_t gw = ref_win8(refw, (uint32_t)(startL + j8));
          cmp8(rw, gw, cnt, mis, mat);
        }
        nmisL += __popc(mis);
        if (depth && mis) emit_mismatches<SINK>(ws, sk, mis, startL + 1 + j8);
        __syncwarp();
        const unsigned mm = __ballot_sync(0xffffffffu, mat != 0);
        if (mm) {
          const int f = j8 + (__clz(mat | 1u) >> 2);          // meaningful only where mat != 0
          const int l = j8 + 7 - ((__ffs(mat | 0x80000000u) - 1) >> 2);
          const unsigned lower = mm & lt;
          const int pl_s = __shfl_sync(0xffffffffu, l, lower ? 31 - __clz(lower) : (int)lane);
          const int pl = lower ? pl_s : lmL;
          if (mat && pl >= 0 && f - pl > T) slow = true;
          const int f0 = __shfl_sync(0xffffffffu, f, __ffs(mm) - 1);
          if (fmL < 0) fmL = f0;
          lmL = __shfl_sync(0xffffffffu, l, 31 - __clz(mm));
        }
      }
      nmisL = __reduce_add_sync(0xffffffffu, nmisL);
      if ((int)lane == L) { fmI = fmL; lmI = lmL; nmis = nmisL; }
    }
    if (op == 7 && nE > 0) { fmI = 0; lmI = nE - 1; }
    if (op == 2 || op == 8) nmis = nE;
    const bool has = fmI >= 0;
    const int fm = base + fmI, lm = base + lmI;

    // last matched position before this op (prev_position of CigarCluster.add at the op's first base)
    const unsigned hasMask = __ballot_sync(0xffffffffu, has);
    const unsigned lowerH = hasMask & lt;
    const int lmPrev = __shfl_sync(0xffffffffu, lm, lowerH ? 31 - __clz(lowerH) : (int)lane);
    const int prev_j = lowerH ? lmPrev : prev;

    // breaks: first match of the op further than T from the previous match
    const bool brk = has && (fm - prev_j > T);
    const unsigned bm = __ballot_sync(0xffffffffu, brk);
    if (bm) {
      const int o = nb + 2 * __popc(bm & lt);
      if (store_br && brk && o + 1 < MAXB) { ws->br[o] = prev_j; ws->br[o + 1] = fm; }
      nb += 2 * __popc(bm);
    }
    if (hasMask) prev = __shfl_sync(0xffffffffu, lm, 31 - __clz(hasMask));
    maps += nE; errs += nmis;

    if (depth) {
      // depth runs as difference events, cancelling "end of previous run == start of this run"
      const bool em = nE > 0;
      const int S_j = base, E_j = base + nE;
      const unsigned emMask = __ballot_sync(0xffffffffu, em);
      const unsigned lowerE = emMask & lt;
      const int pe_s = __shfl_sync(0xffffffffu, E_j, lowerE ? 31 - __clz(lowerE) : (int)lane);
      const int prevE = lowerE ? pe_s : pendingEnd;
      const bool emitStart = em && prevE != S_j;
      int ne = emitStart ? (prevE >= 0 ? 2 : 1) : 0;
      // D / X ops: every emitted position is an error
      const int nerr = (op == 2 || op == 8) ? nE : 0;
      // mapStart / mapEnd: emitted positions up to and including the first match that are > T past prev_j
      int ms_lo = 0, ms_n = 0;
      if (em) {
        const int hi = has ? fmI : nE - 1;
        const int lo = max(0, T + prev_j - base + 1);
        if (lo <= hi) { ms_lo = lo; ms_n = hi - lo + 1; }
      }
      const int tot = ne + nerr + 2 * ms_n;
      if (SINK == 1) {
        if (tot) {
          int o = ev_reserve(ws, tot);
          if (o >= 0) {
            if (emitStart) {
              if (prevE >= 0) ws->ev[o++] = (EV_DEPTH_DEC << 29) | (unsigned)prevE;
              ws->ev[o++] = (EV_DEPTH_INC << 29) | (unsigned)S_j;
            }
            for (int i = 0; i < nerr; i++) ws->ev[o++] = (EV_ERR << 29) | (unsigned)(base + i);
            for (int i = 0; i < ms_n; i++) {
              ws->ev[o++] = (EV_MSTART << 29) | (unsigned)(base + ms_lo + i);
              ws->ev[o++] = (EV_MEND << 29) | (unsigned)prev_j;
            }
          }
        }
      } else if (SINK == 2) {
        if (emitStart) {
          if (prevE >= 0) atomicAdd(sk.cov + prevE, -1);
          atomicAdd(sk.cov + S_j, 1);
        }
        for (int i = 0; i < nerr; i++) atomicAdd(sk.cov + sk.arr_stride + base + i, 1);
        for (int i = 0; i < ms_n; i++) {
          atomicAdd(sk.cov + 2 * sk.arr_stride + base + ms_lo + i, 1);
          atomicAdd(sk.cov + 3 * sk.arr_stride + prev_j, 1);
        }
      }
      if (emMask) pendingEnd = __shfl_sync(0xffffffffu, E_j, 31 - __clz(emMask));
    }

    // read offsets of alignment start / end (htsjdk alignment blocks = M,=,X elements)
    const bool isBlk = valid && (op == 0 || op == 7 || op == 8);
    const int bStart = refPos0 + 1, bEnd = refPos0 + len, rStart = readPos0 + 1;
    if (!stFound) {
      const unsigned m = __ballot_sync(0xffffffffu, isBlk && bEnd >= rc.alnStart);
      if (m) {
        const int l = __ffs(m) - 1;
        const int bs = __shfl_sync(0xffffffffu, bStart, l), rs = __shfl_sync(0xffffffffu, rStart, l);
        st_r = rc.alnStart < bs ? 0 : rc.alnStart - bs + rs;
        stFound = true;
      }
    }
    {
      const int v = isBlk ? bEnd : INT_MIN;
      const int mx = __reduce_max_sync(0xffffffffu, v);
      if (mx > endMax) {
        endMax = mx;
        const int l = __ffs(__ballot_sync(0xffffffffu, isBlk && bEnd == mx)) - 1;
        endBs = __shfl_sync(0xffffffffu, bStart, l);
        endRs = __shfl_sync(0xffffffffu, rStart, l);
      }
    }
  }
  __syncwarp();
  if (__any_sync(0xffffffffu, slow)) return false;
  if (nb + 1 > MAXB) return false;  // let the serial path report it
  rc.alnEnd = refBase;
  if (store_br && lane == 0) ws->br[nb] = rc.alnEnd;
  nb++;
  if (depth && pendingEnd >= 0 && lane == 0) {
    if (SINK == 1) {
      int o = ev_reserve(ws, 1);
      if (o >= 0) ws->ev[o] = (EV_DEPTH_DEC << 29) | (unsigned)pendingEnd;
    } else if (SINK == 2) {
      atomicAdd(sk.cov + pendingEnd, -1);
    }
  }
  __syncwarp();
  rc.nb = nb;
  rc.maps_sum = __reduce_add_sync(0xffffffffu, maps);
  rc.err_sum = __reduce_add_sync(0xffffffffu, errs);
  rc.st_r = (rc.alnStart <= 0) ? 0 : st_r;
  rc.end_r = (rc.alnEnd > 0 && endMax == rc.alnEnd) ? (rc.alnEnd < endBs ? 0 : rc.alnEnd - endBs + endRs) : 0;
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
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) walk_kernel(const DevCfg c, const BatchDev b, const int mode) {
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
    if (lane == 0) { ws->nev = 0; ws->ev_ovf = 0; }
    __syncwarp();

    // coverage plan: short reads stage their events in shared memory during the walk; reads that could overflow the
    // staging buffer are walked without events first and a second time, straight into the cluster's rows, once the
    // cluster is known (their CIGAR and bases are L1/L2-resident by then).
    const CovSink nosink{nullptr, 0};
    bool direct_cov = c.record_depth && (rc.c1 - rc.c0) * 2u > (unsigned)MAXEV;
    bool serial = c.force_serial || rc.alnStart <= 0;
    if (!serial) {
      if (c.record_depth && !direct_cov) {
        serial = !fast_walk<REF_SMEM, 1>(c, refw, b, rc, ws, lane, nosink);
        if (!serial && *(volatile int*)&ws->ev_ovf) direct_cov = true;
      } else {
        serial = !fast_walk<REF_SMEM, 0>(c, refw, b, rc, ws, lane, nosink);
      }
    }
    if (serial && !rc.bad) {
      if (lane == 0) {
        ws->nev = 0; ws->ev_ovf = 0;
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
          fast_walk<REF_SMEM, 2>(c, refw, b, rc2, ws, lane, CovSink{cov, arr_stride});
        } else {
          const int nev = *(volatile int*)&ws->nev;
          for (int i = lane; i < nev; i += 32) {
            const unsigned ev = ws->ev[i];
            const unsigned t = ev >> 29, p = ev & EV_POS_MASK;
            const int arr = t <= EV_DEPTH_DEC ? 0 : (int)t - 1;
            atomicAdd(cov + arr * arr_stride + p, t == EV_DEPTH_DEC ? -1 : 1);
          }
        }
        // CigarCluster.setStartEnd :360-365 (uses the trimmed startPos/endPos)
        if (lane == 0) atomicAdd(cov + 2 * arr_stride + startPos, 1);
        if (lane == 1) atomicAdd(cov + 3 * arr_stride + endPos, 1);
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
