// fastwalk.cuh — the single-pass fast walk (one THREAD per read, two uniform phases per round).
//
// What it replaces: walk_kernel<0> + walk_kernel<1> of kernels.cuh for every read whose CIGAR is made of M I D N S H P
// only (minimap2's default output).  Reads with '=' / 'X' / invalid ops, very long deletions, alignments that leave
// the reference, more than 3 aligned blocks or more than 6 breaks are flagged NPT_RF_SLOWPATH and go through the
// exact serial device walker of kernels.cuh, which handles every case of the reference literally.
//
// Reference semantics restated here (J = /root/reference/java/src/main/java/npTranscript):
//   IdentityProfile1.processCigar      J/cluster/IdentityProfile1.java:482-585   (M emits pos refPos+i+1 with ref==read;
//                                      D advances refPos by len and THEN emits len mismatching positions, :513-522)
//   CigarCluster.add/addStart/addEnd   J/cluster/CigarCluster.java:442-498      (prev = last matching position; a match
//                                      more than break_thresh after prev appends (prev, pos) to breaks; with
//                                      recordStartEnd every emission more than break_thresh after prev counts
//                                      mapStart[pos]++ / mapEnd[prev]++)
//   SAMRecord geometry (htsjdk)        J/cluster/IdentityProfile1.java:581, 634-635
//
// Design (why it looks like this): the round-1 kernel ran a fixed four-phase loop per CIGAR element with every lane
// active in ~60 % of the phases (ncu: 1 700-2 400 warp-instructions per read, 19 of 32 lanes).  Here each lane first
// PRODUCES up to FW_K uniform tasks from its CIGAR (a task = at most 16 emitted positions: a piece of an M element, or
// the positions a D element emits), then every lane CONSUMES its tasks with the same straight-line code:
// normalise 16 read bases and 16 reference bases to the top of a 64-bit word, XOR, nibble-nonzero test.
// Coverage is not credited with atomics at all: the per-position flags (covered, deletion-emitted, mismatch) are ORed
// into a lane-private ring of 4-bit-per-position words in shared memory and leave the kernel as 16-byte ENTRIES
// (32 positions each, on the reference grid).  Once the cluster of the read is known, accumulate_kernel adds the
// entries of all reads of a (cluster, source) group with bit-sliced carry-save adders (fastcov.cuh) — positional
// popcount instead of 250 scattered REDs per read.
#pragma once
#if !defined(__CUDACC__)
// host build (tests/host_fastwalk.cpp runs the lane logic on the CPU against the oracle): the CUDA vector types
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
#endif
#include "npt_device.cuh"

#if defined(__CUDACC__)
#define FW_HD __host__ __device__ __forceinline__
#else
#define FW_HD inline
#endif

namespace npt {

constexpr int FW_TASK_MAX = 16;    // positions per task
constexpr int FW_K = 8;            // tasks per lane per round
constexpr int FW_RING = 32;        // ring of 8-position words per lane = 256 positions (live span of a round <= 175)
constexpr int FW_CIG_TILE = 16;    // CIGAR words staged per lane per round
constexpr int FW_SEQ_TILE = 20;    // base words (8 bases each) staged per lane per round: 160 bases
constexpr int FW_TILE_STRIDE = 20; // words between lane tiles (16-byte aligned, 4-way bank conflicts at worst)
constexpr int FW_MAX_BLOCKS = 3;   // aligned blocks (runs between N elements) per read
constexpr int FW_MAX_JUNC = 2;     // breaks (prev, pos) per read: BRK_INLINE = 6 = start + 2 pairs + end
constexpr int FW_MAX_EXTRA = 2;    // rare extra events per read (see FwHdr)
constexpr int FW_HDR_WORDS = 16;   // per-read header, uint32 words
constexpr int FW_BRK_INLINE = 6;
constexpr unsigned FW_TASK_D = 1u << 27;
constexpr unsigned FW_NIB_MAX = 1u << 27;  // read bases addressable by a task

// entry flags, one nibble per position (position 8W + i of ring word W sits in nibble i counted from the top)
constexpr unsigned FW_COV = 0x11111111u;   // emitted by an M element (depth +1)
constexpr unsigned FW_QRK = 0x22222222u;   // emitted by a D element after it advanced (depth +1, error +1)
constexpr unsigned FW_MIS = 0x88888888u;   // M position whose base differs from the reference (error +1)

// per-read header written by the fast walk, read by accumulate_kernel and fast_events_kernel:
//   [0] index of the read's first entry in the batch's entry arena
//   [1] nblk | njunc << 8 | nextra << 16
//   [2+2k], [3+2k]   block k: first entry index on the reference grid (position >> 5), number of entries
//   [8+2j], [9+2j]   junction j: (prev, pos) = mapEnd[prev]++, mapStart[pos]++
//   [12+2e],[13+2e]  extra e: kind 0 (bit 31 clear) = emission `pos` more than break_thresh after `prev` that was not a
//                    match (mapStart[pos]++, mapEnd[prev]++); kind 1 (bit 31 set) = a position emitted by two D elements
//                    (depth[pos]++, errors[pos]++ on top of the entry's single flag)
enum : unsigned { FWF_ACTIVE = 1, FWF_DONE = 2, FWF_SLOW = 4, FWF_ENDREAD = 8, FWF_ENDBLOCK = 16, FWF_FIRSTREF = 32, FWF_LASTREFM = 64, FWF_BLOCKOPEN = 128 };

struct FwLane {
  long long r;                  // read index in the batch
  unsigned ck, cend, cs;        // CIGAR cursor / end (word indices in b.cigar); first word held by the CIGAR tile
  int alnStart, refPos, readPos, mRem, nbases;
  unsigned long long so;        // byte offset of the read's bases in b.seq4
  long long tb;                 // byte offset in b.seq4 of the base tile (16-byte aligned), -1 = nothing staged
  int st_r, lmRead0, lmLen;     // read offset of the first aligned base; last M element (for the read offset of the alignment end)
  unsigned flags;
  int nb, nblk, nent, njunc, nextra, errs, maps, nt;
  unsigned pmh, pml; int pp0;   // last match: the last set flag of (pmh:pml), a task's match mask, counted from position pp0
  int maxPos, flushE, blkFirstE;
  unsigned entBase, entCap;
};

// ---- small portable intrinsics -----------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
FW_HD unsigned fw_fl(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_l(lo, hi, s); }      // high word of (hi:lo) << s, s < 32
FW_HD unsigned fw_fr(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_r(lo, hi, s); }      // low word of (hi:lo) >> s, s < 32
FW_HD unsigned fw_frc(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_rc(lo, hi, s); }    // ... s clamped to 32
FW_HD unsigned fw_bswap(unsigned x) { return __byte_perm(x, 0, 0x0123); }
FW_HD int fw_popc(unsigned x) { return __popc(x); }
FW_HD int fw_clz(unsigned x) { return __clz(x); }
FW_HD int fw_ctz(unsigned x) { return __ffs(x) - 1; }
#else
FW_HD unsigned fw_fl(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
FW_HD unsigned fw_fr(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
FW_HD unsigned fw_frc(unsigned lo, unsigned hi, unsigned s) { return s >= 32 ? hi : fw_fr(lo, hi, s); }
FW_HD unsigned fw_bswap(unsigned x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }
FW_HD int fw_popc(unsigned x) { return __builtin_popcount(x); }
FW_HD int fw_clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
FW_HD int fw_ctz(unsigned x) { return x ? __builtin_ctz(x) : -1; }
#endif
FW_HD int fw_min(int a, int b) { return a < b ? a : b; }
FW_HD int fw_max(int a, int b) { return a > b ? a : b; }

// index (counted from the first position of the task) of the last / first set match flag of a task's mask
FW_HD int fw_last_idx(unsigned mh, unsigned ml) { return ml ? 8 + ((31 - fw_ctz(ml)) >> 2) : ((31 - fw_ctz(mh)) >> 2); }
FW_HD int fw_first_idx(unsigned mh, unsigned ml, int none) { return mh ? (fw_clz(mh) >> 2) : (ml ? 8 + (fw_clz(ml) >> 2) : none); }

// ---- start of a read ----------------------------------------------------------------------------------------------
FW_HD void fw_start_read(FwLane& L, const BatchDev& b, long long r) {
  L.r = r;
  L.flags = FWF_ACTIVE;
  L.alnStart = b.pos[r];
  L.ck = b.cigar_off[r]; L.cend = b.cigar_off[r + 1];
  L.cs = 0xFFFFFFF0u;   // no CIGAR tile
  L.so = b.seq_off[r];
  const unsigned long long se = b.seq_off[r + 1];
  const unsigned long long nb2 = 2ull * (se - L.so);
  L.nbases = (int)(nb2 < 0x7FFFFFFFull ? nb2 : 0x7FFFFFFFull);
  L.tb = -1;
  L.entBase = (unsigned)((L.so >> 4) + 8ull * (unsigned long long)r);
  L.entCap = (unsigned)(((se >> 4) + 8ull * (unsigned long long)(r + 1)) - ((L.so >> 4) + 8ull * (unsigned long long)r));
  L.refPos = L.alnStart - 1; L.readPos = 0; L.mRem = 0;
  L.st_r = 0; L.lmRead0 = 0; L.lmLen = 0;
  L.nb = 1; L.nblk = 0; L.nent = 0; L.njunc = 0; L.nextra = 0; L.errs = 0; L.maps = 0; L.nt = 0;
  L.pmh = 0x80000000u; L.pml = 0; L.pp0 = L.alnStart;   // CigarCluster.addStart :442-446: prev = alignment start
  L.maxPos = 0; L.flushE = 0; L.blkFirstE = 0;
  if (L.alnStart <= 0) L.flags |= FWF_SLOW;             // not a mapped record: the serial walker flags it
  else b.break_arena[r * FW_BRK_INLINE] = L.alnStart;
}

// ---- phase 1: CIGAR -> tasks ---------------------------------------------------------------------------------------
// cig[i] = CIGAR word (L.cs + i), i < cigWords.  tasks[i * TS] receives task i: x = read base index | D flag | (n-1) << 28,
// y = refPos before the first emitted position (the task emits positions y+1 .. y+n).
template <bool RECORD>
FW_HD void fw_produce(FwLane& L, const DevCfg& c, const uint32_t* cig, int cigWords, uint2* tasks, int TS) {
  L.nt = 0;
  if ((L.flags & (FWF_ACTIVE | FWF_SLOW)) != FWF_ACTIVE) return;
  const int tnibEnd = (int)(2 * (L.tb - (long long)L.so)) + FW_SEQ_TILE * 8;
  const int G = c.G;
  for (;;) {
    if (L.mRem > 0) {
      if (L.readPos + FW_TASK_MAX + 8 > tnibEnd) break;   // bases not staged: next round
      const int n = fw_min(L.mRem, FW_TASK_MAX);
      uint2 t; t.x = (unsigned)L.readPos | ((unsigned)(n - 1) << 28); t.y = (unsigned)L.refPos;
      tasks[L.nt * TS] = t;
      L.refPos += n; L.readPos += n; L.mRem -= n; L.nt++;
      if (L.nt == FW_K) break;
      continue;
    }
    if (L.ck == L.cend) { L.flags |= FWF_ENDREAD; break; }
    if (L.ck - L.cs >= (unsigned)cigWords) break;
    const unsigned w = cig[L.ck - L.cs];
    const int op = (int)(w & 15u), len = (int)(w >> 4);
    if (op == 0) {          // M: emits from its own start (:531-551)
      if (len == 0 || L.readPos + len > L.nbases || L.refPos + len > G || (unsigned)(L.readPos + len) >= FW_NIB_MAX) { L.flags |= FWF_SLOW; break; }
      if (!(L.flags & FWF_FIRSTREF)) { L.flags |= FWF_FIRSTREF; L.st_r = L.readPos + 1; }
      if (!(L.flags & FWF_BLOCKOPEN)) {
        if (L.nblk == FW_MAX_BLOCKS) { L.flags |= FWF_SLOW; break; }
        L.flags |= FWF_BLOCKOPEN; L.flushE = L.blkFirstE = (L.refPos + 1) >> 5; L.maxPos = L.refPos + 1;
      }
      L.lmRead0 = L.readPos; L.lmLen = len; L.flags |= FWF_LASTREFM;
      L.mRem = len; L.maps += len; L.ck++;
      L.maxPos = fw_max(L.maxPos, L.refPos + len);
      continue;
    }
    if (op == 2) {          // D: advance, then emit len mismatching positions (:513-522)
      if (len == 0 || len > FW_TASK_MAX || !(L.flags & FWF_FIRSTREF) || L.refPos + 2 * len > G) { L.flags |= FWF_SLOW; break; }
      L.ck++; L.refPos += len; L.flags &= ~FWF_LASTREFM;
      if (!(L.flags & FWF_BLOCKOPEN)) {
        if (L.nblk == FW_MAX_BLOCKS) { L.flags |= FWF_SLOW; break; }
        L.flags |= FWF_BLOCKOPEN; L.flushE = L.blkFirstE = (L.refPos + 1) >> 5; L.maxPos = L.refPos + 1;
      }
      uint2 t; t.x = FW_TASK_D | ((unsigned)(len - 1) << 28); t.y = (unsigned)L.refPos;
      tasks[L.nt * TS] = t;
      L.maxPos = fw_max(L.maxPos, L.refPos + len);
      L.maps += len; L.errs += len; L.nt++;
      if (L.nt == FW_K) break;
      continue;
    }
    if (op == 3) {          // N: ends the aligned block; the round that closed the block consumes the element next time
      if (L.flags & FWF_BLOCKOPEN) { L.flags |= FWF_ENDBLOCK; break; }
      if (len == 0 || !(L.flags & FWF_FIRSTREF) || (long long)L.refPos + len > (long long)G) { L.flags |= FWF_SLOW; break; }
      L.ck++; L.refPos += len; L.flags &= ~FWF_LASTREFM;
      continue;
    }
    if (op == 1 || op == 4) {  // I, S consume read bases only
      if (len == 0 || (unsigned)L.readPos + (unsigned)len >= FW_NIB_MAX) { L.flags |= FWF_SLOW; break; }
      L.readPos += len; L.ck++;
      continue;
    }
    if (op == 5 || op == 6) { L.ck++; continue; }  // H, P
    L.flags |= FWF_SLOW;    // = X and invalid ops: the serial walker restates them literally
    break;
  }
}

// ---- phase 2: tasks -> match masks, breaks, ring ---------------------------------------------------------------------
// seq[k] = raw (little-endian) word k of the base tile; refg[W] = reference grid word W (positions 8W..8W+7, first in the
// top nibble).  ring[j * RS] = lane-private ring word j.
template <bool RECORD>
FW_HD void fw_task(FwLane& L, const DevCfg& c, const BatchDev& b, const uint2 tk, const uint32_t* seq, const uint32_t* refg, uint32_t* ring, int RS) {
  const bool isD = (tk.x & FW_TASK_D) != 0;
  const int n = (int)(tk.x >> 28) + 1;
  const int p0 = (int)tk.y + 1;              // first emitted position
  const unsigned t4 = ((unsigned)p0 & 7u) << 2;
  const unsigned W0 = (unsigned)p0 >> 3;
  // 16 read bases from the task's first base, normalised to the top of (hi:lo)
  const int rel = isD ? 0 : (int)(tk.x & (FW_NIB_MAX - 1)) - (int)(2 * (L.tb - (long long)L.so));
  const int k0 = rel >> 3;
  const unsigned o4 = ((unsigned)rel & 7u) << 2;
  const unsigned w0 = fw_bswap(seq[k0]), w1 = fw_bswap(seq[k0 + 1]), w2 = fw_bswap(seq[k0 + 2]);
  const unsigned hi = fw_fl(w1, w0, o4), lo = fw_fl(w2, w1, o4);
  // 16 reference bases from position p0
  const unsigned R0 = refg[W0], R1 = refg[W0 + 1], R2 = refg[W0 + 2];
  const unsigned rhi = fw_fl(R1, R0, t4), rlo = fw_fl(R2, R1, t4);
  const unsigned xh = hi ^ rhi, xl = lo ^ rlo;
  const unsigned nzh = ((xh & 0x77777777u) + 0x77777777u) | xh, nzl = ((xl & 0x77777777u) + 0x77777777u) | xl;  // bit 3 of a nibble: bases differ
  const unsigned n4 = (unsigned)n << 2;
  const unsigned vh = ~fw_frc(0xFFFFFFFFu, 0u, n4);                               // top n nibbles
  const unsigned vl = ~fw_frc(0xFFFFFFFFu, 0u, (n4 > 32u ? n4 : 32u) - 32u);
  const unsigned mm = isD ? 0u : FW_MIS;
  const unsigned eh = nzh & mm & vh, el = nzl & mm & vl;     // mismatches
  const unsigned mh = ~nzh & mm & vh, ml = ~nzl & mm & vl;   // matches
  // ---- CigarCluster.add :470-498: only when some position of the task can be more than T past the last match
  if (p0 + n - 1 - L.pp0 > c.T) {
    const int prev = L.pp0 + fw_last_idx(L.pmh, L.pml);
    const int f = fw_first_idx(mh, ml, n);
    const int i0 = fw_max(0, prev + c.T + 1 - p0);
    if (RECORD) {
      for (int i = i0; i < fw_min(f, n); i++) {   // emissions that are not matches: mapStart[pos]++, mapEnd[prev]++
        if (L.nextra == FW_MAX_EXTRA) { L.flags |= FWF_SLOW; break; }
        b.fhdr[L.r * FW_HDR_WORDS + 12 + 2 * L.nextra] = (unsigned)(p0 + i);
        b.fhdr[L.r * FW_HDR_WORDS + 13 + 2 * L.nextra] = (unsigned)prev;
        L.nextra++;
      }
    }
    if (f < n && f >= i0) {                       // the first match closes the break (prev, pos)
      if (L.njunc == FW_MAX_JUNC) L.flags |= FWF_SLOW;
      else {
        b.break_arena[L.r * FW_BRK_INLINE + L.nb] = prev;
        b.break_arena[L.r * FW_BRK_INLINE + L.nb + 1] = p0 + f;
        if (RECORD) { b.fhdr[L.r * FW_HDR_WORDS + 8 + 2 * L.njunc] = (unsigned)prev; b.fhdr[L.r * FW_HDR_WORDS + 9 + 2 * L.njunc] = (unsigned)(p0 + f); }
        L.nb += 2; L.njunc++;
      }
    }
  }
  if (mh | ml) { L.pmh = mh; L.pml = ml; L.pp0 = p0; }
  if (RECORD) {
    const unsigned pat = isD ? FW_QRK : FW_COV;
    const unsigned dh = eh | (pat & vh), dl = el | (pat & vl);
    const unsigned g0 = dh >> t4, g1 = fw_fr(dl, dh, t4), g2 = fw_fr(0u, dl, t4);
    const unsigned j0 = W0 & (FW_RING - 1), j1 = (W0 + 1) & (FW_RING - 1), j2 = (W0 + 2) & (FW_RING - 1);
    const unsigned o0 = ring[j0 * RS], o1 = ring[j1 * RS], o2 = ring[j2 * RS];
    ring[j0 * RS] = o0 | g0; ring[j1 * RS] = o1 | g1; ring[j2 * RS] = o2 | g2;
    const unsigned v0 = o0 & g0 & FW_QRK, v1 = o1 & g1 & FW_QRK, v2 = o2 & g2 & FW_QRK;
    if (v0 | v1 | v2) {   // a position emitted by two D elements: the flag cannot count to 2, record the second emission
      for (int j = 0; j < 3; j++) {
        unsigned v = j == 0 ? v0 : (j == 1 ? v1 : v2);
        while (v) {
          const int bit = 31 - fw_clz(v);
          v &= ~(1u << bit);
          if (L.nextra == FW_MAX_EXTRA) { L.flags |= FWF_SLOW; break; }
          b.fhdr[L.r * FW_HDR_WORDS + 12 + 2 * L.nextra] = 0x80000000u | (unsigned)((W0 + j) * 8 + ((31 - bit) >> 2));
          b.fhdr[L.r * FW_HDR_WORDS + 13 + 2 * L.nextra] = 0;
          L.nextra++;
        }
      }
    }
  }
}

// ---- end of round: completed entries leave the ring; blocks and reads are closed ------------------------------------
template <bool RECORD>
FW_HD void fw_flush(FwLane& L, const DevCfg& c, const BatchDev& b, uint32_t* ring, int RS) {
  if (!(L.flags & FWF_ACTIVE)) return;
  const bool fin = (L.flags & (FWF_ENDREAD | FWF_ENDBLOCK | FWF_SLOW)) != 0;
  if (RECORD && (L.flags & FWF_BLOCKOPEN)) {
    // an entry (32 positions) is complete when every position of it lies before the next position an M element can emit
    int endE = fin ? (L.maxPos >> 5) + 1 : (L.refPos + 1) >> 5;
    while (L.flushE < endE) {
      const int j = (L.flushE & (FW_RING / 4 - 1)) * 4;
      uint4 e;
      e.x = ring[(j + 0) * RS]; e.y = ring[(j + 1) * RS]; e.z = ring[(j + 2) * RS]; e.w = ring[(j + 3) * RS];
      ring[(j + 0) * RS] = 0; ring[(j + 1) * RS] = 0; ring[(j + 2) * RS] = 0; ring[(j + 3) * RS] = 0;
      if ((unsigned)L.nent < L.entCap) b.fent[(size_t)L.entBase + (unsigned)L.nent] = e;
      else { L.flags |= FWF_SLOW; endE = (L.maxPos >> 5) + 1; }   // does not fit the read's share of the arena: empty the ring, serial path
      L.errs += fw_popc((e.x & FW_MIS) | ((e.y & FW_MIS) >> 1) | ((e.z & FW_MIS) >> 2) | ((e.w & FW_MIS) >> 3));
      L.nent++; L.flushE++;
    }
    if (fin && L.nblk < FW_MAX_BLOCKS) {
      b.fhdr[L.r * FW_HDR_WORDS + 2 + 2 * L.nblk] = (unsigned)L.blkFirstE;
      b.fhdr[L.r * FW_HDR_WORDS + 3 + 2 * L.nblk] = (unsigned)(L.flushE - L.blkFirstE);
      L.nblk++;
    }
  }
  if (fin || (L.flags & FWF_SLOW)) L.flags &= ~(FWF_BLOCKOPEN | FWF_ENDBLOCK);
  if (L.flags & (FWF_ENDREAD | FWF_SLOW)) {
    const long long r = L.r;
    if (!(L.flags & FWF_SLOW) && L.nb + 1 > FW_BRK_INLINE) L.flags |= FWF_SLOW;
    if (L.flags & FWF_SLOW) {
      b.rflags[r] = 0x20u;    // NPT_RF_SLOWPATH: walked again by the exact serial device path
#if defined(__CUDA_ARCH__)
      const unsigned k = atomicAdd(b.bctr + 7, 1u);
#else
      const unsigned k = b.bctr[7]++;
#endif
      b.slow_list[k] = (unsigned)r;
    } else {
      const int alnEnd = L.refPos;  // CigarCluster.addEnd :466-469
      b.break_arena[r * FW_BRK_INLINE + L.nb] = alnEnd;
      b.nbreaks[r] = L.nb + 1;
      b.break_loc[r] = 0xFFFFFFFFu;
      b.read_st[r] = L.st_r;
      b.read_en[r] = (L.flags & FWF_LASTREFM) ? L.lmRead0 + L.lmLen : 0;   // getReadPositionAtReferencePosition(alignmentEnd)
      b.maps_sum[r] = c.record_depth ? L.maps : 0;
      b.err_sum[r] = c.record_depth ? L.errs : 0;
      b.rflags[r] = 0;
      if (RECORD) {
        b.fhdr[r * FW_HDR_WORDS + 0] = L.entBase;
        b.fhdr[r * FW_HDR_WORDS + 1] = (unsigned)L.nblk | ((unsigned)L.njunc << 8) | ((unsigned)L.nextra << 16);
      }
    }
    L.flags &= FWF_DONE;
  }
}

}  // namespace npt
