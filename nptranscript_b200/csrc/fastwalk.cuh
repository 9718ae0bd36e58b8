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
//   [1+k]            block k < 3: first entry index on the reference grid (position >> 5) | number of entries << 16; 0 = no block
//   [4] nblk | njunc << 8 | nextra << 16 | ndbl << 24
//   [5+d]            d < 3: a position emitted by two D elements (depth[pos]++, errors[pos]++ on top of the entry's single flag);
//                    a read's 4th and 5th such position go to the extras as kind 1
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
  int tnib0;                    // read base index of the tile's first base (2 * (tb - so), may be negative)
  int st_r, lmRead0, lmLen;     // read offset of the first aligned base; last M element (for the read offset of the alignment end)
  unsigned flags;
  int nb, nblk, nent, njunc, nextra, ndbl, errs, maps, nt;
  unsigned pmh, pml; int pp0;   // last match: the last set flag of (pmh:pml), a task's match mask, counted from position pp0
  int maxPos, flushE, blkFirstE;
  unsigned blk0, blk1, blk2;    // closed blocks: first entry | entries << 16
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
  L.tb = -1; L.tnib0 = 0;
  L.entBase = (unsigned)((L.so >> 4) + 8ull * (unsigned long long)r);
  L.entCap = (unsigned)(((se >> 4) + 8ull * (unsigned long long)(r + 1)) - ((L.so >> 4) + 8ull * (unsigned long long)r));
  L.refPos = L.alnStart - 1; L.readPos = 0; L.mRem = 0;
  L.st_r = 0; L.lmRead0 = 0; L.lmLen = 0;
  L.nb = 1; L.nblk = 0; L.nent = 0; L.njunc = 0; L.nextra = 0; L.ndbl = 0; L.errs = 0; L.maps = 0; L.nt = 0;
  L.pmh = 0x80000000u; L.pml = 0; L.pp0 = L.alnStart;   // CigarCluster.addStart :442-446: prev = alignment start
  L.maxPos = 0; L.flushE = 0; L.blkFirstE = 0; L.blk0 = L.blk1 = L.blk2 = 0;
  if (L.alnStart <= 0) L.flags |= FWF_SLOW;             // not a mapped record: the serial walker flags it
  else b.break_arena[r * FW_BRK_INLINE] = L.alnStart;
}

// ---- phase 1: CIGAR -> tasks ---------------------------------------------------------------------------------------
// cig[i] = CIGAR word (L.cs + i), i < cigWords.  tasks[i * TS] receives task i: x = read base index | D flag | (n-1) << 28,
// y = refPos before the first emitted position (the task emits positions y+1 .. y+n).
// One iteration = [take one CIGAR element if the current M element is used up] + [emit at most one task], written as
// straight-line arithmetic: lanes sit on different element types all the time, and a branch per type serialises the warp
// (ncu r02b: 3-12 of 32 lanes active in the branchy version, half of the kernel's instructions).  Everything the fast walk
// does not restate only sets a sticky word that is looked at once per round; a round that ran on such an element produced
// tasks nobody uses (fw_task clamps its reference index, the read goes to the serial walker).
template <bool RECORD>
FW_HD void fw_produce(FwLane& L, const DevCfg& c, const uint32_t* cig, int cigWords, uint2* tasks, int TS) {
  L.nt = 0;
  if ((L.flags & (FWF_ACTIVE | FWF_SLOW)) != FWF_ACTIVE) return;
  const int tnibLast = L.tnib0 + FW_SEQ_TILE * 8 - (FW_TASK_MAX + 8);   // last read base a task may start at with this tile
  bool open = (L.flags & FWF_BLOCKOPEN) != 0, firstref = (L.flags & FWF_FIRSTREF) != 0, lastM = (L.flags & FWF_LASTREFM) != 0;
  unsigned viol = 0, orRef = 0, orRead = 0;
  unsigned endf = 0;
  int nt = 0, refPos = L.refPos, readPos = L.readPos, mRem = L.mRem;
  unsigned ck = L.ck;
  bool go = true;
  while (go) {
    // ---- (a) one CIGAR element; lanes that do not take one see a no-op (P of length 0)
    const bool need = mRem == 0;
    const bool atEnd = need && ck == L.cend;
    const bool fetch = need && !atEnd && (ck - L.cs < (unsigned)cigWords);
    unsigned w = cig[fetch ? (ck - L.cs) : 0u];
    w = fetch ? w : 6u;
    const unsigned op = w & 15u, len = w >> 4;
    const unsigned pr = (0x00084281u >> ((op & 7u) << 2)) & 15u;   // 1 = M, 2 = D, 4 = N, 8 = I / S, 0 = H / P
    const bool isM = (pr & 1u) != 0, isD = (pr & 2u) != 0, isN = (pr & 4u) != 0;
    const bool nStop = isN && open;                 // N ends the aligned block: the round that closed it takes the element next time
    const unsigned lenT = nStop ? 0u : len;
    viol |= (op + 9u) >> 4;                         // = X and invalid ops
    viol |= (pr != 0u && len == 0u) ? 1u : 0u;      // zero-length elements
    viol |= (isD && len > (unsigned)FW_TASK_MAX) ? 1u : 0u;
    viol |= ((pr & 6u) && !firstref) ? 1u : 0u;     // D / N before the first M
    viol |= ((pr & 3u) && !open && L.nblk == FW_MAX_BLOCKS) ? 1u : 0u;
    ck += (fetch && !nStop) ? 1u : 0u;
    refPos += (pr & 6u) ? (int)lenT : 0;            // D, N advance first; D then emits len mismatching positions (:513-522)
    if (isM) {                                      // M: emits from its own start (:531-551)
      L.st_r = firstref ? L.st_r : readPos + 1;
      L.lmRead0 = readPos; L.lmLen = (int)len; mRem = (int)len;
    }
    firstref = firstref || isM;
    lastM = isM || (lastM && !(pr & 6u));
    L.maps += (pr & 3u) ? (int)lenT : 0;
    L.errs += isD ? (int)lenT : 0;
    readPos += (pr & 8u) ? (int)lenT : 0;
    if (pr & 3u) {
      if (!open) { open = true; L.flushE = L.blkFirstE = (refPos + 1) >> 5; L.maxPos = refPos + 1; }
      L.maxPos = fw_max(L.maxPos, refPos + (int)len);
    }
    orRef |= (unsigned)refPos | len; orRead |= (unsigned)readPos;
    // ---- (b) at most one task: the positions of the D element just taken, or the next piece of the current M element
    const bool haveM = mRem > 0;
    const bool staged = readPos <= tnibLast;
    const bool emitM = haveM && staged;
    const int n = isD ? (int)len : fw_min(mRem, FW_TASK_MAX);
    if (isD || emitM) {
      uint2 t;
      t.x = (isD ? FW_TASK_D : (unsigned)readPos) | ((unsigned)(n - 1) << 28);
      t.y = (unsigned)refPos;
      tasks[nt * TS] = t;
      nt++;
    }
    const int adv = emitM ? n : 0;
    refPos += adv; readPos += adv; mRem -= adv;
    // ---- (c) the round ends at the end of the read / block / staged data, or when the task buffer is full
    endf |= (atEnd ? FWF_ENDREAD : 0u) | (nStop ? FWF_ENDBLOCK : 0u);
    go = !(atEnd || nStop || (need && !atEnd && !fetch) || (haveM && !staged) || nt == FW_K || viol != 0u);
  }
  L.nt = nt; L.refPos = refPos; L.readPos = readPos; L.mRem = mRem; L.ck = ck;
  // ranges: positions <= G (a round adds < 2^28 per element to positions < 2^29, so the sticky OR sees every excursion before
  // an int can wrap), read bases < 2^27 (the task encoding)
  const bool slow = viol != 0u || (orRef >> 29) != 0u || (orRead >> 27) != 0u || L.maxPos > c.G || refPos > c.G;
  L.flags = (L.flags & ~(FWF_BLOCKOPEN | FWF_FIRSTREF | FWF_LASTREFM)) | (open ? FWF_BLOCKOPEN : 0u) | (firstref ? FWF_FIRSTREF : 0u) |
            (lastM ? FWF_LASTREFM : 0u) | endf | (slow ? FWF_SLOW : 0u);
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
  const unsigned Wr = W0 < (unsigned)(c.refg_nwords - 3) ? W0 : (unsigned)(c.refg_nwords - 3);   // memory safety only
  // 16 read bases from the task's first base, normalised to the top of (hi:lo)
  const int rel = isD ? 0 : (int)(tk.x & (FW_NIB_MAX - 1)) - L.tnib0;
  const int k0 = rel >> 3;
  const unsigned o4 = ((unsigned)rel & 7u) << 2;
  const unsigned w0 = fw_bswap(seq[k0]), w1 = fw_bswap(seq[k0 + 1]), w2 = fw_bswap(seq[k0 + 2]);
  const unsigned hi = fw_fl(w1, w0, o4), lo = fw_fl(w2, w1, o4);
  // 16 reference bases from position p0
  const unsigned R0 = refg[Wr], R1 = refg[Wr + 1], R2 = refg[Wr + 2];
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
          const unsigned q = (unsigned)((W0 + j) * 8 + ((31 - bit) >> 2));
          if (L.ndbl < 3) { b.fhdr[L.r * FW_HDR_WORDS + 5 + L.ndbl] = q; L.ndbl++; continue; }
          if (L.nextra == FW_MAX_EXTRA) { L.flags |= FWF_SLOW; break; }
          b.fhdr[L.r * FW_HDR_WORDS + 12 + 2 * L.nextra] = 0x80000000u | q;
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
  if (RECORD && (L.flags & FWF_SLOW)) {
    for (int j = 0; j < FW_RING; j++) ring[j * RS] = 0;   // whatever the read left behind; its entries are never looked at
  } else if (RECORD && (L.flags & FWF_BLOCKOPEN)) {
    // an entry (32 positions) is complete when every position of it lies before the next position an M element can emit
    const int endE = fin ? (L.maxPos >> 5) + 1 : (L.refPos + 1) >> 5;
    while (L.flushE < endE) {
      const int j = (L.flushE & (FW_RING / 4 - 1)) * 4;
      uint4 e;
      e.x = ring[(j + 0) * RS]; e.y = ring[(j + 1) * RS]; e.z = ring[(j + 2) * RS]; e.w = ring[(j + 3) * RS];
      ring[(j + 0) * RS] = 0; ring[(j + 1) * RS] = 0; ring[(j + 2) * RS] = 0; ring[(j + 3) * RS] = 0;
      if ((unsigned)L.nent < L.entCap) b.fent[(size_t)L.entBase + (unsigned)L.nent] = e;
      else L.flags |= FWF_SLOW;   // does not fit the read's share of the arena (the ring is emptied all the same): serial path
      L.errs += fw_popc((e.x & FW_MIS) | ((e.y & FW_MIS) >> 1) | ((e.z & FW_MIS) >> 2) | ((e.w & FW_MIS) >> 3));
      L.nent++; L.flushE++;
    }
    if (fin && L.nblk < FW_MAX_BLOCKS) {
      const unsigned cnt = (unsigned)(L.flushE - L.blkFirstE);
      const unsigned v = (unsigned)L.blkFirstE | (cnt << 16);
      if (cnt >= 0x10000u || (unsigned)L.blkFirstE >= 0x10000u) L.flags |= FWF_SLOW;
      if (L.nblk == 0) L.blk0 = v; else if (L.nblk == 1) L.blk1 = v; else L.blk2 = v;
      L.nblk++;
    }
  }
  if (fin || (L.flags & FWF_SLOW)) L.flags &= ~(FWF_BLOCKOPEN | FWF_ENDBLOCK);
  if (L.flags & (FWF_ENDREAD | FWF_SLOW)) {
    const long long r = L.r;
    if (!(L.flags & FWF_SLOW) && (L.nb + 1 > FW_BRK_INLINE || L.lmRead0 + L.lmLen > L.nbases)) L.flags |= FWF_SLOW;
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
        uint4 h; h.x = L.entBase; h.y = L.blk0; h.z = L.blk1; h.w = L.blk2;
        *reinterpret_cast<uint4*>(b.fhdr + r * FW_HDR_WORDS) = h;
        b.fhdr[r * FW_HDR_WORDS + 4] = (unsigned)L.nblk | ((unsigned)L.njunc << 8) | ((unsigned)L.nextra << 16) | ((unsigned)L.ndbl << 24);
      }
    }
    L.flags &= FWF_DONE;
  }
}

}  // namespace npt
