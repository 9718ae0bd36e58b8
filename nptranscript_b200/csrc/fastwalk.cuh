// fastwalk.cuh — the single-pass fast walk (one THREAD per read, one fused straight-line step per M piece).
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
// Design (why it looks like this).  ncu history: the round-1 kernel ran a four-phase loop per CIGAR element with 19 of 32
// lanes active (1 700-2 400 warp-instructions per read and walk); the first version of this file produced "tasks" from the
// CIGAR in one loop and consumed them in a second one (profiles/r02d: 2 900 warp-instructions per read, 116 instructions
// per CIGAR element in the branch-free producer alone).  Now ONE step does both, shaped after what a nanopore CIGAR looks
// like — M pieces separated by single I / D elements:
//     [take one I / S / H / P / D element]  +  [start the M element behind it]  +  [compare <= 16 bases of the current M]
// A D element never makes a step of its own: the positions it emits (after it advanced, the reference's quirk) start at
// the same position as the M piece behind it, so its flags go into the same three ring words.  Everything rare (N, end of
// read, staged data used up) parks the lane until the end of the round, where a short handler runs once.
// Coverage is not credited with atomics: the per-position flags (covered, deletion-emitted, mismatch) are ORed into a
// lane-private ring of 4-bit-per-position words in shared memory and leave the kernel as 16-byte ENTRIES (32 positions
// each, on the reference grid).  Once the cluster of the read is known, accumulate_kernel adds the entries of all reads of
// a (cluster, source) group with bit-sliced carry-save adders (fastcov.cuh) — positional popcount instead of 250 scattered
// REDs per read.
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

constexpr int FW_PIECE = 16;       // positions per step
#ifndef NPT_FW_R
#define NPT_FW_R 16
#endif
constexpr int FW_R = NPT_FW_R;     // steps per round (between two stagings of a lane's CIGAR / base tiles)
constexpr int FW_RING = 32;        // ring of 8-position words per lane = 256 positions ...
constexpr int FW_RING_PAD = 2;     // ... + two words behind it that alias words 0 and 1 (a step writes three consecutive words)
// staged per lane and round: CIGAR words (a step looks at two and takes 1.2 on average; the tile starts up to 3 words before
// the cursor) and base words (8 bases each; a step takes at most 16 bases, 10 on average, plus what I / S elements skip; the
// tile starts up to 31 bases before the cursor and a step reads 24 bases from its first one)
constexpr int FW_CIG_TILE = FW_R <= 8 ? 20 : (FW_R <= 12 ? 24 : 28);
constexpr int FW_SEQ_TILE = FW_R <= 8 ? 24 : (FW_R <= 12 ? 32 : 40);
constexpr int FW_TILE_STRIDE = FW_CIG_TILE + FW_SEQ_TILE;   // words between lane tiles (16-byte aligned)
constexpr int FW_MAX_BLOCKS = 3;   // aligned blocks (runs between N elements) per read
constexpr int FW_MAX_JUNC = 2;     // breaks (prev, pos) per read: BRK_INLINE = 6 = start + 2 pairs + end
constexpr int FW_MAX_EXTRA = 4;    // rare extra events per read (see below)
constexpr int FW_RARE_WORDS = 16;  // per-read record of rare events, uint32 words
constexpr int FW_BRK_INLINE = 6;
// Without depth recording (host mode, opts_host.txt) nothing is kept per aligned block, so the number of blocks is free and
// the break list may be long: a spliced read has two breaks per junction.  The lane keeps them in FW_BRK_BUF words of a
// lane-private scratch in global memory (written at junctions only: a few stores per read) and moves the list to the batch's
// overflow area at the end of the read, like walk_kernel<2> does for the serial walker.
constexpr int FW_LONG_JUNC = 127;
constexpr int FW_BRK_BUF = 2 * FW_LONG_JUNC + 2;

// entry flags, one nibble per position (position 8W + i of ring word W sits in nibble i counted from the top)
constexpr unsigned FW_COV = 0x11111111u;   // emitted by an M element (depth +1)
constexpr unsigned FW_QRK = 0x22222222u;   // emitted by a D element after it advanced (depth +1, error +1)
constexpr unsigned FW_MIS = 0x88888888u;   // M position whose base differs from the reference (error +1)

// What the fast walk leaves per read (read by cov_key_kernel and accumulate_kernel):
//   fhdrA[r]   = { index of the read's first entry in the batch's entry arena,
//                  block 0, 1, 2: first entry index on the reference grid (position >> 5) | number of entries << 16; 0 = no block }
//   fmeta[r]   = nblk | njunc << 8 | nextra << 16 | ndbl << 24
//   frare[r][] = [d], d < 3       a position emitted by two D elements (depth[pos]++, errors[pos]++ on top of the entry's single
//                                 flag); further ones go to the extras as kind 1
//                [4+2j], [5+2j]   junction j: (prev, pos) = mapEnd[prev]++, mapStart[pos]++
//                [8+2e], [9+2e]   extra e < 4: kind 0 (bit 31 clear) = emission `pos` more than break_thresh after `prev` that was
//                                 not a match (mapStart[pos]++, mapEnd[prev]++); kind 1 (bit 31 set) = a position emitted by two
//                                 D elements
enum : unsigned { FWF_ACTIVE = 1, FWF_DONE = 2, FWF_SLOW = 4 };

struct FwLane {
  long long r;                  // read index in the batch
  unsigned ck, cend, cs;        // CIGAR cursor / end (word indices in b.cigar); first word held by the CIGAR tile
  int alnStart, refPos, mRem, nbases;
  unsigned readPos;
  unsigned long long so;        // byte offset of the read's bases in b.seq4
  long long tb;                 // byte offset in b.seq4 of the base tile (16-byte aligned), -1 = nothing staged
  int tnib0;                    // read base index of the tile's first base (2 * (tb - so), may be negative)
  int st_r;                     // read offset of the first aligned base, 0 = no M element yet
  int rdEndM, refEndM;          // read offset / reference position behind the last M element
  unsigned flags;
  int nb, nblk, nent, njunc, nextra, ndbl, errs, sumN;
  unsigned pmh, pml; int pp0;   // last match: the last set flag of (pmh:pml), a step's match mask, counted from position pp0
  int qEnd;                     // first position behind everything D elements have emitted so far
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

// index (counted from the first position of the step) of the last / first set match flag of a step's mask
FW_HD int fw_last_idx(unsigned mh, unsigned ml) { return ml ? 8 + ((31 - fw_ctz(ml)) >> 2) : ((31 - fw_ctz(mh)) >> 2); }
FW_HD int fw_first_idx(unsigned mh, unsigned ml, int none) { return mh ? (fw_clz(mh) >> 2) : (ml ? 8 + (fw_clz(ml) >> 2) : none); }
// the top n nibbles of a 16-nibble (hi:lo) pair, n <= 16
FW_HD unsigned fw_top_hi(unsigned n4) { return ~fw_frc(0xFFFFFFFFu, 0u, n4); }
FW_HD unsigned fw_top_lo(unsigned n4) { return ~fw_frc(0xFFFFFFFFu, 0u, (n4 > 32u ? n4 : 32u) - 32u); }

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
  L.st_r = 0; L.rdEndM = 0; L.refEndM = -1;
  L.nb = 1; L.nblk = 0; L.nent = 0; L.njunc = 0; L.nextra = 0; L.ndbl = 0; L.errs = 0; L.sumN = 0;
  L.pmh = 0x80000000u; L.pml = 0; L.pp0 = L.alnStart;   // CigarCluster.addStart :442-446: prev = alignment start
  L.qEnd = 0;
  // the first aligned block is open from the start: its first entry is the one that holds the alignment start
  L.maxPos = L.alnStart - 1; L.flushE = L.blkFirstE = L.alnStart >> 5; L.blk0 = L.blk1 = L.blk2 = 0;
  if (L.alnStart <= 0 || L.alnStart >= (1 << 29)) L.flags |= FWF_SLOW;   // not a mapped record: the serial walker flags it
  else b.break_arena[r * FW_BRK_INLINE] = L.alnStart;
}

// a lane that cannot go on (more events than its header holds) stops taking elements; the end of the round hands it over
FW_HD void fw_give_up(FwLane& L) { L.flags |= FWF_SLOW; L.ck = L.cend; L.mRem = 0; }

// ---- the rare part of a step: some emission lies more than break_thresh behind the last match (CigarCluster.add :470-498) ----
// emissions in the reference's order: the dq positions of the D element (never matches), then the n positions of the M piece
template <bool RECORD>
FW_HD void fw_far_step(FwLane& L, const DevCfg& c, const BatchDev& b, int p0, int n, int dq, unsigned mh, unsigned ml, uint32_t* aux, int RS) {
  const int prev = L.pp0 + fw_last_idx(L.pmh, L.pml);
  const int f = fw_first_idx(mh, ml, n);            // first match of the M piece (n = none)
  const int i0 = fw_max(0, prev + c.T + 1 - p0);    // first offset that is more than T behind prev
  if (RECORD) {
    for (int pass = 0; pass < 2; pass++) {          // emissions that are not matches: mapStart[pos]++, mapEnd[prev]++
      const int hi = pass == 0 ? dq : fw_min(f, n);
      for (int i = i0; i < hi; i++) {
        if (L.nextra == FW_MAX_EXTRA) { fw_give_up(L); return; }
        b.frare[L.r * FW_RARE_WORDS + 8 + 2 * L.nextra] = (unsigned)(p0 + i);
        b.frare[L.r * FW_RARE_WORDS + 9 + 2 * L.nextra] = (unsigned)prev;
        L.nextra++;
      }
    }
  }
  if (f < n && f >= i0) {                           // the first match closes the break (prev, pos)
    if (L.njunc == (RECORD ? FW_MAX_JUNC : FW_LONG_JUNC)) { fw_give_up(L); return; }
    if (L.nb + 1 < FW_BRK_INLINE) {
      b.break_arena[L.r * FW_BRK_INLINE + L.nb] = prev;
      b.break_arena[L.r * FW_BRK_INLINE + L.nb + 1] = p0 + f;
    }
    if (!RECORD) { aux[L.nb * RS] = (uint32_t)prev; aux[(L.nb + 1) * RS] = (uint32_t)(p0 + f); }
    if (RECORD) { b.frare[L.r * FW_RARE_WORDS + 4 + 2 * L.njunc] = (unsigned)prev; b.frare[L.r * FW_RARE_WORDS + 5 + 2 * L.njunc] = (unsigned)(p0 + f); }
    L.nb += 2; L.njunc++;
  }
}

// positions [p0, qe) were emitted by an earlier D element and are emitted again by this one: the entry's flag cannot count
// to 2, every further emission is recorded by position
FW_HD void fw_double_qrk(FwLane& L, const BatchDev& b, int p0, int qe) {
  for (int q = p0; q < qe; q++) {
    if (L.ndbl < 3) { b.frare[L.r * FW_RARE_WORDS + L.ndbl] = (unsigned)q; L.ndbl++; continue; }
    if (L.nextra == FW_MAX_EXTRA) { fw_give_up(L); return; }
    b.frare[L.r * FW_RARE_WORDS + 8 + 2 * L.nextra] = 0x80000000u | (unsigned)q;
    b.frare[L.r * FW_RARE_WORDS + 9 + 2 * L.nextra] = 0;
    L.nextra++;
  }
}

// ---- one step ---------------------------------------------------------------------------------------------------------
// cig[i] = CIGAR word (L.cs + i), i < FW_CIG_TILE; seq[k] = raw (little-endian) word k of the base tile; refg[W] = reference
// grid word W (positions 8W..8W+7, first in the top nibble); ring[j * RS] = lane-private ring word j, j < FW_RING + FW_RING_PAD
// (RECORD) or lane-private break scratch word j, j < FW_BRK_BUF (!RECORD; global memory, RS = 1).
// `live` = the lane holds a read the fast walk has not given up; ringLimit = last refPos a step may start from without
// running past the 256 positions the ring holds behind the first unflushed entry.
// The step takes an element only when it can restate it: I S H P always, D of at most 16 positions after the first M, M.  At
// anything else (N, = X, invalid ops, zero-length elements, end of the CIGAR, end of the staged words / bases) the lane simply
// does nothing for the rest of the round — no flag is kept; fw_round_end looks at the element the cursor stands on.
template <bool RECORD>
FW_HD void fw_step(FwLane& L, const DevCfg& c, const BatchDev& b, const uint32_t* cig, const uint32_t* seq, const uint32_t* refg, uint32_t* ring,
                   int RS, bool live, int ringLimit) {
  const bool can = RECORD ? (live & (L.refPos <= ringLimit)) : live;
  // ---- (a) CIGAR: one element that is not M (applied at once) and / or the start of the M element behind it
  const unsigned ci = L.ck - L.cs;
  const unsigned cic = ci < (unsigned)(FW_CIG_TILE - 2) ? ci : (unsigned)(FW_CIG_TILE - 2);   // memory safety
  const unsigned w0 = cig[cic], w1 = cig[cic + 1];
  const unsigned ob = 1u << (w0 & 15u);
  const unsigned len0 = w0 >> 4;
  const bool isD = (ob & 4u) != 0u;
  bool take = can & (L.mRem == 0) & (ci < (unsigned)(FW_CIG_TILE - 1)) & (L.ck < L.cend) & (w0 >= 16u) & ((ob & 0xFF88u) == 0u);
  take = take & !(isD & ((len0 > (unsigned)FW_PIECE) | (L.st_r == 0)));
  const bool nonM = take & ((ob & 1u) == 0u);
  L.readPos += (nonM & ((ob & 0x12u) != 0u)) ? len0 : 0u;          // I, S: read bases only (H, P: nothing)
  const int dq = (nonM & isD) ? (int)len0 : 0;
  L.refPos += dq;                                                  // D advances first and then emits dq mismatching positions (:513-522)
  L.errs += dq;
  const unsigned wM = nonM ? w1 : w0;
  const bool startM = take & ((wM & 15u) == 0u) & (wM >= 16u) & (!nonM | (L.ck + 1u < L.cend));
  const int lenM = (int)(wM >> 4);
  L.ck += (nonM ? 1u : 0u) + (startM ? 1u : 0u);
  if (startM) {                                                    // M emits from its own start (:531-551)
    L.mRem = lenM; L.rdEndM = (int)L.readPos + lenM; L.refEndM = L.refPos + lenM;
    if (L.st_r == 0) L.st_r = (int)L.readPos + 1;
  }
  // ---- (b) at most 16 positions of the current M element
  const unsigned tn = L.readPos - (unsigned)L.tnib0;               // base index inside the tile
  const bool staged = tn <= (unsigned)(FW_SEQ_TILE * 8 - FW_PIECE - 8);
  const int n = (can & (L.mRem > 0) & staged) ? fw_min(L.mRem, FW_PIECE) : 0;
  const int p0 = L.refPos + 1;                                     // first emitted position (of the D element and of the M piece)
  const unsigned t4 = ((unsigned)p0 & 7u) << 2;
  const unsigned W0 = (unsigned)p0 >> 3;
  const unsigned Wr = W0 < (unsigned)(c.refg_nwords - 3) ? W0 : (unsigned)(c.refg_nwords - 3);   // memory safety only
  const unsigned k0 = (tn >> 3) < (unsigned)(FW_SEQ_TILE - 3) ? (tn >> 3) : (unsigned)(FW_SEQ_TILE - 3);
  const unsigned o4 = (tn & 7u) << 2;
  const unsigned s0 = fw_bswap(seq[k0]), s1 = fw_bswap(seq[k0 + 1]), s2 = fw_bswap(seq[k0 + 2]);
  const unsigned hi = fw_fl(s1, s0, o4), lo = fw_fl(s2, s1, o4);                // 16 read bases from readPos
  const unsigned R0 = refg[Wr], R1 = refg[Wr + 1], R2 = refg[Wr + 2];
  const unsigned rhi = fw_fl(R1, R0, t4), rlo = fw_fl(R2, R1, t4);              // 16 reference bases from p0
  const unsigned xh = hi ^ rhi, xl = lo ^ rlo;
  const unsigned nzh = ((xh & 0x77777777u) + 0x77777777u) | xh, nzl = ((xl & 0x77777777u) + 0x77777777u) | xl;  // bit 3 of a nibble: bases differ
  const unsigned n4 = (unsigned)n << 2;
  const unsigned vh = fw_top_hi(n4) & FW_MIS, vl = fw_top_lo(n4) & FW_MIS;
  const unsigned eh = nzh & vh, el = nzl & vl;       // mismatches
  const unsigned mh = vh & ~nzh, ml = vl & ~nzl;     // matches
  const int m = n > dq ? n : dq;
  const int last = p0 + m - 1;
  if ((m > 0) & (last - L.pp0 > c.T)) fw_far_step<RECORD>(L, c, b, p0, n, dq, mh, ml, ring, RS);   // some emission is more than T behind the last match
  if (mh | ml) { L.pmh = mh; L.pml = ml; L.pp0 = p0; }
  if (RECORD) {
    if ((dq > 0) & (p0 < L.qEnd)) fw_double_qrk(L, b, p0, fw_min(L.qEnd, p0 + dq));
    L.qEnd = dq > 0 ? fw_max(L.qEnd, p0 + dq) : L.qEnd;
    const unsigned q4 = (unsigned)dq << 2;
    const unsigned qh = fw_top_hi(q4), ql = fw_top_lo(q4);
    const unsigned dh = eh | (vh >> 3) | (qh & FW_QRK), dl = el | (vl >> 3) | (ql & FW_QRK);
    const unsigned g0 = dh >> t4, g1 = fw_fr(dl, dh, t4), g2 = fw_fr(0u, dl, t4);
    uint32_t* rw = ring + (W0 & (FW_RING - 1)) * RS;               // three consecutive words (the last two may be the pad)
    rw[0] |= g0; rw[RS] |= g1; rw[2 * RS] |= g2;
  }
  L.maxPos = fw_max(L.maxPos, last);
  L.refPos += n; L.readPos += (unsigned)n; L.mRem -= n;
}

// ---- end of round: N elements, completed entries, closed blocks, finished reads -------------------------------------
template <bool RECORD>
FW_HD void fw_round_end(FwLane& L, const DevCfg& c, const BatchDev& b, const uint32_t* cig, uint32_t* ring, int RS) {
  if (!(L.flags & FWF_ACTIVE)) return;
  bool slow = (L.flags & FWF_SLOW) != 0u;
  // the alignment leaves the reference (the serial walker clamps), or read offsets leave the range a step handles
  slow = slow || L.maxPos > c.G || L.refPos > c.G || L.readPos >= (1u << 28);
  bool endRead = false, isN = false;
  unsigned wN = 0;
  if (!slow && L.mRem == 0) {                      // what does the cursor stand on?
    const unsigned ci = L.ck - L.cs;
    if (L.ck >= L.cend) endRead = true;
    else if (ci < (unsigned)(FW_CIG_TILE - 1)) {
      const unsigned w = cig[ci], op = w & 15u, len = w >> 4;
      if (op == 3u && len != 0u && L.st_r != 0) { isN = true; wN = w; }
      else if (op > 6u || len == 0u || op == 3u || (op == 2u && (len > (unsigned)FW_PIECE || L.st_r == 0))) slow = true;   // not restated here
    }
  }
  const bool fin = endRead || isN || slow;
  if (RECORD) {
    if (slow) {
      for (int j = 0; j < FW_RING + FW_RING_PAD; j++) ring[j * RS] = 0;   // whatever the read left behind; its entries are never looked at
    } else {
      {  // fold the pad words back onto the words they alias
        const unsigned a = ring[FW_RING * RS], bb = ring[(FW_RING + 1) * RS];
        if (a | bb) { ring[0] |= a; ring[RS] |= bb; ring[FW_RING * RS] = 0; ring[(FW_RING + 1) * RS] = 0; }
      }
      // an entry (32 positions) is complete when every position of it lies before the next position an element can emit
      const int endE = fin ? (L.maxPos >> 5) + 1 : (L.refPos + 1) >> 5;
      while (L.flushE < endE) {
        const int j = (L.flushE & (FW_RING / 4 - 1)) * 4;
        uint4 e;
        e.x = ring[(j + 0) * RS]; e.y = ring[(j + 1) * RS]; e.z = ring[(j + 2) * RS]; e.w = ring[(j + 3) * RS];
        ring[(j + 0) * RS] = 0; ring[(j + 1) * RS] = 0; ring[(j + 2) * RS] = 0; ring[(j + 3) * RS] = 0;
        if ((unsigned)L.nent < L.entCap) b.fent[(size_t)L.entBase + (unsigned)L.nent] = e;
        else slow = true;   // does not fit the read's share of the arena (the ring is emptied all the same): serial path
        L.errs += fw_popc((e.x & FW_MIS) | ((e.y & FW_MIS) >> 1) | ((e.z & FW_MIS) >> 2) | ((e.w & FW_MIS) >> 3));
        L.nent++; L.flushE++;
      }
    }
  }
  if (fin && !slow) {                              // the block ends here
    const unsigned cnt = RECORD ? (unsigned)(L.flushE - L.blkFirstE) : 0u;
    const unsigned v = (unsigned)L.blkFirstE | (cnt << 16);
    if (RECORD && (cnt >= 0x10000u || (unsigned)L.blkFirstE >= 0x10000u)) slow = true;   // the packed block header holds 16 bits each
    if (L.nblk == 0) L.blk0 = v; else if (L.nblk == 1) L.blk1 = v; else L.blk2 = v;
    if (RECORD || L.nblk < FW_MAX_BLOCKS) L.nblk++;
  }
  if (isN && !slow) {                              // the N element itself: skips reference positions (:523-525)
    if (RECORD && L.nblk == FW_MAX_BLOCKS) slow = true;      // a fourth block would open behind it
    const int len = (int)(wN >> 4);
    L.refPos += len; L.sumN += len;
    L.ck++;
    L.maxPos = L.refPos; L.flushE = L.blkFirstE = (L.refPos + 1) >> 5;   // the next block is open from here
    if (L.refPos > c.G) slow = true;
  }
  if (endRead && !slow && ((RECORD && L.nb + 1 > FW_BRK_INLINE) || L.rdEndM > L.nbases)) slow = true;
  if (slow) {
    const long long r = L.r;
    b.rflags[r] = 0x20u;    // NPT_RF_SLOWPATH: walked again by the exact serial device path
#if defined(__CUDA_ARCH__)
    const unsigned k = atomicAdd(b.bctr + 7, 1u);
#else
    const unsigned k = b.bctr[7]++;
#endif
    b.slow_list[k] = (unsigned)r;
    L.flags &= FWF_DONE;
  } else if (endRead) {
    const long long r = L.r;
    const int alnEnd = L.refPos;  // CigarCluster.addEnd :466-469
    const int n_all = L.nb + 1;
    b.nbreaks[r] = n_all;
    if (n_all <= FW_BRK_INLINE) {
      b.break_arena[r * FW_BRK_INLINE + L.nb] = alnEnd;
      b.break_loc[r] = 0xFFFFFFFFu;
    } else {   // !RECORD only: the whole list goes to the overflow area
#if defined(__CUDA_ARCH__)
      const unsigned loc = atomicAdd(b.bctr, (unsigned)n_all);
#else
      const unsigned loc = b.bctr[0]; b.bctr[0] += (unsigned)n_all;
#endif
      if (loc + (unsigned)n_all > b.break_cap) {
#if defined(__CUDA_ARCH__)
        atomicOr(b.bctr + 1, 1u);   // too small: the host enlarges it (bctr[0] = the need) and the serial walker redoes the long lists
#else
        b.bctr[1] |= 1u;
#endif
        b.break_loc[r] = 0xFFFFFFFEu;
      } else {
        int* dst = b.break_arena + b.n * FW_BRK_INLINE + loc;
        dst[0] = L.alnStart;
        for (int i = 1; i < L.nb; i++) dst[i] = (int)ring[i * RS];
        dst[L.nb] = alnEnd;
        b.break_loc[r] = loc;
      }
    }
    b.read_st[r] = L.st_r;
    b.read_en[r] = (L.st_r != 0 && L.refPos == L.refEndM) ? L.rdEndM : 0;   // getReadPositionAtReferencePosition(alignmentEnd)
    b.maps_sum[r] = c.record_depth ? alnEnd - (L.alnStart - 1) - L.sumN : 0;   // positions M and D elements emitted
    b.err_sum[r] = c.record_depth ? L.errs : 0;
    b.rflags[r] = 0;
    if (RECORD) {
      uint4 h; h.x = L.entBase; h.y = L.blk0; h.z = L.blk1; h.w = L.blk2;
      b.fhdrA[r] = h;
      b.fmeta[r] = (unsigned)L.nblk | ((unsigned)L.njunc << 8) | ((unsigned)L.nextra << 16) | ((unsigned)L.ndbl << 24);
    }
    L.flags &= FWF_DONE;
  }
}

}  // namespace npt
