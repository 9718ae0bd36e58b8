// fast_inflate.h — raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for the BAM packer.
//
// Why: BAM decode is inflate-bound (profiles/NOTES.md: 250 MB/s per core with zlib); this decoder keeps a 64-bit bit buffer
// that is refilled with one unaligned 8-byte load, decodes literals / lengths with one 11-bit table lookup (sub-tables for
// the rare longer codes), offsets with an 8-bit one, and copies matches 8 bytes at a time.  It is deliberately strict: any
// condition it does not handle (incomplete or over-subscribed codes, input or output overrun, bad symbols) makes it return
// -1, and the caller falls back to zlib for that block, so the result is never worse than zlib's; the BGZF CRC32 is checked
// on top either way.
#pragma once
#include <cstdint>
#include <cstring>

namespace nptinf {

constexpr int LL_BITS = 11, OFF_BITS = 8, PRE_BITS = 7;
constexpr int LL_MAX_ENTRIES = (1 << LL_BITS) + 1024, OFF_MAX_ENTRIES = (1 << OFF_BITS) + 512;

// table entry: bits 0..4 = code length (full), bit 5 = literal, bit 6 = end of block, bit 7 = sub-table pointer,
// bits 8..12 = extra bits (or sub-table index bits), bits 16..31 = literal / base value / sub-table start
constexpr uint32_t F_LIT = 32, F_EOB = 64, F_SUB = 128;
constexpr uint32_t F_LIT2 = 1u << 13;   // literal entry that carries a second literal in bits 24..31 (length = both codes)

struct Tables {
  uint32_t ll[LL_MAX_ENTRIES];
  uint32_t off[OFF_MAX_ENTRIES];
};

static inline uint32_t bitrev(uint32_t c, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; }
  return r;
}

// Builds a decode table for `n` symbols with code lengths `lens` (0 = unused).  payload(sym) gives bits 5..31 of the entry.
// Returns false for over-subscribed or incomplete codes (except the RFC's single-code case, which zlib also accepts: left to zlib).
template <typename Payload>
static bool build_table(const uint8_t* lens, int n, int tbits, uint32_t* tab, int max_entries, Payload payload) {
  int count[16] = {0};
  for (int i = 0; i < n; i++) count[lens[i]]++;
  if (count[0] == n) return false;
  long left = 1;
  for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return false; }
  if (left != 0) return false;   // incomplete: zlib decides
  uint32_t next[16];
  {
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (l > 1 ? (uint32_t)count[l - 1] : 0u)) << 1; next[l] = code; }
  }
  const int tsize = 1 << tbits;
  for (int i = 0; i < tsize; i++) tab[i] = 0;
  // sub-table sizes: the longest code under every primary index
  uint8_t submax[1 << LL_BITS];
  memset(submax, 0, (size_t)tsize);
  {
    uint32_t nx[16];
    memcpy(nx, next, sizeof nx);
    for (int s = 0; s < n; s++) {
      const int l = lens[s];
      if (l > tbits) { const uint32_t r = bitrev(nx[l], l); const uint32_t p = r & (tsize - 1); if (l - tbits > submax[p]) submax[p] = (uint8_t)(l - tbits); }
      if (l) nx[l]++;
    }
  }
  int used = tsize;
  for (int p = 0; p < tsize; p++)
    if (submax[p]) {
      const int sb = submax[p];
      if (used + (1 << sb) > max_entries) return false;
      tab[p] = F_SUB | ((uint32_t)sb << 8) | ((uint32_t)used << 16) | (uint32_t)tbits;
      for (int k = 0; k < (1 << sb); k++) tab[used + k] = 0;
      used += 1 << sb;
    }
  for (int s = 0; s < n; s++) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t r = bitrev(next[l]++, l);
    const uint32_t e = payload(s) | (uint32_t)l;
    if (l <= tbits) {
      for (uint32_t k = r; k < (uint32_t)tsize; k += 1u << l) tab[k] = e;
    } else {
      const uint32_t p = r & (tsize - 1);
      const uint32_t sub = tab[p] >> 16;
      const int sb = (tab[p] >> 8) & 31;
      for (uint32_t k = r >> tbits; k < (1u << sb); k += 1u << (l - tbits)) tab[sub + k] = e;
    }
  }
  return true;
}

// Two literals per lookup: where a literal's code leaves enough of the 11 index bits to determine the NEXT code completely and
// that one is a literal too, the entry carries both (BAM blocks are literal-heavy: qualities and packed bases).
static inline void pair_literals(uint32_t* tab) {
  uint32_t single[1 << LL_BITS];
  memcpy(single, tab, sizeof single);
  for (uint32_t idx = 0; idx < (1u << LL_BITS); idx++) {
    const uint32_t e = single[idx];
    if (!(e & F_LIT) || (e & F_SUB)) continue;
    const int l1 = e & 31;
    if (l1 >= LL_BITS) continue;
    const uint32_t e2 = single[idx >> l1];
    const int l2 = e2 & 31;
    if ((e2 & F_LIT) && !(e2 & F_SUB) && l2 && l1 + l2 <= LL_BITS)
      tab[idx] = F_LIT | F_LIT2 | (e & 0x00FF0000u) | ((e2 & 0x00FF0000u) << 8) | (uint32_t)(l1 + l2);
  }
}

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t OFF_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t OFF_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t ll_payload(int s) {
  if (s < 256) return F_LIT | ((uint32_t)s << 16);
  if (s == 256) return F_EOB;
  if (s > 285) return 0xFFFF0000u | (31u << 8);   // invalid symbol: caught by the decoder (extra = 31)
  return ((uint32_t)LEN_BASE[s - 257] << 16) | ((uint32_t)LEN_EXTRA[s - 257] << 8);
}
static inline uint32_t off_payload(int s) {
  if (s > 29) return 0xFFFF0000u | (31u << 8);
  return ((uint32_t)OFF_BASE[s] << 16) | ((uint32_t)OFF_EXTRA[s] << 8);
}

static inline bool build_fixed(Tables& t) {
  uint8_t l[288];
  for (int i = 0; i < 144; i++) l[i] = 8;
  for (int i = 144; i < 256; i++) l[i] = 9;
  for (int i = 256; i < 280; i++) l[i] = 7;
  for (int i = 280; i < 288; i++) l[i] = 8;
  uint8_t d[32];
  for (int i = 0; i < 32; i++) d[i] = 5;
  if (!(build_table(l, 288, LL_BITS, t.ll, LL_MAX_ENTRIES, ll_payload) && build_table(d, 32, OFF_BITS, t.off, OFF_MAX_ENTRIES, off_payload))) return false;
  pair_literals(t.ll);
  return true;
}

struct BitReader {
  const uint8_t* p; const uint8_t* end;
  uint64_t buf = 0; int left = 0;
  long overrun = 0;   // zero bytes supplied beyond the end of the input
  inline void refill() {
    if (end - p >= 8) {
      uint64_t v; memcpy(&v, p, 8);
      buf |= v << left;
      p += (63 - left) >> 3;
      left |= 56;
    } else {
      while (left <= 56) {
        if (p < end) buf |= (uint64_t)*p++ << left; else overrun++;
        left += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void skip(int n) { buf >>= n; left -= n; }
};

// Returns the number of bytes written (the whole stream up to its final block), or -1.
static long inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap) {
  static Tables fixed;
  static const bool fixed_ok = build_fixed(fixed);
  if (!fixed_ok) return -1;
  Tables dyn;
  BitReader br{in, in + in_len};
  uint8_t* o = out;
  uint8_t* const oend = out + out_cap;
  for (;;) {
    br.refill();
    const uint32_t hdr = br.peek(3); br.skip(3);
    const int last = hdr & 1, type = hdr >> 1;
    const Tables* T = nullptr;
    if (type == 0) {   // stored: back to a byte boundary of the input
      br.skip(br.left & 7);
      const long inbuf = (long)(br.left >> 3) - br.overrun;   // real whole bytes still unread in the bit buffer
      if (inbuf < 0) return -1;
      const uint8_t* q = br.p - inbuf;
      br.overrun = 0;
      if (br.end - q < 4) return -1;
      const unsigned len = q[0] | (q[1] << 8), nlen = q[2] | (q[3] << 8);
      if ((len ^ 0xFFFFu) != nlen) return -1;
      q += 4;
      if ((size_t)(br.end - q) < len || (size_t)(oend - o) < len) return -1;
      memcpy(o, q, len);
      o += len;
      br.p = q + len; br.buf = 0; br.left = 0;
      if (last) break;
      continue;
    } else if (type == 1) {
      T = &fixed;
    } else if (type == 2) {
      const int hlit = br.peek(5) + 257; br.skip(5);
      const int hdist = br.peek(5) + 1; br.skip(5);
      const int hclen = br.peek(4) + 4; br.skip(4);
      if (hlit > 286 || hdist > 30) return -1;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t pl[19] = {0};
      br.refill();
      for (int i = 0; i < hclen; i++) { if (br.left < 3) br.refill(); pl[order[i]] = (uint8_t)br.peek(3); br.skip(3); }
      uint32_t pre[1 << PRE_BITS];
      if (!build_table(pl, 19, PRE_BITS, pre, 1 << PRE_BITS, [](int s) { return (uint32_t)s << 16; })) return -1;
      uint8_t lens[320];
      int i = 0;
      const int total = hlit + hdist;
      while (i < total) {
        br.refill();
        const uint32_t e = pre[br.peek(PRE_BITS)];
        const int l = e & 31;
        if (!l) return -1;
        br.skip(l);
        const int s = e >> 16;
        if (s < 16) lens[i++] = (uint8_t)s;
        else {
          int rep, v = 0;
          if (s == 16) { if (i == 0) return -1; v = lens[i - 1]; rep = 3 + br.peek(2); br.skip(2); }
          else if (s == 17) { rep = 3 + br.peek(3); br.skip(3); }
          else { rep = 11 + br.peek(7); br.skip(7); }
          if (i + rep > total) return -1;
          while (rep--) lens[i++] = (uint8_t)v;
        }
      }
      if (lens[256] == 0) return -1;
      if (!build_table(lens, hlit, LL_BITS, dyn.ll, LL_MAX_ENTRIES, ll_payload)) return -1;
      if (!build_table(lens + hlit, hdist, OFF_BITS, dyn.off, OFF_MAX_ENTRIES, off_payload)) return -1;
      pair_literals(dyn.ll);
      T = &dyn;
    } else {
      return -1;
    }
    // ---- symbols of a Huffman block
    // Fast loop: while at least 16 input bytes and 320 output bytes remain nothing inside needs a bounds check (a refill reads
    // 8 bytes, a symbol writes at most 258 bytes + 7 of slack for the 8-byte copies ... which stay inside this block's output:
    // 320 > 258 + 8 + 6 literals).  Up to three literal entries are taken per refill (3 x 15 bits < 56).
    bool eob = false;
    while (br.end - br.p >= 16 && oend - o >= 320) {
      br.refill();
      uint32_t e = T->ll[br.peek(LL_BITS)];
      if (e & F_SUB) e = T->ll[(e >> 16) + ((uint32_t)(br.buf >> LL_BITS) & ((1u << ((e >> 8) & 31)) - 1))];
      if (e & F_LIT) {   // one or two literals per entry, three entries per refill
        br.skip(e & 31); o[0] = (uint8_t)(e >> 16); o[1] = (uint8_t)(e >> 24); o += 1 + ((e >> 13) & 1);
        e = T->ll[br.peek(LL_BITS)];
        if (e & F_SUB) e = T->ll[(e >> 16) + ((uint32_t)(br.buf >> LL_BITS) & ((1u << ((e >> 8) & 31)) - 1))];
        if (e & F_LIT) {
          br.skip(e & 31); o[0] = (uint8_t)(e >> 16); o[1] = (uint8_t)(e >> 24); o += 1 + ((e >> 13) & 1);
          e = T->ll[br.peek(LL_BITS)];
          if (e & F_SUB) e = T->ll[(e >> 16) + ((uint32_t)(br.buf >> LL_BITS) & ((1u << ((e >> 8) & 31)) - 1))];
          if (e & F_LIT) { br.skip(e & 31); o[0] = (uint8_t)(e >> 16); o[1] = (uint8_t)(e >> 24); o += 1 + ((e >> 13) & 1); continue; }
        }
        if (br.left < 48) br.refill();   // the literals used up to 30 bits: top up before a length / offset pair (48 bits)
      }
      const int l = e & 31;
      if (!l) return -1;
      br.skip(l);
      if (e & F_EOB) { eob = true; break; }
      const int xb = (e >> 8) & 31;
      if (xb == 31) return -1;
      unsigned len = (e >> 16) + br.peek(xb);
      br.skip(xb);
      uint32_t d = T->off[br.peek(OFF_BITS)];
      if (d & F_SUB) d = T->off[(d >> 16) + ((uint32_t)(br.buf >> OFF_BITS) & ((1u << ((d >> 8) & 31)) - 1))];
      const int dl = d & 31;
      if (!dl) return -1;
      br.skip(dl);
      const int dx = (d >> 8) & 31;
      if (dx == 31) return -1;
      const unsigned dist = (d >> 16) + br.peek(dx);
      br.skip(dx);
      if (dist > (size_t)(o - out)) return -1;
      const uint8_t* src = o - dist;
      uint8_t* const stop = o + len;
      if (dist >= 8) {
        do { uint64_t v; memcpy(&v, src, 8); memcpy(o, &v, 8); src += 8; o += 8; } while (o < stop);   // may write up to 7 bytes of slack
        o = stop;
      } else if (dist == 1) {
        memset(o, *src, len);
        o = stop;
      } else {
        while (o < stop) *o++ = *src++;
      }
    }
    // ---- careful loop for the ends of the input / output
    while (!eob) {
      if (br.left < 48) br.refill();
      uint32_t e = T->ll[br.peek(LL_BITS)];
      if (e & F_SUB) e = T->ll[(e >> 16) + ((uint32_t)(br.buf >> LL_BITS) & ((1u << ((e >> 8) & 31)) - 1))];
      const int l = e & 31;
      if (!l) return -1;
      br.skip(l);
      if (e & F_LIT) {
        if (o == oend) return -1;
        *o++ = (uint8_t)(e >> 16);
        if (e & F_LIT2) { if (o == oend) return -1; *o++ = (uint8_t)(e >> 24); }
        continue;
      }
      if (e & F_EOB) { eob = true; break; }
      const int xb = (e >> 8) & 31;
      if (xb == 31) return -1;
      unsigned len = (e >> 16) + br.peek(xb);
      br.skip(xb);
      uint32_t d = T->off[br.peek(OFF_BITS)];
      if (d & F_SUB) d = T->off[(d >> 16) + ((uint32_t)(br.buf >> OFF_BITS) & ((1u << ((d >> 8) & 31)) - 1))];
      const int dl = d & 31;
      if (!dl) return -1;
      br.skip(dl);
      const int dx = (d >> 8) & 31;
      if (dx == 31) return -1;
      const unsigned dist = (d >> 16) + br.peek(dx);
      br.skip(dx);
      if (dist > (size_t)(o - out) || len > (size_t)(oend - o)) return -1;
      const uint8_t* src = o - dist;
      if (dist >= 8) {
        while (len >= 8) { uint64_t v; memcpy(&v, src, 8); memcpy(o, &v, 8); src += 8; o += 8; len -= 8; }
        while (len--) *o++ = *src++;
      } else if (dist == 1) {
        memset(o, *src, len);
        o += len;
      } else {
        while (len--) *o++ = *src++;
      }
    }
    if (br.overrun * 8 > br.left) return -1;   // symbols were decoded from bits that are not there
    if (last) break;
  }
  if (br.overrun * 8 > br.left) return -1;
  return (long)(o - out);
}

}  // namespace nptinf
