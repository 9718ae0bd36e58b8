// BAM (BGZF) -> structure-of-arrays packer: see include/npt_bam.h for the reference lines this replaces.
//
// Layout of the work: the file is memory-mapped; BGZF block headers are scanned sequentially (18 bytes each), a window of
// blocks is inflated in parallel (one zlib raw-inflate stream per thread, each block straight into its place of one
// contiguous buffer: the ISIZE trailers give the offsets), and records are parsed from that buffer in file order.  A BAM
// record already holds CIGAR as (len<<4)|op words and SEQ as 4-bit codes, so every field is one memcpy.
#include <zlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/npt_bam.h"
#include "fast_inflate.h"

namespace {

constexpr size_t WINDOW_BLOCKS = 4096;  // BGZF blocks inflated per refill (<= 256 MB of BAM bytes)

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

template <typename T>
struct Buf {  // grow-only array from the caller's allocator
  T* p = nullptr;
  size_t cap = 0;
  void* (*alloc)(size_t) = nullptr;
  void (*free_)(void*) = nullptr;
  bool reserve(size_t n, size_t keep) {
    if (n <= cap) return true;
    size_t nc = std::max<size_t>(n, cap + cap / 2 + 1024);
    T* q = (T*)alloc(nc * sizeof(T));
    if (!q) return false;
    if (keep) memcpy(q, p, keep * sizeof(T));
    if (p) free_(p);
    p = q; cap = nc;
    return true;
  }
  void release() { if (p) free_(p); p = nullptr; cap = 0; }
};

struct BlockRef { size_t coff; uint32_t csize, usize; size_t uoff; };

// one scanned record: where it is, and what the history-free part of the filter chain says about it
enum : uint8_t { KEEP, UNMAPPED, EXCLUDED, LOWQ, SHORT, STOP, LOWMAPQ, SECONDARY, MALFORMED };
struct RecInfo { size_t off; double q1; int32_t refID; uint32_t n_cig, sb, nl; uint8_t what; bool lowmapq, secondary; };

// fn(lo, hi) over [0, count) on up to nthreads threads (inline when the range is small)
template <typename F>
void parallel_for(size_t count, int nthreads, F fn) {
  const size_t nt = std::min<size_t>((size_t)std::max(1, nthreads), count / 2048 + 1);
  if (nt <= 1) { if (count) fn(0, count); return; }
  std::vector<std::thread> th;
  const size_t per = (count + nt - 1) / nt;
  for (size_t t = 1; t < nt; t++) th.emplace_back([=]() { fn(std::min(count, t * per), std::min(count, (t + 1) * per)); });
  fn(0, std::min(count, per));
  for (auto& x : th) x.join();
}

}  // namespace

struct nptb_reader {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t fsize = 0, cpos = 0;
  struct RawBuf {   // grow-only byte buffer that is never zero-filled (std::vector::resize would touch every new byte)
    uint8_t* p = nullptr; size_t cap = 0;
    uint8_t* data() { return p; }
    bool reserve(size_t n, size_t keep) {
      if (n <= cap) return true;
      const size_t nc = n + n / 8 + 4096;
      uint8_t* q = (uint8_t*)malloc(nc);
      if (!q) return false;
      if (keep) memcpy(q, p, keep);
      free(p);
      p = q; cap = nc;
      return true;
    }
    ~RawBuf() { free(p); }
  } ubuf;
  size_t upos = 0, uend = 0;
  bool ceof = false;
  int nthreads = 1;
  std::vector<std::string> ref_names;
  std::vector<int64_t> ref_lens;
  std::string err;
  // batch buffers
  Buf<int32_t> pos; Buf<uint16_t> flag, src; Buf<uint32_t> cigar_off, cigar, name_off; Buf<uint64_t> seq_off; Buf<uint8_t> seq4, mapq;
  Buf<double> q1; Buf<char> names;
  // compact transfer format of the current batch (nptb_pack)
  Buf<uint16_t> cigar16; Buf<uint32_t> wide_idx, wide_word; Buf<uint8_t> seq2, exc_val; Buf<uint64_t> exc_byte;
  std::vector<RecInfo> scan;
  std::vector<uint32_t> kept;
  // totals
  int64_t n_records = 0, n_unmapped = 0, n_excluded_ref = 0, n_low_quality = 0, n_short = 0, n_low_mapq = 0, n_secondary = 0;
  int64_t num_reads = 0;  // "numReads" of the reference's loop
  double err_prob[256];   // pow(10, -q/10) for every byte value: the same calls the per-base loop would make, made once
  bool stopped = false;
  bool fast_inflate = true;   // NPTB_ZLIB_ONLY=1 turns the table-driven decoder off (A/B, debugging)
};

namespace {

int fail(nptb_reader* r, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  r->err = buf;
  return -1;
}

// Inflate the next window of BGZF blocks behind the unread bytes.  Returns 0 ok (possibly nothing added at EOF), -1 error.
int refill(nptb_reader* r) {
  if (r->upos > 0) {  // keep the unread tail at the front
    const size_t left = r->uend - r->upos;
    if (left) memmove(r->ubuf.data(), r->ubuf.data() + r->upos, left);
    r->upos = 0; r->uend = left;
  }
  std::vector<BlockRef> blocks;
  size_t uoff = r->uend;
  while (blocks.size() < WINDOW_BLOCKS && r->cpos < r->fsize) {
    const uint8_t* h = r->map + r->cpos;
    if (r->fsize - r->cpos < 28) return fail(r, "truncated BGZF block header at offset %zu", r->cpos);
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return fail(r, "not a BGZF block at offset %zu", r->cpos);
    const unsigned xlen = le16(h + 10);
    if (r->fsize - r->cpos < 12 + (size_t)xlen + 8) return fail(r, "truncated BGZF extra field at offset %zu", r->cpos);
    int bsize = -1;
    for (unsigned o = 0; o + 4 <= xlen;) {
      const uint8_t* x = h + 12 + o;
      const unsigned slen = le16(x + 2);
      if (x[0] == 'B' && x[1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = le16(x + 4);
      o += 4 + slen;
    }
    if (bsize < 0) return fail(r, "gzip member without a BGZF BC field at offset %zu", r->cpos);
    const size_t total = (size_t)bsize + 1;
    if (total < 12 + (size_t)xlen + 8 || r->fsize - r->cpos < total) return fail(r, "truncated BGZF block at offset %zu", r->cpos);
    const uint32_t isize = le32(h + total - 4);
    if (isize > 65536) return fail(r, "BGZF block larger than 64 KiB at offset %zu", r->cpos);
    blocks.push_back({r->cpos + 12 + xlen, (uint32_t)(total - 12 - xlen - 8), isize, uoff});
    uoff += isize;
    r->cpos += total;
  }
  if (r->cpos >= r->fsize) r->ceof = true;
  if (blocks.empty()) return 0;
  if (!r->ubuf.reserve(uoff + 16, r->uend)) return fail(r, "out of memory");
  std::atomic<size_t> next{0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= blocks.size() || bad.load()) break;
      const BlockRef& b = blocks[i];
      if (b.usize == 0) continue;  // the empty EOF marker block
      const uint32_t crc = le32(r->map + b.coff + b.csize);
      // the table-driven decoder first (fast_inflate.h); anything it refuses, or a CRC that does not match, goes to zlib
      bool done = false;
      if (r->fast_inflate && nptinf::inflate_raw(r->map + b.coff, b.csize, r->ubuf.data() + b.uoff, b.usize) == (long)b.usize)
        done = (uint32_t)crc32(crc32(0L, Z_NULL, 0), r->ubuf.data() + b.uoff, b.usize) == crc;
      if (done) continue;
      zs.next_in = const_cast<Bytef*>(r->map + b.coff); zs.avail_in = b.csize;
      zs.next_out = r->ubuf.data() + b.uoff; zs.avail_out = b.usize;
      const int rc = inflate(&zs, Z_FINISH);
      if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = 1; break; }
      if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), r->ubuf.data() + b.uoff, b.usize) != crc) { bad = 2; break; }
      inflateReset(&zs);
    }
    inflateEnd(&zs);
  };
  const int nt = (int)std::min<size_t>((size_t)r->nthreads, blocks.size());
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (bad.load()) return fail(r, bad.load() == 2 ? "BGZF block fails its CRC32" : "corrupt deflate data in a BGZF block");
  r->uend = uoff;
  return 0;
}

// n contiguous unread bytes, or nullptr at end of file (*short_read set if some but fewer than n bytes remain)
const uint8_t* peek(nptb_reader* r, size_t n, bool* short_read, int* rc) {
  *short_read = false; *rc = 0;
  while (r->uend - r->upos < n) {
    if (r->ceof) { *short_read = r->uend > r->upos; return nullptr; }
    if (refill(r) < 0) { *rc = -1; return nullptr; }
  }
  return r->ubuf.data() + r->upos;
}

int read_header(nptb_reader* r) {
  bool sh; int rc;
  const uint8_t* p = peek(r, 12, &sh, &rc);
  if (!p) return rc ? rc : fail(r, "file too short for a BAM header");
  if (memcmp(p, "BAM\1", 4) != 0) return fail(r, "not a BAM file (magic)");
  const uint32_t l_text = le32(p + 4);
  p = peek(r, 12 + (size_t)l_text, &sh, &rc);
  if (!p) return rc ? rc : fail(r, "truncated BAM header text");
  const uint32_t n_ref = le32(p + 8 + l_text);
  r->upos += 12 + (size_t)l_text;
  for (uint32_t i = 0; i < n_ref; i++) {
    p = peek(r, 4, &sh, &rc);
    if (!p) return rc ? rc : fail(r, "truncated BAM reference list");
    const uint32_t l_name = le32(p);
    p = peek(r, 8 + (size_t)l_name, &sh, &rc);
    if (!p) return rc ? rc : fail(r, "truncated BAM reference list");
    r->ref_names.emplace_back((const char*)p + 4, l_name ? l_name - 1 : 0);
    r->ref_lens.push_back(le32(p + 4 + l_name));
    r->upos += 8 + (size_t)l_name;
  }
  return 0;
}

}  // namespace

extern "C" {

nptb_reader* nptb_open(const char* path, int nthreads, void* (*alloc)(size_t), void (*free_)(void*), char* err, size_t errlen) {
  auto say = [&](const char* m) { if (err && errlen) snprintf(err, errlen, "%s: %s", path ? path : "(null)", m); };
  if (!path) { say("null path"); return nullptr; }
  if ((alloc == nullptr) != (free_ == nullptr)) { say("alloc and free must be given together"); return nullptr; }
  nptb_reader* r = new nptb_reader();
  r->fd = open(path, O_RDONLY);
  struct stat st;
  if (r->fd < 0 || fstat(r->fd, &st) != 0) { say("cannot open"); if (r->fd >= 0) close(r->fd); delete r; return nullptr; }
  r->fsize = (size_t)st.st_size;
  if (r->fsize) {
    void* m = mmap(nullptr, r->fsize, PROT_READ, MAP_PRIVATE, r->fd, 0);
    if (m == MAP_FAILED) { say("mmap failed"); close(r->fd); delete r; return nullptr; }
    r->map = (const uint8_t*)m;
    madvise(m, r->fsize, MADV_SEQUENTIAL);
  }
  r->nthreads = nthreads > 0 ? nthreads : (int)std::max(1u, std::thread::hardware_concurrency());
  if (const char* e = getenv("NPTB_ZLIB_ONLY")) r->fast_inflate = atoi(e) == 0;
  void* (*a)(size_t) = alloc ? alloc : malloc;
  void (*f)(void*) = free_ ? free_ : free;
#define INIT(b) r->b.alloc = a; r->b.free_ = f
  INIT(pos); INIT(flag); INIT(src); INIT(cigar_off); INIT(cigar); INIT(name_off); INIT(seq_off); INIT(seq4); INIT(mapq); INIT(q1); INIT(names);
  INIT(cigar16); INIT(wide_idx); INIT(wide_word); INIT(seq2); INIT(exc_val); INIT(exc_byte);
#undef INIT
  for (int v = 0; v < 256; v++) r->err_prob[v] = std::pow(10.0, -(double)(int8_t)(uint8_t)v / 10.0);
  if (read_header(r) < 0) { say(r->err.c_str()); nptb_close(r); return nullptr; }
  return r;
}

void nptb_close(nptb_reader* r) {
  if (!r) return;
  r->pos.release(); r->flag.release(); r->src.release(); r->cigar_off.release(); r->cigar.release(); r->name_off.release();
  r->seq_off.release(); r->seq4.release(); r->mapq.release(); r->q1.release(); r->names.release();
  r->cigar16.release(); r->wide_idx.release(); r->wide_word.release(); r->seq2.release(); r->exc_val.release(); r->exc_byte.release();
  if (r->map) munmap((void*)r->map, r->fsize);
  if (r->fd >= 0) close(r->fd);
  delete r;
}

int32_t nptb_n_refs(const nptb_reader* r) { return r ? (int32_t)r->ref_names.size() : 0; }
const char* nptb_ref_name(const nptb_reader* r, int32_t i) { return (r && i >= 0 && i < (int32_t)r->ref_names.size()) ? r->ref_names[i].c_str() : nullptr; }
int64_t nptb_ref_len(const nptb_reader* r, int32_t i) { return (r && i >= 0 && i < (int32_t)r->ref_lens.size()) ? r->ref_lens[i] : -1; }
const char* nptb_last_error(const nptb_reader* r) { return r ? r->err.c_str() : "null reader"; }

int nptb_next_batch(nptb_reader* r, const nptb_filter* f, int32_t src, int64_t max_batch_reads, nptb_batch* out) {
  if (!r || !f || !out || max_batch_reads < 1) return r ? fail(r, "bad argument") : -1;
  memset(out, 0, sizeof *out);
  int64_t n = 0;
  size_t ncig = 0, nseq = 0, nname = 0;
  int32_t ref = -2;
  const int32_t nrefs = (int32_t)r->ref_names.size();
  if (!r->cigar_off.reserve(1, 0) || !r->seq_off.reserve(1, 0) || !r->name_off.reserve(1, 0)) return fail(r, "out of memory");
  r->cigar_off.p[0] = 0; r->seq_off.p[0] = 0; r->name_off.p[0] = 0;
  typedef RecInfo Info;
  std::vector<Info>& recs = r->scan;
  std::vector<uint32_t>& kept = r->kept;
  bool batch_closed = false;
  while (n < max_batch_reads && !r->stopped && !batch_closed) {
    // ---- 1. the complete records at hand, in file order (one 4-byte hop each)
    bool sh; int rc;
    const uint8_t* p = peek(r, 4, &sh, &rc);
    if (!p) { if (rc) return rc; if (sh) return fail(r, "truncated BAM record"); break; }
    const uint32_t bs0 = le32(p);
    if (bs0 < 32) return fail(r, "BAM record shorter than its fixed fields");
    if (!peek(r, 4 + (size_t)bs0, &sh, &rc)) return rc ? rc : fail(r, "truncated BAM record");
    recs.clear();
    const size_t want = (size_t)(max_batch_reads - n) + 1024;  // a few more than fit: some will be filtered out
    for (size_t o = r->upos; recs.size() < want && r->uend - o >= 4;) {
      const uint32_t bs = le32(r->ubuf.data() + o);
      if (bs < 32) return fail(r, "BAM record shorter than its fixed fields");
      if (r->uend - o < 4 + (size_t)bs) break;
      Info in{}; in.off = o;
      recs.push_back(in);
      o += 4 + (size_t)bs;
    }
    // ---- 2. per record, in parallel: the part of the filter chain that needs no history, and q1
    auto classify = [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++) {
        Info& in = recs[i];
        const uint8_t* q = r->ubuf.data() + in.off;
        const uint32_t bs = le32(q);
        in.refID = (int32_t)le32(q + 4);
        const unsigned l_name = q[12], mq = q[13], n_cig = le16(q + 16), flg = le16(q + 18);
        const uint32_t l_seq = le32(q + 20);
        const size_t need = 32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
        if (need > bs) { in.what = MALFORMED; continue; }
        const uint8_t* qual = q + 36 + l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2;
        in.n_cig = n_cig; in.sb = (l_seq + 1) / 2; in.nl = l_name ? l_name - 1 : 0;
        in.q1 = 500; in.what = KEEP; in.lowmapq = false; in.secondary = false;
        if (flg & 0x4) { in.what = UNMAPPED; continue; }                                                             // :716
        if (f->ref_include && (in.refID < 0 || in.refID >= nrefs || !f->ref_include[in.refID])) { in.what = EXCLUDED; continue; }  // :721-730
        if (l_seq > 0 && qual[0] != 0xFF) {                                                                          // :738-752
          double sump = 0;
          for (uint32_t k = 0; k < l_seq; k++) sump += r->err_prob[qual[k]];  // Math.pow(10, -b[i]/10.0), b[i] a signed byte
          sump = sump / (double)l_seq;
          in.q1 = -10.0 * std::log10(sump);
        }
        if (in.q1 < f->fail_thresh) { in.what = LOWQ; continue; }
        if (l_seq <= 1) { in.what = SHORT; continue; }                                                               // :797
        in.lowmapq = (int)mq < f->min_mapq;                                                                          // :813
        in.secondary = (flg & 0x900) != 0;                                                                           // :938-959
      }
    };
    parallel_for(recs.size(), r->nthreads, classify);
    // ---- 3. in file order: the stateful part (numReads / maxReads, chromosome switch, batch capacity) and output offsets.
    // A record that opens the next chromosome, or does not fit, is left unread (and uncounted) for the next call.
    kept.clear();
    size_t consumed_end = r->upos;
    const int64_t n0 = n;
    const size_t ncig0 = ncig, nseq0 = nseq, nname0 = nname;
    for (size_t i = 0; i < recs.size(); i++) {
      Info& in = recs[i];
      if (in.what == MALFORMED) return fail(r, "BAM record fields exceed its block_size");
      uint8_t what = in.what;
      bool counted = false, pre = false;
      if (what == KEEP) {
        counted = true;                                                                                              // :804-808
        if (f->max_reads > 0 && r->num_reads + 1 > f->max_reads) what = STOP;
        else if (in.lowmapq) what = LOWMAPQ;
        else { pre = true; if (in.secondary) what = SECONDARY; }
      }
      if (pre && n > 0 && in.refID != ref) { batch_closed = true; break; }            // chromosome switch (:830)
      if (what == KEEP && (n >= max_batch_reads || ncig + in.n_cig >= (1ull << 32))) { batch_closed = true; break; }
      if (pre && n == 0) ref = in.refID;
      if (counted) r->num_reads++;
      if (what == STOP) { r->stopped = true; break; }
      r->n_records++;
      const uint8_t* q = r->ubuf.data() + in.off;
      consumed_end = in.off + 4 + (size_t)le32(q);
      switch (what) {
        case UNMAPPED: r->n_unmapped++; break;
        case EXCLUDED: r->n_excluded_ref++; break;
        case LOWQ: r->n_low_quality++; break;
        case SHORT: r->n_short++; break;
        case LOWMAPQ: r->n_low_mapq++; break;
        case SECONDARY: r->n_secondary++; break;
        default: break;
      }
      if (what != KEEP) continue;
      in.what = KEEP;
      kept.push_back((uint32_t)i);
      n++; ncig += in.n_cig; nseq += in.sb; nname += in.nl;
    }
    // ---- 4. copy the kept records, in parallel (every field is a memcpy)
    if (!kept.empty()) {
      // Size the batch arrays once for where the batch is heading (what is left of the file at the rate seen so far, capped by
      // max_batch_reads) instead of growing them window by window: every growth copies the batch and faults in fresh pages
      // (1.6 s of a 3.2 s run with 1 Mi-read batches before this).
      {
        const double frac = r->fsize ? std::min(1.0, std::max(1e-6, (double)r->cpos / (double)r->fsize)) : 1.0;
        const double n_file = (double)std::max<int64_t>(r->n_records, 1) / frac;                // records of the whole file, roughly
        const double share = (double)n / (double)std::max<int64_t>(r->n_records + (int64_t)recs.size(), 1);   // of which kept
        const int64_t n_goal = std::min<int64_t>(max_batch_reads, n + (int64_t)(1.1 * share * std::max(0.0, n_file - (double)r->n_records)) + 1024);
        const double per = (double)n_goal / (double)n * 1.05;
        if ((size_t)n_goal > r->pos.cap) {
          r->pos.reserve(n_goal, n0); r->flag.reserve(n_goal, n0); r->src.reserve(n_goal, n0); r->mapq.reserve(n_goal, n0); r->q1.reserve(n_goal, n0);
          r->cigar_off.reserve(n_goal + 1, n0 + 1); r->seq_off.reserve(n_goal + 1, n0 + 1); r->name_off.reserve(n_goal + 1, n0 + 1);
        }
        if (ncig > r->cigar.cap) r->cigar.reserve((size_t)(ncig * per) + 1024, ncig0);
        if (nseq + 16 > r->seq4.cap) r->seq4.reserve((size_t)(nseq * per) + 4096, nseq0);
        if (nname + 1 > r->names.cap) r->names.reserve((size_t)(nname * per) + 1024, nname0);
      }
      if (!r->pos.reserve(n, n0) || !r->flag.reserve(n, n0) || !r->src.reserve(n, n0) || !r->mapq.reserve(n, n0) || !r->q1.reserve(n, n0) ||
          !r->cigar_off.reserve(n + 1, n0 + 1) || !r->seq_off.reserve(n + 1, n0 + 1) || !r->name_off.reserve(n + 1, n0 + 1) ||
          !r->cigar.reserve(ncig, ncig0) || !r->seq4.reserve(nseq + 16, nseq0) || !r->names.reserve(nname + 1, nname0))
        return fail(r, "out of memory");
      // sequential prefix of the three offset arrays (cheap), then the copies
      {
        size_t c = ncig0, s2 = nseq0, m = nname0;
        for (size_t k = 0; k < kept.size(); k++) {
          const Info& in = recs[kept[k]];
          c += in.n_cig; s2 += in.sb; m += in.nl;
          r->cigar_off.p[n0 + k + 1] = (uint32_t)c; r->seq_off.p[n0 + k + 1] = s2; r->name_off.p[n0 + k + 1] = (uint32_t)m;
        }
      }
      auto copy = [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
          const Info& in = recs[kept[k]];
          const size_t j = (size_t)n0 + k;
          const uint8_t* q = r->ubuf.data() + in.off;
          const unsigned l_name = q[12];
          const uint8_t* name = q + 36;
          const uint8_t* cig = name + l_name;
          const uint8_t* seq = cig + 4 * (size_t)in.n_cig;
          r->pos.p[j] = (int32_t)le32(q + 8) + 1; r->flag.p[j] = le16(q + 18); r->src.p[j] = (uint16_t)src; r->mapq.p[j] = q[13]; r->q1.p[j] = in.q1;
          memcpy(r->cigar.p + r->cigar_off.p[j], cig, 4 * (size_t)in.n_cig);
          memcpy(r->seq4.p + r->seq_off.p[j], seq, in.sb);
          memcpy(r->names.p + r->name_off.p[j], name, in.nl);
        }
      };
      parallel_for(kept.size(), r->nthreads, copy);
    }
    r->upos = consumed_end;
  }
  out->ref_id = ref == -2 ? -1 : ref;
  out->n = n;
  out->pos = r->pos.p; out->flag = r->flag.p; out->src = r->src.p; out->cigar_off = r->cigar_off.p; out->cigar = r->cigar.p;
  out->seq_off = r->seq_off.p; out->seq4 = r->seq4.p; out->q1 = r->q1.p; out->mapq = r->mapq.p; out->name_off = r->name_off.p; out->names = r->names.p;
  out->n_records = r->n_records; out->n_unmapped = r->n_unmapped; out->n_excluded_ref = r->n_excluded_ref; out->n_low_quality = r->n_low_quality;
  out->n_short = r->n_short; out->n_low_mapq = r->n_low_mapq; out->n_secondary = r->n_secondary;
  out->stopped_at_max_reads = r->stopped ? 1 : 0;
  return n > 0 ? 1 : 0;
}

// ---- compact transfer format (npt_packed_batch of include/npt.h): every output element depends on one input element, so the
// arrays are cut into ranges for the thread pool; the (rare) wide elements and exception bytes are collected per range and
// concatenated in order
int nptb_pack(nptb_reader* r, const nptb_batch* b, nptb_packed* out) {
  if (!r || !b || !out) return -1;
  const int64_t n = b->n;
  const size_t ncig = n ? b->cigar_off[n] : 0, nseq = n ? (size_t)b->seq_off[n] : 0, nseq2 = (nseq + 1) / 2;
  if (!r->cigar16.reserve(ncig + 8, 0) || !r->seq2.reserve(nseq2 + 16, 0)) return fail(r, "out of memory");
  const int nt = std::max(1, r->nthreads);
  // CIGAR: (len << 4 | op) fits 16 bits when len < 4095; longer elements keep their op under an all-ones length
  std::vector<std::vector<uint32_t>> widx(nt), wword(nt);
  {
    const size_t per = (ncig + nt - 1) / nt;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        const size_t lo = std::min(ncig, t * per), hi = std::min(ncig, (t + 1) * per);
        for (size_t i = lo; i < hi; i++) {
          const uint32_t w = b->cigar[i];
          if ((w >> 4) >= 4095u) { r->cigar16.p[i] = (uint16_t)(0xFFF0u | (w & 15u)); widx[t].push_back((uint32_t)i); wword[t].push_back(w); }
          else r->cigar16.p[i] = (uint16_t)w;
        }
      });
    for (auto& x : th) x.join();
  }
  // bases: seq4 byte j (two 4-bit codes) -> nibble j of the seq2 stream (two 2-bit codes); A C G T = 1 2 4 8 -> 0 1 2 3
  static const uint8_t code2[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0};
  static const uint8_t ok4[16] = {0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
  std::vector<std::vector<uint64_t>> eidx(nt);
  std::vector<std::vector<uint8_t>> eval(nt);
  {
    const size_t per = ((nseq2 + nt - 1) / nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        const size_t lo = std::min(nseq2, t * per), hi = std::min(nseq2, (t + 1) * per);
        for (size_t m = lo; m < hi; m++) {
          uint8_t o = 0;
          for (int k = 0; k < 2; k++) {
            const size_t j = 2 * m + k;
            uint8_t v = 0;
            if (j < nseq) {
              v = b->seq4[j];
              if (!(ok4[v >> 4] & ok4[v & 15])) { eidx[t].push_back((uint64_t)j); eval[t].push_back(v); }
            }
            o = (uint8_t)((o << 4) | (code2[v >> 4] << 2) | code2[v & 15]);
          }
          r->seq2.p[m] = o;
        }
      });
    for (auto& x : th) x.join();
  }
  size_t nw = 0, ne = 0;
  for (int t = 0; t < nt; t++) { nw += widx[t].size(); ne += eidx[t].size(); }
  if (!r->wide_idx.reserve(nw + 1, 0) || !r->wide_word.reserve(nw + 1, 0) || !r->exc_byte.reserve(ne + 1, 0) || !r->exc_val.reserve(ne + 1, 0))
    return fail(r, "out of memory");
  size_t ow = 0, oe = 0;
  for (int t = 0; t < nt; t++) {
    if (!widx[t].empty()) { memcpy(r->wide_idx.p + ow, widx[t].data(), widx[t].size() * 4); memcpy(r->wide_word.p + ow, wword[t].data(), wword[t].size() * 4); ow += widx[t].size(); }
    if (!eidx[t].empty()) { memcpy(r->exc_byte.p + oe, eidx[t].data(), eidx[t].size() * 8); memcpy(r->exc_val.p + oe, eval[t].data(), eval[t].size()); oe += eidx[t].size(); }
  }
  out->n = n; out->pos = b->pos; out->flag = b->flag; out->src = b->src; out->cigar_off = b->cigar_off; out->cigar16 = r->cigar16.p;
  out->n_wide = (int64_t)nw; out->wide_idx = r->wide_idx.p; out->wide_word = r->wide_word.p;
  out->seq_off = b->seq_off; out->seq2 = r->seq2.p; out->n_exc = (int64_t)ne; out->exc_byte = r->exc_byte.p; out->exc_val = r->exc_val.p;
  return 0;
}

// test hook: one raw DEFLATE stream through the table-driven decoder (mode 1) or zlib (mode 0); returns bytes written or -1
long nptb_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, int mode) {
  if (mode == 1) return nptinf::inflate_raw(in, in_len, out, out_cap);
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) return -1;
  zs.next_in = const_cast<Bytef*>(in); zs.avail_in = (uInt)in_len; zs.next_out = out; zs.avail_out = (uInt)out_cap;
  const int rc = inflate(&zs, Z_FINISH);
  const long n = (long)(out_cap - zs.avail_out);
  inflateEnd(&zs);
  return rc == Z_STREAM_END ? n : -1;
}

}  // extern "C"
