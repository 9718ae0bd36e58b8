// synth.cpp — synthetic direct-RNA workload generator (host side, multi-threaded, deterministic).
//
// Produces the structure-of-arrays batches of include/npt.h for the benchmark configurations of
// BASELINE.md §4 / SURVEY.md §8(d): a mixture of canonical sgRNA reads (leader→TRS-B junction),
// 5'-truncated leaderless reads, genomic fragments and double-junction reads, with per-base
// mismatch/insertion/deletion noise, soft clips and M-op CIGARs (minimap2 default, no --eqx).
// The reference ships no sample BAM, so every workload in this repo comes from here.
//
// Read i of a stream depends only on (seed, i): the output is identical for any thread count and any
// batching, so the oracle and the device path always see the same reads.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64 seeding + xoshiro256**
  uint64_t s[4];
  static uint64_t sm(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t idx) {
    uint64_t x = seed * 0xD1342543DE82EF95ull + idx * 0x2545F4914F6CDD1Dull + 0x1234567;
    for (int i = 0; i < 4; i++) s[i] = sm(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }  // inclusive
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

}  // namespace

extern "C" {

#define NPTS_MAX_ACCEPTORS 32

typedef struct npts_params {
  int32_t genome_len;
  int32_t polya_len;          // A-tail of the genome (33 for VIC01)
  int32_t donor;              // leader donor site (69 for SARS-CoV-2)
  int32_t n_acceptors;
  int32_t acceptor[NPTS_MAX_ACCEPTORS];
  double acceptor_w[NPTS_MAX_ACCEPTORS];  // weights (normalised inside)
  double w_canonical, w_leaderless, w_genomic, w_double;
  double p_mismatch, p_ins, p_del;
  double indel_mean;          // geometric mean length
  double p_neg;               // fraction with flag 0x10
  int32_t num_sources;        // src = read index % num_sources unless src_fixed >= 0
  int32_t src_fixed;
  int32_t use_eqx;            // emit =/X instead of M (exercises the =/X quirk path)
  double leaderless_len;      // lognormal median of the leaderless (3'-anchored) read length
  double genomic_len;         // lognormal median of genomic fragment length
} npts_params;

// genome: uniform ACGT from the seed + polyA tail; BAM 4-bit codes A=1 C=2 G=4 T=8
void npts_genome(uint64_t seed, int32_t len, int32_t polya, uint8_t* codes) {
  Rng r(seed, 0xC0FFEEull);
  static const uint8_t c4[4] = {1, 2, 4, 8};
  for (int i = 0; i < len; i++) codes[i] = c4[r.next() & 3];
  for (int i = std::max(0, len - polya); i < len; i++) codes[i] = 1;
}

struct npts_out {
  std::vector<int32_t> pos;
  std::vector<uint16_t> flag, src;
  std::vector<uint32_t> cigar_off, cigar;
  std::vector<uint64_t> seq_off;
  std::vector<uint8_t> seq4;
};

struct Noise { double p_mismatch, p_ins, p_del, indel_mean; bool eqx, clips; };
static void emit_blocks(const Noise& P, const uint8_t* genome, Rng& r, const int* bs, const int* be, int nb, std::vector<uint32_t>& cig,
                        std::vector<uint8_t>& bases);

static void gen_read(const npts_params& P, const uint8_t* genome, uint64_t seed, uint64_t idx, int32_t& pos, uint16_t& flag, uint16_t& src,
                     std::vector<uint32_t>& cig, std::vector<uint8_t>& bases) {
  Rng r(seed, idx);
  const int G = P.genome_len;
  cig.clear(); bases.clear();
  // exon blocks, 1-based inclusive
  int bs[3], be[3], nb = 0;
  double u = r.uni();
  const double wsum = P.w_canonical + P.w_leaderless + P.w_genomic + P.w_double;
  u *= wsum;
  auto pick_acceptor = [&]() {
    double ws = 0; for (int i = 0; i < P.n_acceptors; i++) ws += P.acceptor_w[i];
    double v = r.uni() * ws; int k = 0;
    for (; k < P.n_acceptors - 1; k++) { v -= P.acceptor_w[k]; if (v < 0) break; }
    return P.acceptor[k] + r.range(-3, 3);
  };
  int end3 = G - P.polya_len - r.range(0, 40);
  if (u < P.w_canonical) {
    int st = r.range(1, 30), don = P.donor + r.range(-3, 3), acc = pick_acceptor();
    bs[0] = st; be[0] = don; bs[1] = acc; be[1] = end3; nb = 2;
  } else if (u < P.w_canonical + P.w_leaderless) {
    int len = (int)std::exp(std::log(P.leaderless_len) + 0.6 * r.normal());  // 3'-anchored, 5'-truncated
    len = std::min(std::max(200, len), G / 2);
    bs[0] = std::max(1, end3 - len); be[0] = end3; nb = 1;
  } else if (u < P.w_canonical + P.w_leaderless + P.w_genomic) {
    int st = r.range(1, std::min(20000, G * 2 / 3));
    int len = (int)std::exp(std::log(P.genomic_len) + 0.5 * r.normal());
    len = std::max(50, len);
    bs[0] = st; be[0] = std::min(G - 1, st + len); nb = 1;
  } else {
    int st = r.range(1, 30), don = P.donor + r.range(-3, 3);
    int a1 = r.range(2000, G / 2), d2 = a1 + r.range(200, 1500);
    int a2 = std::max(d2 + 1200, pick_acceptor());
    if (a2 + 100 > end3) a2 = end3 - 100;
    if (d2 + 1 >= a2) d2 = a2 - 1100;
    bs[0] = st; be[0] = don; bs[1] = a1; be[1] = d2; bs[2] = a2; be[2] = end3; nb = 3;
  }
  for (int b = 0; b < nb; b++) { if (be[b] < bs[b]) be[b] = bs[b]; if (be[b] > G) be[b] = G; }
  pos = bs[0];
  flag = (uint16_t)(r.uni() < P.p_neg ? 16 : 0);
  src = (uint16_t)(P.src_fixed >= 0 ? P.src_fixed : (int)(idx % (uint64_t)std::max(1, P.num_sources)));
  Noise N{P.p_mismatch, P.p_ins, P.p_del, P.indel_mean, P.use_eqx != 0, true};
  emit_blocks(N, genome, r, bs, be, nb, cig, bases);
}

// aligned blocks -> CIGAR + bases with per-base noise and soft clips
static void emit_blocks(const Noise& P, const uint8_t* genome, Rng& r, const int* bs, const int* be, int nb, std::vector<uint32_t>& cig,
                        std::vector<uint8_t>& bases) {
  static const uint8_t c4[4] = {1, 2, 4, 8};
  auto push = [&](int len, int op) {
    if (len <= 0) return;
    if (!cig.empty() && (int)(cig.back() & 0xF) == op) cig.back() += (uint32_t)len << 4;
    else cig.push_back(((uint32_t)len << 4) | (uint32_t)op);
  };
  auto geo = [&]() { int l = 1; double q = 1.0 - 1.0 / P.indel_mean; while (r.uni() < q && l < 30) l++; return l; };
  int clip5 = P.clips ? r.range(0, 30) : 0, clip3 = P.clips ? r.range(10, 80) : 0;
  if (clip5) { push(clip5, 4); for (int i = 0; i < clip5; i++) bases.push_back(c4[r.next() & 3]); }
  for (int b = 0; b < nb; b++) {
    if (b > 0) push(bs[b] - be[b - 1] - 1, 3);
    int p = bs[b];
    while (p <= be[b]) {
      bool edge = (p == bs[b] || p == be[b]);
      double v = edge ? 1.0 : r.uni();
      if (v < P.p_ins) {
        int l = geo(); push(l, 1);
        for (int i = 0; i < l; i++) bases.push_back(c4[r.next() & 3]);
      } else if (v < P.p_ins + P.p_del) {
        int l = std::min(geo(), be[b] - p);  // keep the block's last base aligned
        if (l > 0) { push(l, 2); p += l; }
      } else {
        uint8_t g = genome[p - 1];
        bool mm = !edge && (r.uni() < P.p_mismatch);
        uint8_t c = g;
        if (mm) { do { c = c4[r.next() & 3]; } while (c == g); }
        if (P.eqx) push(1, mm ? 8 : 7); else push(1, 0);
        bases.push_back(c); p++;
      }
    }
  }
  push(clip3, 4);
  for (int i = 0; i < clip3; i++) bases.push_back(c4[r.next() & 3]);
}

}  // extern "C"

template <typename Gen>
static void* generate_with(Gen gen, int64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  npts_out* O = new npts_out();
  O->pos.resize(n); O->flag.resize(n); O->src.resize(n);
  O->cigar_off.assign(n + 1, 0); O->seq_off.assign(n + 1, 0);
  const int64_t CH = 4096;
  const int64_t nch = (n + CH - 1) / CH;
  std::vector<std::vector<uint32_t>> ccig(nch);
  std::vector<std::vector<uint8_t>> cseq(nch);
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    std::vector<uint32_t> cig; std::vector<uint8_t> bases;
    for (;;) {
      int64_t c = next.fetch_add(1);
      if (c >= nch) break;
      auto& vc = ccig[c]; auto& vs = cseq[c];
      for (int64_t i = c * CH; i < std::min(n, (c + 1) * CH); i++) {
        gen(i, O->pos[i], O->flag[i], O->src[i], cig, bases);
        O->cigar_off[i + 1] = (uint32_t)cig.size();          // lengths for now
        O->seq_off[i + 1] = (uint64_t)((bases.size() + 1) / 2);
        vc.insert(vc.end(), cig.begin(), cig.end());
        size_t nb = bases.size();
        for (size_t k = 0; k < nb; k += 2) vs.push_back((uint8_t)((bases[k] << 4) | (k + 1 < nb ? bases[k + 1] : 0)));
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(work);
  for (auto& t : th) t.join();
  for (int64_t i = 0; i < n; i++) { O->cigar_off[i + 1] += O->cigar_off[i]; O->seq_off[i + 1] += O->seq_off[i]; }
  O->cigar.resize(O->cigar_off[n]); O->seq4.resize(O->seq_off[n]);
  for (int64_t c = 0; c < nch; c++) {
    int64_t i0 = c * CH;
    if (!ccig[c].empty()) memcpy(O->cigar.data() + O->cigar_off[i0], ccig[c].data(), ccig[c].size() * 4);
    if (!cseq[c].empty()) memcpy(O->seq4.data() + O->seq_off[i0], cseq[c].data(), cseq[c].size());
  }
  return O;
}

extern "C" {

void* npts_generate(const npts_params* P, const uint8_t* genome, uint64_t seed, uint64_t first_index, int64_t n, int nthreads) {
  return generate_with([=](int64_t i, int32_t& pos, uint16_t& flag, uint16_t& src, std::vector<uint32_t>& cig, std::vector<uint8_t>& bases) {
    gen_read(*P, genome, seed, first_index + (uint64_t)i, pos, flag, src, cig, bases);
  }, n, nthreads);
}

// ---- host mode (BASELINE.json configs[4]): spliced long reads on one chromosome of a gene model.
// read = random gene (popularity skewed), its exons with internal ones skipped at random, 5'-truncated; source 0 (with
// transcriptome != 0) carries the full transcripts aligned error-free, the "transcriptome BAM" of --firstIsTranscriptome.
typedef struct npts_host_params {
  int32_t chrom_len;
  int32_t n_genes;
  const int32_t* gene_exon_off;   // n_genes + 1
  const int32_t* exon_start;      // 1-based inclusive, ascending inside a gene
  const int32_t* exon_end;
  double p_skip, p_trunc5, p_intergenic, gene_skew;
  double p_mismatch, p_ins, p_del, indel_mean, p_neg;
  int32_t num_sources, src_fixed, transcriptome;
} npts_host_params;

#define NPTS_MAX_BLOCKS 64
void* npts_generate_host(const npts_host_params* P, const uint8_t* genome, uint64_t seed, uint64_t first_index, int64_t n, int nthreads) {
  const npts_host_params H = *P;
  return generate_with([=](int64_t i, int32_t& pos, uint16_t& flag, uint16_t& src, std::vector<uint32_t>& cig, std::vector<uint8_t>& bases) {
    const uint64_t idx = first_index + (uint64_t)i;
    Rng r(seed, idx);
    cig.clear(); bases.clear();
    src = (uint16_t)(H.src_fixed >= 0 ? H.src_fixed : (int)(idx % (uint64_t)std::max(1, H.num_sources)));
    const bool exact = H.transcriptome && src == 0;
    int bs[NPTS_MAX_BLOCKS], be[NPTS_MAX_BLOCKS], nb = 0;
    if (H.n_genes <= 0 || (!exact && r.uni() < H.p_intergenic)) {   // unspliced fragment anywhere
      int len = std::max(100, (int)std::exp(std::log(900.0) + 0.5 * r.normal()));
      len = std::min(len, std::max(1, H.chrom_len / 2));
      int st = r.range(1, std::max(1, H.chrom_len - len - 1));
      bs[0] = st; be[0] = std::min(H.chrom_len, st + len); nb = 1;
    } else {
      const int g = exact ? (int)(idx / (uint64_t)std::max(1, H.num_sources) % (uint64_t)H.n_genes)
                          : std::min(H.n_genes - 1, (int)(H.n_genes * std::pow(r.uni(), H.gene_skew)));
      const int e0 = H.gene_exon_off[g], e1 = H.gene_exon_off[g + 1];
      int first = e0;
      if (!exact && e1 - e0 > 1 && r.uni() < H.p_trunc5) first = r.range(e0, e1 - 1);   // direct RNA: 3'-anchored
      for (int e = first; e < e1 && nb < NPTS_MAX_BLOCKS; e++) {
        const bool inner = e > first && e < e1 - 1;
        if (!exact && inner && r.uni() < H.p_skip) continue;
        bs[nb] = H.exon_start[e]; be[nb] = H.exon_end[e]; nb++;
      }
      if (!exact && nb > 0 && r.uni() < 0.7) {   // the read starts somewhere inside its first exon
        const int L = be[0] - bs[0];
        if (L > 20) bs[0] += r.range(0, L - 20);
      }
    }
    for (int b = 0; b < nb; b++) { if (be[b] < bs[b]) be[b] = bs[b]; if (be[b] > H.chrom_len) be[b] = H.chrom_len; }
    pos = bs[0];
    flag = (uint16_t)(r.uni() < H.p_neg ? 16 : 0);
    Noise N{exact ? 0.0 : H.p_mismatch, exact ? 0.0 : H.p_ins, exact ? 0.0 : H.p_del, H.indel_mean, false, !exact};
    emit_blocks(N, genome, r, bs, be, nb, cig, bases);
  }, n, nthreads);
}

void npts_sizes(void* h, int64_t* n, int64_t* n_cigar, int64_t* n_seq_bytes) {
  npts_out* O = (npts_out*)h;
  *n = (int64_t)O->pos.size(); *n_cigar = (int64_t)O->cigar.size(); *n_seq_bytes = (int64_t)O->seq4.size();
}

void npts_copy(void* h, int32_t* pos, uint16_t* flag, uint16_t* src, uint32_t* cigar_off, uint32_t* cigar, uint64_t* seq_off, uint8_t* seq4) {
  npts_out* O = (npts_out*)h;
  size_t n = O->pos.size();
  memcpy(pos, O->pos.data(), n * 4); memcpy(flag, O->flag.data(), n * 2); memcpy(src, O->src.data(), n * 2);
  memcpy(cigar_off, O->cigar_off.data(), (n + 1) * 4); memcpy(cigar, O->cigar.data(), O->cigar.size() * 4);
  memcpy(seq_off, O->seq_off.data(), (n + 1) * 8); memcpy(seq4, O->seq4.data(), O->seq4.size());
}

void npts_free(void* h) { delete (npts_out*)h; }

}  // extern "C"
