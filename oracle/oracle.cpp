// oracle.cpp — CPU restatement of npTranscript's read→transcript-cluster path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this.  The product (nptranscript_b200/) never does.
//
// PARITY UNPINNED: the reference (Java 8 + htsjdk + japsa) cannot be built or run here (no JVM,
// two vendored jars missing) and ships no tests, fixtures or golden vectors for this path, so this
// file is a line-by-line restatement derived by reading.  J = /root/reference/java/src/main/java/npTranscript.
// Third-party semantics restated from their published behaviour:
//   htsjdk 3.0.1 (java/pom.xml:39-43): SAMRecord.getAlignmentEnd = start + sum(len of M,D,N,=,X) - 1;
//     getAlignmentBlocks (one block per M/=/X element, readStart 1-based counting S and I, not H);
//     getReadPositionAtReferencePosition(pos) (first block whose end >= pos; 0 if pos precedes it or none).
//   japsa 1.9.4e (java/pom.xml:54-58): Sequence.getBase equality == same IUPAC letter (case-insensitive);
//     here both sides are BAM 4-bit codes.
//
// The restatement keeps the Java data structures' meaning (string cluster keys, list-of-Integer
// sub-cluster keys, sequential arrival order, doubles for Count.true_breaks) but not their cost.

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/npt.h"

namespace {

// ---------------------------------------------------------------- Annotation (J/cluster/Annotation.java)
struct Annot {
  int kind = NPT_ANNOT_CSV;
  std::vector<int> start, end, name_id;
  std::vector<uint8_t> strand;
  std::vector<uint32_t> name_hash;  // GFF: HashMap.hash(String.hashCode()) of the row's gene string
  int seqlen = 0;
  int tolerance = 10;  // Annotation.tolerance = --bin (ViralTranscriptAnalysisCmd2.java:497)
  bool enforceStrand = false;
  int round = 100;     // CigarHash2.round
  std::vector<std::vector<int>> spliced, unspliced;  // [src][orf]  (Annotation.java:186-187)

  // TranscriptUtils.round (J/cluster/TranscriptUtils.java:20-23): floor of a double division
  static int jround(int pos, int round) { return (int)std::floor((double)pos / (double)round); }

  // Annotation.nextDownstream :62-72 / EmptyAnnotation :18-20.  Returns a token (npt.h NPT_TOK_*).
  int nextDownstream(int rightBreak, bool forward) const {
    if (kind == NPT_ANNOT_EMPTY) return jround(rightBreak, round);
    for (size_t i = 0; i < start.size(); i++) {
      if (!enforceStrand || forward == (bool)strand[i]) {
        if (rightBreak - tolerance <= start[i]) return name_id[i];
      }
    }
    return NPT_TOK_END;
  }
  // Annotation.nextUpstream :87-97 / EmptyAnnotation :22-24
  int nextUpstream(int leftBreak, bool forward) const {
    if (kind == NPT_ANNOT_EMPTY) return jround(leftBreak, round);
    for (int i = (int)start.size() - 1; i >= 0; i--) {
      if (!enforceStrand || forward == (bool)strand[i]) {
        if (leftBreak + tolerance >= start[i]) return name_id[i];
      }
    }
    return NPT_TOK_ST;
  }
  // GFFAnnotation.contains :72-76
  bool gffContains(int l2, int r2, int i) const {
    return r2 - tolerance <= end[i] && r2 + tolerance >= start[i] && l2 - tolerance <= end[i] && l2 + tolerance >= start[i];
  }
  // GFFAnnotation.nextUpstream(l1, r1, chrom_index, forward) :80-100.  Returns the exon row or -1 (null).
  int gffUpstream(int l1, int r1, bool forward_known, bool forward) const {
    if (l1 < 0) return -1;
    std::map<int, int> indices;  // TreeMap<score, row>: a later put (lower row) replaces an equal score
    for (int i = (int)start.size() - 1; i >= 0; i--) {
      if (!forward_known || forward == (bool)strand[i]) {
        if (gffContains(l1, r1, i)) indices[std::abs(l1 - start[i]) + std::abs(r1 - end[i])] = i;
      }
    }
    if (!indices.empty()) {
      if (indices.begin()->first <= 2 * tolerance) return indices.begin()->second;
    }
    return -1;
  }
  // Annotation.isLeader :214-216 / EmptyAnnotation :28-30
  bool isLeader(int prev_pos) const {
    if (kind == NPT_ANNOT_EMPTY) return prev_pos < 100;
    return prev_pos < start.at(1) - tolerance;
  }
  // Annotation.getTypeInd :233-236
  int getTypeInd(int st, int en, int startThresh, int endThresh) const {
    if (st <= startThresh) return en >= seqlen - endThresh ? 0 : 1;
    return en >= seqlen - endThresh ? 2 : 3;
  }
};

// ---------------------------------------------------------------- per-read scratch (the "coRefPositions" CigarCluster)
struct Scratch {
  std::vector<int> breaks;  // CigarHash2 breaks
  int prev_position = -1;
  // depth events of the read, in emission order (map/maps[src]/errors[src] of the scratch cluster)
  std::vector<int> ev_pos;
  std::vector<uint8_t> ev_match;
  std::vector<int> ms_pos, me_pos;  // mapStart / mapEnd single-count events
  void clear() {
    breaks.clear(); prev_position = -1; ev_pos.clear(); ev_match.clear(); ms_pos.clear(); me_pos.clear();
  }
};

struct Cfg {
  npt_config c;
  int min_first_last_exon_length;  // 0 virus / 10 host (ViralTranscriptAnalysisCmd2.java:339,360)
  bool recordStartEnd;
};

// CigarCluster.add  (J/cluster/CigarCluster.java:470-498)
inline void cc_add(Scratch& s, const Cfg& cfg, int pos, bool match) {
  bool break_p = s.prev_position < 0 || pos - s.prev_position > cfg.c.break_thresh;
  if (cfg.c.record_depth) {
    s.ev_pos.push_back(pos);
    s.ev_match.push_back(match ? 1 : 0);
    if (cfg.recordStartEnd && break_p && s.prev_position > 0) {
      s.ms_pos.push_back(pos);
      s.me_pos.push_back(s.prev_position);
    }
  }
  if (match) {
    if (break_p) {
      if (s.prev_position > 0) s.breaks.push_back(s.prev_position);
      s.breaks.push_back(pos);
    }
    s.prev_position = pos;
  }
}

inline int seq4_at(const uint8_t* seq4, int64_t i) {
  uint8_t b = seq4[i >> 1];
  return (i & 1) ? (b & 0xF) : (b >> 4);
}

// java.util.HashSet<String> as far as iteration order goes (OpenJDK HashMap.putVal / resize / treeifyBin): table of 16
// buckets, doubled when size exceeds 0.75*capacity or when a bucket chain already holds 8 nodes at an append (tables
// below 64 buckets are resized instead of treeified); a split keeps the relative order inside each half, so a bucket lists
// its members in insertion order.  Strings are represented by (name id, spread hash).  Tree bins (only possible from 64
// buckets on) are not restated: ok=false.
struct JavaStringSet {
  std::vector<std::vector<std::pair<int, uint32_t>>> tab;
  size_t size = 0; bool ok = true;
  JavaStringSet() : tab(16) {}
  void resize() {
    std::vector<std::vector<std::pair<int, uint32_t>>> nt(tab.size() * 2);
    for (auto& bkt : tab) for (auto& e : bkt) nt[e.second & (nt.size() - 1)].push_back(e);
    tab.swap(nt);
  }
  void add(int id, uint32_t h) {
    auto& bkt = tab[h & (tab.size() - 1)];
    for (auto& e : bkt) if (e.first == id) return;
    const size_t before = bkt.size();
    bkt.emplace_back(id, h);
    if (before >= 8) { if (tab.size() < 64) resize(); else ok = false; }
    if (++size > tab.size() * 3 / 4) resize();
  }
  template <class F> void for_each(F f) const { for (auto& bkt : tab) for (auto& e : bkt) f(e.first); }
};

struct ReadOut {
  bool ok = false;            // false: bad cigar (reference throws) → not clustered
  int startPos = 0, endPos = 0, readSt = 0, readEn = 0;
  int type_ind = 0;
  bool leader = false, neg = false, forward_known = false, forward = false;
  int maps_sum = 0, err_sum = 0;
  std::vector<int> key;       // key tokens (rendered to the Java string by render_key)
  std::vector<int> bp;        // break-point events: level, st, en triples (IdentityProfile1.java:276-286)
  std::vector<int> orf_spliced, orf_unspliced;  // ORF indices to count (Annotation.addCount)
};

// htsjdk SAMRecord.getReadPositionAtReferencePosition(pos) over the alignment blocks of a cigar.
int read_pos_at_ref_pos(const uint32_t* cig, int nc, int alnStart, int pos) {
  if (pos <= 0) return 0;
  int readBase = 1, refBase = alnStart;
  for (int k = 0; k < nc; k++) {
    int len = (int)(cig[k] >> 4), op = (int)(cig[k] & 0xF);
    switch (op) {
      case 5: break;                       // H
      case 6: break;                       // P
      case 4: readBase += len; break;      // S
      case 3: refBase += len; break;       // N
      case 2: refBase += len; break;       // D
      case 1: readBase += len; break;      // I
      case 0: case 7: case 8: {            // M = X : an AlignmentBlock(readBase, refBase, len)
        int bend = refBase + len - 1;      // CoordMath.getEnd
        if (bend >= pos) {
          if (pos < refBase) return 0;
          return pos - refBase + readBase;
        }
        readBase += len; refBase += len;
        break;
      }
      default: break;
    }
  }
  return 0;
}

// IdentityProfile1.processCigar (J/cluster/IdentityProfile1.java:482-585)
bool process_cigar(Scratch& s, const Cfg& cfg, int alnStart, const uint32_t* cig, int nc, const uint8_t* seq4, int64_t nbases,
                   const uint8_t* ref, int64_t refLen, int* alnEndOut) {
  int64_t readPos = 0;
  int64_t refPos = (int64_t)alnStart - 1;
  // CigarCluster.addStart :442-446
  s.breaks.push_back(alnStart);
  s.prev_position = alnStart;
  int64_t refSpan = 0;
  for (int k = 0; k < nc; k++) {
    const int length = (int)(cig[k] >> 4);
    const int op = (int)(cig[k] & 0xF);
    switch (op) {
      case 5: break;  // H
      case 6: break;  // P
      case 4: readPos += length; break;  // S
      case 3: refPos += length; refSpan += length; break;  // N
      case 2:  // D: advance first, then emit (quirk, :513-522)
        refPos += length; refSpan += length;
        for (int i = 0; i < length && refPos + i < refLen; i++) cc_add(s, cfg, (int)(refPos + i + 1), false);
        break;
      case 1: readPos += length; break;  // I
      case 0:  // M :531-550
        // readSeq.getBase(readPos+i) past the end of the read throws in the reference: record not clustered.  nbases counts
        // the pad nibble of an odd-length read as a base (npt_batch carries bytes, not l_seq).
        if (length > 0 && refPos < refLen && readPos + std::min<int64_t>(length, refLen - refPos) > nbases) return false;
        for (int i = 0; i < length && refPos + i < refLen; i++) {
          bool m = ref[refPos + i] == seq4_at(seq4, readPos + i);
          cc_add(s, cfg, (int)(refPos + i + 1), m);
        }
        readPos += length; refPos += length; refSpan += length;
        break;
      case 7:  // EQ: advance first, then emit as matches (quirk, :552-562)
        readPos += length; refPos += length; refSpan += length;
        for (int i = 0; i < length && refPos + i < refLen; i++) cc_add(s, cfg, (int)(refPos + i + 1), true);
        break;
      case 8:  // X: advance first, then emit as mismatches (quirk, :564-576)
        readPos += length; refPos += length; refSpan += length;
        for (int i = 0; i < length && refPos + i < refLen; i++) cc_add(s, cfg, (int)(refPos + i + 1), false);
        break;
      default:
        return false;  // IllegalStateException :578
    }
  }
  int alnEnd = (int)(alnStart + refSpan - 1);  // htsjdk getAlignmentEnd
  s.breaks.push_back(alnEnd);                  // CigarCluster.addEnd :466-469
  *alnEndOut = alnEnd;
  return true;  // breaks.size() is even by construction (:582-584 never fires)
}

// IdentityProfile1.identity1 :594-643 + processRefPositions :149-356 for one read (the order-free part).
void per_read(const Cfg& cfg, const Annot& annot, const uint8_t* ref, int64_t refLen, int alnStart, int flag, int src,
              const uint32_t* cig, int nc, const uint8_t* seq4, int64_t nbases, Scratch& s, ReadOut& o) {
  s.clear();
  o = ReadOut();
  int alnEnd = 0;
  if (alnStart <= 0) return;  // unmapped-style record: the caller's filter chain drops these (ViralTranscriptAnalysisCmd2.java:716)
  if (!process_cigar(s, cfg, alnStart, cig, nc, seq4, nbases, ref, refLen, &alnEnd)) return;
  o.ok = true;
  // :634-635
  int st_r = read_pos_at_ref_pos(cig, nc, alnStart, alnStart);
  int end_r = read_pos_at_ref_pos(cig, nc, alnStart, alnEnd);
  o.neg = (flag & 0x10) != 0;
  std::vector<int>& breaks = s.breaks;
  // processRefPositions :159-170
  o.startPos = breaks[0];
  o.endPos = breaks[breaks.size() - 1];
  o.readSt = st_r; o.readEn = end_r;
  if (cfg.c.rna[src]) { o.forward_known = true; o.forward = !o.neg; } else { o.forward_known = false; }
  // :205-228 small first/last exon removal
  if (cfg.min_first_last_exon_length > 0 && breaks.size() > 2) {
    int sze = (int)breaks.size();
    int diff0 = breaks[1] - o.startPos;
    int diff1 = o.endPos - breaks[sze - 2];
    const int m = cfg.min_first_last_exon_length;
    if (sze == 4 && diff0 < m && diff1 < m) {
      if (diff1 < diff0) { breaks.erase(breaks.begin() + sze - 1); breaks.erase(breaks.begin() + sze - 2); }
      else { breaks.erase(breaks.begin() + 1); breaks.erase(breaks.begin()); }
    } else {
      if (diff1 < m) { breaks.erase(breaks.begin() + sze - 1); breaks.erase(breaks.begin() + sze - 2); }
      if (diff0 < m) { breaks.erase(breaks.begin() + 1); breaks.erase(breaks.begin()); }
    }
    o.endPos = breaks[breaks.size() - 1];
    o.startPos = breaks[0];
  }
  const int nB = (int)breaks.size();
  // :256
  o.leader = cfg.c.coronavirus ? (nB > 1 && annot.isLeader(breaks[1])) : false;
  const bool fwd = o.forward;  // auto-unboxed Boolean; callers guarantee forward_known in coronavirus mode
  if (cfg.c.coronavirus) {  // :266-292
    if (cfg.c.include_start || nB == 2) o.key.push_back(annot.nextUpstream(o.startPos, fwd));
    else o.key.push_back(NPT_TOK_NOENDS);
    bool firstBreak = true;
    for (int i = 1; i < nB - 1; i += 2) {
      int gap = breaks[i + 1] - breaks[i];
      o.key.push_back(annot.nextUpstream(breaks[i], fwd));
      o.key.push_back(annot.nextDownstream(breaks[i + 1], fwd));
      if (gap > cfg.c.break_thresh) {
        if (firstBreak) {
          firstBreak = false;
          o.bp.push_back(0); o.bp.push_back(breaks[i]); o.bp.push_back(breaks[i + 1]);
          o.leader = true;
        } else {
          o.bp.push_back(1); o.bp.push_back(breaks[i]); o.bp.push_back(breaks[i + 1]);
        }
      }
    }
    if (cfg.c.include_start || nB == 2) o.key.push_back(annot.nextUpstream(breaks[nB - 1], fwd));
  } else if (annot.kind == NPT_ANNOT_GFF) {  // :300-323
    JavaStringSet genes;
    std::vector<int> coords;  // GFFAnnotation.getCoord(pos, chrom_index, forward) :132-135 -> the bin; chrom prefix at render
    const bool fwdOrNull = !o.forward_known || o.forward;
    for (int i = 0; i + 1 < nB; i += 2) {
      int upst = annot.gffUpstream(breaks[i], breaks[i + 1], o.forward_known, o.forward);
      if (upst >= 0) genes.add(annot.name_id[upst], annot.name_hash[upst]);
      else coords.push_back(Annot::jround(fwdOrNull ? breaks[i + 1] : breaks[i], annot.round));
    }
    if (!genes.ok) { o.ok = false; o.startPos = o.endPos = 0; return; }  // tree bins: not restated, record dropped
    if (genes.size > 0) {
      o.key.push_back(NPT_TOK_GENESET);
      genes.for_each([&](int id) { o.key.push_back(id); });
    } else {
      o.key.push_back(fwdOrNull ? coords.back() : coords.front());
    }
  } else {  // :293-299 (EmptyAnnotation)
    if (o.forward_known) {
      if (o.forward) o.key.push_back(annot.nextUpstream(breaks[nB - 1], fwd));
      else o.key.push_back(annot.nextUpstream(breaks[0], fwd));
    } else {
      o.key.push_back(annot.nextUpstream(breaks[nB - 1], fwd));
    }
  }
  if (cfg.c.enforce_strand) o.key.push_back(fwd ? NPT_TOK_PLUS : NPT_TOK_MINUS);  // :325-327
  o.type_ind = cfg.c.coronavirus ? annot.getTypeInd(o.startPos, o.endPos, cfg.c.start_thresh, cfg.c.end_thresh) : 0;  // :330
  // :332 setStartEnd (CigarCluster.java:360-365)
  if (cfg.recordStartEnd) { s.ms_pos.push_back(o.startPos); s.me_pos.push_back(o.endPos); }
  // :349-356 ORF counts
  if (cfg.c.coronavirus) {
    int st1 = breaks[nB - 2];
    for (int i = (int)annot.start.size() - 1; i >= 0; i--) {
      if (annot.start[i] < st1 - 10) break;
      if (o.endPos > annot.start[i] + 100) {
        if (o.startPos <= cfg.c.start_thresh) o.orf_spliced.push_back(i); else o.orf_unspliced.push_back(i);
      }
    }
  }
  // CigarCluster.getError :538-544 inputs
  if (cfg.c.record_depth) {
    o.maps_sum = (int)s.ev_pos.size();
    int e = 0; for (uint8_t m : s.ev_match) e += !m;
    o.err_sum = e;
  }
}

// ---------------------------------------------------------------- cluster side (CigarCluster as accumulator)
struct Count {  // CigarCluster.Count :63-135
  std::vector<int> count;
  std::vector<double> true_breaks;  // in 1kb
  std::vector<long long> isum;      // exact integer sums (not in the reference; used to check the device's int64 sums)
  bool has_tb = false;
  int break_sum = 0;
  int k = 0;                        // creation index inside the cluster (the N of ".tN")
  long long first_read = 0;
  bool forward = false, neg = false;
  void addBreaks(const std::vector<int>& breaks, int src, bool firstIsTranscriptome) {  // :84-96
    if (!firstIsTranscriptome || src == 0 || count[0] == 0) {
      break_sum += 1;
      if (!has_tb) { true_breaks.assign(breaks.size(), 0.0); isum.assign(breaks.size(), 0); has_tb = true; }
      if (true_breaks.size() != breaks.size()) {
        // "warning breaks have different lengths, so not updating"
      } else {
        for (size_t i = 0; i < true_breaks.size(); i++) { true_breaks[i] += breaks[i] / 1e3; isum[i] += breaks[i]; }
      }
    }
  }
};

struct Cluster {
  int idx = 0;                      // n of "ID<chr>.<n>"
  std::vector<int> key;
  std::vector<int> breaks;          // first read's cloneBreaks()
  bool forward_known = false, forward = false, neg = false;
  std::vector<int> readCount;
  int readCountSum = 0;
  long long first_read = 0;
  std::map<std::vector<int>, int> sub_index;  // HashMap<CigarHash2,Count>: list equality
  std::vector<Count> subs;                    // in creation order
  std::vector<std::vector<int>> sub_keys;
  // coverage: dense [src][seqlen+2]
  std::vector<int> depth, errors, mapStart, mapEnd;
};

// CigarCluster.cloneBreaks :620-631 with CigarHash2.addAllR (:61-68) / TranscriptUtils.round
std::vector<int> clone_breaks(const std::vector<int>& breaks, bool includeStart, bool forward_known, bool forward, int round) {
  size_t a = 0, b = breaks.size();
  if (!includeStart && forward_known) { if (forward) a = 1; else b = breaks.size() - 1; }
  std::vector<int> r;
  for (size_t i = a; i < b; i++) r.push_back(Annot::jround(breaks[i], round));
  return r;
}

struct Run {
  Cfg cfg;
  Annot annot;
  std::vector<uint8_t> ref;
  int chrom_index = 0;
  int stride = 0;
  bool calcBreaks1 = false;
  std::unordered_map<std::string, int> l;  // Map<CigarHash,CigarCluster>: String equality (CigarHash.java:42-62)
  std::vector<Cluster> clusters;
  std::vector<int> totalCounts;
  std::map<std::array<int, 4>, int> pairs;  // (src, level, st, en) -> count   BreakPoints.java:102-108
  long long n_reads = 0;
  // per read outputs
  std::vector<int> r_cluster, r_sub, r_start, r_end, r_rst, r_ren, r_maps, r_err, r_breaks;
  std::vector<uint32_t> r_flags;
  std::vector<uint64_t> r_boff;
  bool keep_reads = true;
};

// the Java string of a key, for the map (delimiters :113-114, grammar :266-292)
std::string render_key(const std::vector<int>& tok, const Cfg& cfg, const Annot& annot, int chrom_index) {
  std::string s;
  auto name = [&](int t) -> std::string {
    if (annot.kind == NPT_ANNOT_EMPTY) return std::to_string(chrom_index) + "." + std::to_string(t);
    if (t == NPT_TOK_ST) return std::string("st") + (chrom_index > 0 ? "." + std::to_string(chrom_index) : "");
    if (t == NPT_TOK_END) return std::string("end") + (chrom_index > 0 ? "." + std::to_string(chrom_index) : "");
    return "#" + std::to_string(t);  // gene strings are not known here; name ids are unique per string
  };
  int n = (int)tok.size();
  bool strandTok = n > 0 && (tok[n - 1] == NPT_TOK_PLUS || tok[n - 1] == NPT_TOK_MINUS);
  int m = strandTok ? n - 1 : n;
  if (annot.kind == NPT_ANNOT_GFF) {
    if (m > 0 && tok[0] == NPT_TOK_GENESET) { for (int i = 1; i < m; i++) { s += name(tok[i]); s += '_'; } }
    else s += std::to_string(chrom_index) + "." + std::to_string(tok[0]);
  } else if (!cfg.c.coronavirus) {
    for (int i = 0; i < m; i++) s += name(tok[i]);
  } else if (m > 0 && tok[0] == NPT_TOK_NOENDS) {  // (up,down)* only
    for (int i = 1; i + 1 < m; i += 2) { s += name(tok[i]); s += ','; s += name(tok[i + 1]); s += '_'; }
  } else {  // up(start) '_' { up ',' down '_' }* up(end)
    s += name(tok[0]); s += '_';
    for (int i = 1; i + 1 < m - 1; i += 2) { s += name(tok[i]); s += ','; s += name(tok[i + 1]); s += '_'; }
    s += name(tok[m - 1]);
  }
  if (strandTok) s += (tok[n - 1] == NPT_TOK_PLUS ? '+' : '-');
  return s;
}

// IdentityProfile1.commit :128-144 → CigarClusters.matchCluster :75-102 → copy-ctor :366-405 / merge :633-698
void commit(Run& R, const ReadOut& o, const Scratch& s, int src) {
  const Cfg& cfg = R.cfg;
  const long long ridx = R.n_reads;
  int cl_idx = -1, sub_k = -1;
  if (o.ok) {
    // break points (IdentityProfileHolder.addBreakPoint → BreakPoints.addBreakPoint)
    if (R.calcBreaks1)
      for (size_t i = 0; i + 3 <= o.bp.size(); i += 3) R.pairs[{src, o.bp[i], o.bp[i + 1], o.bp[i + 2]}] += 1;
    for (int i : o.orf_spliced) R.annot.spliced[src][i] += 1;
    for (int i : o.orf_unspliced) R.annot.unspliced[src][i] += 1;
    std::string key = render_key(o.key, cfg, R.annot, R.chrom_index);
    R.totalCounts[src]++;
    auto it = R.l.find(key);
    std::vector<int> br = clone_breaks(s.breaks, cfg.c.include_start, o.forward_known, o.forward, cfg.c.bin);
    Cluster* c;
    if (it == R.l.end()) {
      Cluster nc;
      nc.idx = (int)R.l.size();  // "ID"+chrom_index+"."+l.keySet().size()  (rem_count stays 0 without --sorted)
      nc.key = o.key;
      nc.forward_known = o.forward_known; nc.forward = o.forward; nc.neg = o.neg;
      nc.breaks = br;
      nc.readCount.assign(cfg.c.num_sources, 0);
      nc.first_read = ridx;
      if (cfg.c.record_depth) {
        size_t sz = (size_t)cfg.c.num_sources * R.stride;
        nc.depth.assign(sz, 0); nc.errors.assign(sz, 0); nc.mapStart.assign(sz, 0); nc.mapEnd.assign(sz, 0);
      }
      Count cnt; cnt.count.assign(cfg.c.num_sources, 0); cnt.count[src] = 1; cnt.k = 0; cnt.first_read = ridx;
      cnt.forward = o.forward; cnt.neg = o.neg;
      if (cfg.c.track_break_sums) cnt.addBreaks(s.breaks, src, cfg.c.first_is_transcriptome);
      nc.sub_index[br] = 0; nc.subs.push_back(cnt); nc.sub_keys.push_back(br);
      nc.readCount[src]++; nc.readCountSum++;  // addReadCount
      R.l[key] = (int)R.clusters.size();
      R.clusters.push_back(std::move(nc));
      c = &R.clusters.back();
      sub_k = 0;
    } else {
      c = &R.clusters[it->second];
      auto si = c->sub_index.find(br);
      if (si == c->sub_index.end()) {
        Count cnt; cnt.count.assign(cfg.c.num_sources, 0); cnt.count[src] = 1; cnt.k = (int)c->subs.size(); cnt.first_read = ridx;
        cnt.forward = o.forward; cnt.neg = o.neg;
        c->sub_index[br] = cnt.k; c->subs.push_back(cnt); c->sub_keys.push_back(br);
        sub_k = cnt.k;
      } else {
        sub_k = si->second;
        c->subs[sub_k].count[src] += 1;  // Count.increment
      }
      if (cfg.c.track_break_sums) c->subs[sub_k].addBreaks(s.breaks, src, cfg.c.first_is_transcriptome);
      c->readCount[src] += 1;  // c1.readCount[src] == 1 (updateSourceIndex :458-462)
      c->readCountSum += 1;    // c1.readCountSum == 1 after clear() (CigarCluster.java:432)
    }
    cl_idx = c->idx;
    if (cfg.c.record_depth) {  // SparseVector.merge :82-87, position by position
      int* d = c->depth.data() + (size_t)src * R.stride;
      int* e = c->errors.data() + (size_t)src * R.stride;
      for (size_t i = 0; i < s.ev_pos.size(); i++) { d[s.ev_pos[i]] += 1; if (!s.ev_match[i]) e[s.ev_pos[i]] += 1; }
      int* ms = c->mapStart.data() + (size_t)src * R.stride;
      int* me = c->mapEnd.data() + (size_t)src * R.stride;
      // Dense rows hold positions 0..seqlen+1.  An alignment that runs past the end of the reference (invalid SAM) can put
      // its start/end beyond that; the reference's sparse TreeMap would keep such a key, the dense rows drop it.
      for (int p : s.ms_pos) if (p >= 0 && p < R.stride) ms[p] += 1;
      for (int p : s.me_pos) if (p >= 0 && p < R.stride) me[p] += 1;
    }
  }
  if (R.keep_reads) {
    R.r_cluster.push_back(cl_idx); R.r_sub.push_back(sub_k);
    R.r_start.push_back(o.startPos); R.r_end.push_back(o.endPos); R.r_rst.push_back(o.readSt); R.r_ren.push_back(o.readEn);
    uint32_t f = o.ok ? ((uint32_t)o.type_ind | (o.leader ? NPT_RF_LEADER : 0) | (o.neg ? NPT_RF_NEG : 0)) : NPT_RF_BADCIGAR;
    R.r_flags.push_back(f);
    R.r_maps.push_back(o.maps_sum); R.r_err.push_back(o.err_sum);
    if (o.ok) R.r_breaks.insert(R.r_breaks.end(), s.breaks.begin(), s.breaks.end());
    R.r_boff.push_back(R.r_breaks.size());
  }
  R.n_reads++;
}

}  // namespace

// ================================================================= C entry points (ctypes)
extern "C" {

struct npto_run;  // opaque = Run

void* npto_create(const npt_config* c, int chrom_index, const uint8_t* ref_codes, int64_t ref_len, const npt_annotation* a,
                  int keep_reads) {
  Run* R = new Run();
  R->cfg.c = *c;
  R->cfg.min_first_last_exon_length = c->coronavirus ? 0 : 10;
  R->cfg.recordStartEnd = c->record_depth != 0;
  R->chrom_index = chrom_index;
  R->ref.assign(ref_codes, ref_codes + ref_len);
  R->stride = (int)ref_len + 2;
  R->annot.kind = a->kind;
  R->annot.seqlen = (int)ref_len;
  R->annot.tolerance = c->bin;
  R->annot.round = c->bin;
  R->annot.enforceStrand = c->enforce_strand != 0;
  for (int i = 0; i < a->n; i++) {
    R->annot.start.push_back(a->start[i]); R->annot.end.push_back(a->end[i]);
    R->annot.strand.push_back(a->strand[i]); R->annot.name_id.push_back(a->name_id[i]);
    R->annot.name_hash.push_back(a->kind == NPT_ANNOT_GFF && a->name_hash ? a->name_hash[i] : 0u);
  }
  const int n_orf = a->kind == NPT_ANNOT_GFF ? 0 : a->n;  // ORF counters exist for the CSV table only (Annotation.java:186-187)
  R->annot.spliced.assign(c->num_sources, std::vector<int>(n_orf, 0));
  R->annot.unspliced.assign(c->num_sources, std::vector<int>(n_orf, 0));
  R->calcBreaks1 = c->calc_breaks && c->break_thresh < ref_len;  // ViralTranscriptAnalysisCmd2.java:882
  R->totalCounts.assign(c->num_sources, 0);
  R->keep_reads = keep_reads != 0;
  R->r_boff.push_back(0);
  return R;
}

void npto_destroy(void* h) { delete (Run*)h; }

// Processes one batch.  nthreads > 1 parallelises only the order-free per-read part; commits stay sequential.
void npto_submit(void* h, const npt_batch* b, int nthreads) {
  Run& R = *(Run*)h;
  const int64_t n = b->n;
  const int64_t CH = 1 << 14;
  if (nthreads < 1) nthreads = 1;
  std::vector<ReadOut> outs;
  std::vector<Scratch> scr;
  for (int64_t base = 0; base < n; base += CH) {
    int64_t m = std::min(CH, n - base);
    if (nthreads == 1) {
      Scratch s; ReadOut o;
      for (int64_t i = base; i < base + m; i++) {
        per_read(R.cfg, R.annot, R.ref.data(), (int64_t)R.ref.size(), b->pos[i], b->flag[i], b->src[i], b->cigar + b->cigar_off[i],
                 (int)(b->cigar_off[i + 1] - b->cigar_off[i]), b->seq4 + b->seq_off[i], 2 * (int64_t)(b->seq_off[i + 1] - b->seq_off[i]), s, o);
        commit(R, o, s, b->src[i]);
      }
    } else {
      outs.assign(m, ReadOut()); scr.resize(m);
      std::atomic<int64_t> next(0);
      auto work = [&]() {
        for (;;) {
          int64_t j = next.fetch_add(64);
          if (j >= m) break;
          for (int64_t k = j; k < std::min(j + 64, m); k++) {
            int64_t i = base + k;
            per_read(R.cfg, R.annot, R.ref.data(), (int64_t)R.ref.size(), b->pos[i], b->flag[i], b->src[i],
                     b->cigar + b->cigar_off[i], (int)(b->cigar_off[i + 1] - b->cigar_off[i]), b->seq4 + b->seq_off[i],
                     2 * (int64_t)(b->seq_off[i + 1] - b->seq_off[i]), scr[k],
                     outs[k]);
          }
        }
      };
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back(work);
      for (auto& t : th) t.join();
      for (int64_t k = 0; k < m; k++) commit(R, outs[k], scr[k], b->src[base + k]);
    }
  }
}

int64_t npto_n_reads(void* h) { return ((Run*)h)->n_reads; }
int32_t npto_n_clusters(void* h) { return (int32_t)((Run*)h)->clusters.size(); }
int32_t npto_n_subs(void* h) { int n = 0; for (auto& c : ((Run*)h)->clusters) n += (int)c.subs.size(); return n; }
int64_t npto_n_pairs(void* h) { return (int64_t)((Run*)h)->pairs.size(); }
int64_t npto_n_breaks(void* h) { return (int64_t)((Run*)h)->r_breaks.size(); }
int64_t npto_key_tokens(void* h) { int64_t n = 0; for (auto& c : ((Run*)h)->clusters) n += (int64_t)c.key.size(); return n; }
int64_t npto_sub_key_len(void* h) { int64_t n = 0; for (auto& c : ((Run*)h)->clusters) for (auto& k : c.sub_keys) n += (int64_t)k.size(); return n; }
int64_t npto_sub_sum_len(void* h) { int64_t n = 0; for (auto& c : ((Run*)h)->clusters) for (auto& s : c.subs) n += (int64_t)s.true_breaks.size(); return n; }

// per-read arrays (caller allocates n_reads entries; break_off n_reads+1; breaks n_breaks)
void npto_get_reads(void* h, int32_t* cluster, int32_t* sub, int32_t* start, int32_t* end, int32_t* rst, int32_t* ren, uint32_t* flags,
                    int32_t* maps, int32_t* err, uint64_t* boff, int32_t* breaks) {
  Run& R = *(Run*)h;
  size_t n = R.r_cluster.size();
  memcpy(cluster, R.r_cluster.data(), n * 4); memcpy(sub, R.r_sub.data(), n * 4);
  memcpy(start, R.r_start.data(), n * 4); memcpy(end, R.r_end.data(), n * 4);
  memcpy(rst, R.r_rst.data(), n * 4); memcpy(ren, R.r_ren.data(), n * 4);
  memcpy(flags, R.r_flags.data(), n * 4); memcpy(maps, R.r_maps.data(), n * 4); memcpy(err, R.r_err.data(), n * 4);
  memcpy(boff, R.r_boff.data(), (n + 1) * 8); memcpy(breaks, R.r_breaks.data(), R.r_breaks.size() * 4);
}

// cluster / sub-cluster tables in the layout of npt_results
void npto_get_tables(void* h, int64_t* cl_first, uint32_t* cl_key_off, int32_t* cl_key_tok, int32_t* cl_read_count, uint32_t* cl_sub_off,
                     int64_t* sub_first, uint32_t* sub_key_off, int32_t* sub_key, int32_t* sub_count, int32_t* sub_break_sum,
                     uint32_t* sub_sum_off, int64_t* sub_sums, double* sub_true_breaks, uint8_t* sub_flags, int32_t* total_counts,
                     int32_t* spliced, int32_t* unspliced) {
  Run& R = *(Run*)h;
  const int S = R.cfg.c.num_sources;
  uint32_t ko = 0, so = 0, sko = 0, sso = 0;
  size_t ci = 0;
  for (auto& c : R.clusters) {
    cl_first[ci] = c.first_read; cl_key_off[ci] = ko;
    for (int t : c.key) cl_key_tok[ko++] = t;
    for (int s = 0; s < S; s++) cl_read_count[ci * S + s] = c.readCount[s];
    cl_sub_off[ci] = so;
    for (size_t k = 0; k < c.subs.size(); k++) {
      const Count& cnt = c.subs[k];
      sub_first[so] = cnt.first_read; sub_key_off[so] = sko;
      for (int v : c.sub_keys[k]) sub_key[sko++] = v;
      for (int s = 0; s < S; s++) sub_count[(size_t)so * S + s] = cnt.count[s];
      sub_break_sum[so] = cnt.break_sum;
      sub_sum_off[so] = sso;
      for (size_t j = 0; j < cnt.true_breaks.size(); j++) { sub_sums[sso] = cnt.isum[j]; sub_true_breaks[sso] = cnt.true_breaks[j]; sso++; }
      sub_flags[so] = (uint8_t)((cnt.forward ? 1 : 0) | (cnt.neg ? 2 : 0));
      so++;
    }
    ci++;
  }
  cl_key_off[ci] = ko; cl_sub_off[ci] = so; sub_key_off[so] = sko; sub_sum_off[so] = sso;
  for (int s = 0; s < S; s++) total_counts[s] = R.totalCounts[s];
  const int no = S ? (int)R.annot.spliced[0].size() : 0;
  for (int s = 0; s < S; s++) for (int i = 0; i < no; i++) { spliced[s * no + i] = R.annot.spliced[s][i]; unspliced[s * no + i] = R.annot.unspliced[s][i]; }
}

void npto_get_pairs(void* h, int32_t* src, int32_t* level, int32_t* st, int32_t* en, int32_t* count) {
  Run& R = *(Run*)h;
  size_t i = 0;
  for (auto& kv : R.pairs) { src[i] = kv.first[0]; level[i] = kv.first[1]; st[i] = kv.first[2]; en[i] = kv.first[3]; count[i] = kv.second; i++; }
}

// coverage of one cluster: 4 arrays of [num_sources][seqlen+2]
int npto_get_coverage(void* h, int32_t cluster, int32_t* depth, int32_t* errors, int32_t* map_start, int32_t* map_end) {
  Run& R = *(Run*)h;
  if (!R.cfg.c.record_depth || cluster < 0 || cluster >= (int)R.clusters.size()) return -1;
  const Cluster& c = R.clusters[cluster];
  size_t sz = c.depth.size() * 4;
  memcpy(depth, c.depth.data(), sz); memcpy(errors, c.errors.data(), sz); memcpy(map_start, c.mapStart.data(), sz); memcpy(map_end, c.mapEnd.data(), sz);
  return 0;
}

// Count.getBreaks :99-108 for sub-cluster `so` (global sub index): res[i] = Math.round(true_breaks[i] * (1e3/break_sum))
int npto_sub_consensus(void* h, int32_t so, int32_t* out, int32_t cap) {
  Run& R = *(Run*)h;
  int k = 0;
  for (auto& c : R.clusters) for (auto& cnt : c.subs) {
    if (k == so) {
      if (!cnt.has_tb) return -1;
      double mult = 1e3 / (double)cnt.break_sum;
      int n = (int)cnt.true_breaks.size();
      for (int i = 0; i < n && i < cap; i++) out[i] = (int)std::llround(cnt.true_breaks[i] * mult);  // Math.round (values are positive)
      return n;
    }
    k++;
  }
  return -1;
}

int npto_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
