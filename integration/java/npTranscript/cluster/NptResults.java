package npTranscript.cluster;

import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Read-only view of npt_results (include/npt.h): everything the reference's writers consume, as the library left it in host
 * memory.  Array accessors take (index) and read straight from the native buffers; nothing is copied.
 */
final class NptResults {
  final MemorySegment s;
  final int numSources;

  NptResults(MemorySegment struct, int numSources) {
    this.s = struct;
    this.numSources = numSources;
  }

  private MemorySegment arr(long fieldOffset, long bytes) {
    MemorySegment p = s.get(ADDRESS, fieldOffset);
    return p.equals(MemorySegment.NULL) ? MemorySegment.NULL : p.reinterpret(bytes);
  }
  private int i32(long fieldOffset, long index) { return s.get(ADDRESS, fieldOffset).reinterpret(4 * (index + 1)).get(JAVA_INT, 4 * index); }
  private long i64(long fieldOffset, long index) { return s.get(ADDRESS, fieldOffset).reinterpret(8 * (index + 1)).get(JAVA_LONG, 8 * index); }

  long nReads() { return s.get(JAVA_LONG, NptLayouts.NPT_RESULTS__N_READS); }
  int nClusters() { return s.get(JAVA_INT, NptLayouts.NPT_RESULTS__N_CLUSTERS); }
  int nSubs() { return s.get(JAVA_INT, NptLayouts.NPT_RESULTS__N_SUBS); }
  int nOrf() { return s.get(JAVA_INT, NptLayouts.NPT_RESULTS__N_ORF); }
  long nPairs() { return s.get(JAVA_LONG, NptLayouts.NPT_RESULTS__N_PAIRS); }
  int covStride() { return s.get(JAVA_INT, NptLayouts.NPT_RESULTS__COV_STRIDE); }

  // ---- per read (submit order)
  int clusterId(long r) { return i32(NptLayouts.NPT_RESULTS__CLUSTER_ID, r); }
  int subId(long r) { return i32(NptLayouts.NPT_RESULTS__SUB_ID, r); }
  int startPos(long r) { return i32(NptLayouts.NPT_RESULTS__START_POS, r); }
  int endPos(long r) { return i32(NptLayouts.NPT_RESULTS__END_POS, r); }
  int readSt(long r) { return i32(NptLayouts.NPT_RESULTS__READ_ST, r); }
  int readEn(long r) { return i32(NptLayouts.NPT_RESULTS__READ_EN, r); }
  int readFlags(long r) { return i32(NptLayouts.NPT_RESULTS__READ_FLAGS, r); }
  int mapsSum(long r) { return i32(NptLayouts.NPT_RESULTS__MAPS_SUM, r); }
  int errSum(long r) { return i32(NptLayouts.NPT_RESULTS__ERR_SUM, r); }
  long breakOff(long r) { return i64(NptLayouts.NPT_RESULTS__BREAK_OFF, r); }
  int breakAt(long k) { return i32(NptLayouts.NPT_RESULTS__BREAKS, k); }
  int[] breaks(long r) {
    long a = breakOff(r), b = breakOff(r + 1);
    int[] out = new int[(int) (b - a)];
    for (int i = 0; i < out.length; i++) out[i] = breakAt(a + i);
    return out;
  }

  // ---- per cluster (index = cluster id)
  long clFirstRead(int c) { return i64(NptLayouts.NPT_RESULTS__CL_FIRST_READ, c); }
  int[] clKeyTokens(int c) {
    int a = i32(NptLayouts.NPT_RESULTS__CL_KEY_OFF, c), b = i32(NptLayouts.NPT_RESULTS__CL_KEY_OFF, c + 1);
    int[] out = new int[b - a];
    for (int i = 0; i < out.length; i++) out[i] = i32(NptLayouts.NPT_RESULTS__CL_KEY_TOK, a + i);
    return out;
  }
  int clReadCount(int c, int src) { return i32(NptLayouts.NPT_RESULTS__CL_READ_COUNT, (long) c * numSources + src); }
  int clSubOff(int c) { return i32(NptLayouts.NPT_RESULTS__CL_SUB_OFF, c); }

  // ---- per sub-cluster (contiguous per cluster, in .t order)
  long subFirstRead(int k) { return i64(NptLayouts.NPT_RESULTS__SUB_FIRST_READ, k); }
  int[] subKey(int k) {
    int a = i32(NptLayouts.NPT_RESULTS__SUB_KEY_OFF, k), b = i32(NptLayouts.NPT_RESULTS__SUB_KEY_OFF, k + 1);
    int[] out = new int[b - a];
    for (int i = 0; i < out.length; i++) out[i] = i32(NptLayouts.NPT_RESULTS__SUB_KEY, a + i);
    return out;
  }
  int subCount(int k, int src) { return i32(NptLayouts.NPT_RESULTS__SUB_COUNT, (long) k * numSources + src); }
  int subBreakSum(int k) { return i32(NptLayouts.NPT_RESULTS__SUB_BREAK_SUM, k); }
  long[] subSums(int k) {
    int a = i32(NptLayouts.NPT_RESULTS__SUB_SUM_OFF, k), b = i32(NptLayouts.NPT_RESULTS__SUB_SUM_OFF, k + 1);
    long[] out = new long[b - a];
    for (int i = 0; i < out.length; i++) out[i] = i64(NptLayouts.NPT_RESULTS__SUB_SUMS, a + i);
    return out;
  }
  int subFlags(int k) { return arr(NptLayouts.NPT_RESULTS__SUB_FLAGS, k + 1).get(JAVA_BYTE, k); }

  // ---- per chromosome
  int totalCount(int src) { return i32(NptLayouts.NPT_RESULTS__TOTAL_COUNTS, src); }
  int spliced(int src, int orf) { return i32(NptLayouts.NPT_RESULTS__SPLICED, (long) src * nOrf() + orf); }
  int unspliced(int src, int orf) { return i32(NptLayouts.NPT_RESULTS__UNSPLICED, (long) src * nOrf() + orf); }
  int pairSrc(long i) { return i32(NptLayouts.NPT_RESULTS__PAIR_SRC, i); }
  int pairLevel(long i) { return i32(NptLayouts.NPT_RESULTS__PAIR_LEVEL, i); }
  int pairStart(long i) { return i32(NptLayouts.NPT_RESULTS__PAIR_START, i); }
  int pairEnd(long i) { return i32(NptLayouts.NPT_RESULTS__PAIR_END, i); }
  int pairCount(long i) { return i32(NptLayouts.NPT_RESULTS__PAIR_COUNT, i); }

  // ---- coverage rows [cluster][source][covStride], index = 1-based position
  private int cov(long field, int c, int src, int pos) { return i32(field, ((long) c * numSources + src) * covStride() + pos); }
  int depth(int c, int src, int pos) { return cov(NptLayouts.NPT_RESULTS__DEPTH, c, src, pos); }
  int errors(int c, int src, int pos) { return cov(NptLayouts.NPT_RESULTS__ERRORS, c, src, pos); }
  int mapStart(int c, int src, int pos) { return cov(NptLayouts.NPT_RESULTS__MAP_START, c, src, pos); }
  int mapEnd(int c, int src, int pos) { return cov(NptLayouts.NPT_RESULTS__MAP_END, c, src, pos); }
}
