package npTranscript.cluster;

import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

/**
 * One structure-of-arrays batch of records in pinned host memory (npt_alloc_pinned), laid out as npt_batch wants it
 * (include/npt.h): pos int32[n], flag u16[n], src u16[n], cigar_off u32[n+1], cigar u32[sum] = (len << 4) | op,
 * seq_off u64[n+1], seq4 u8[sum] = BAM 4-bit bases, two per byte, high nibble first, each read byte-aligned.
 * Replaces the per-record hand-over of IdentityProfileHolder.identity1 (reference IdentityProfileHolder.java:170-196).
 */
final class NptBatch {
  final int maxReads;
  final long maxCigar, maxSeqBytes;
  final MemorySegment pos, flag, src, cigarOff, cigar, seqOff, seq4, struct;
  int n = 0;
  long nCigar = 0, nSeq = 0;

  private static MemorySegment pinned(long bytes) {
    try {
      MemorySegment p = (MemorySegment) Npt.allocPinned.invokeExact(bytes);
      if (p.equals(MemorySegment.NULL)) throw new OutOfMemoryError("npt_alloc_pinned(" + bytes + ")");
      return p.reinterpret(bytes);
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }

  /** room for maxReads records of avgCigarOps / avgBases on average */
  NptBatch(int maxReads, int avgCigarOps, int avgBases) {
    this.maxReads = maxReads;
    this.maxCigar = (long) maxReads * avgCigarOps;
    this.maxSeqBytes = (long) maxReads * ((avgBases + 1) / 2);
    pos = pinned(4L * maxReads);
    flag = pinned(2L * maxReads);
    src = pinned(2L * maxReads);
    cigarOff = pinned(4L * (maxReads + 1));
    cigar = pinned(4L * maxCigar);
    seqOff = pinned(8L * (maxReads + 1));
    seq4 = pinned(maxSeqBytes + 16);
    struct = pinned(NptLayouts.NPT_BATCH__SIZEOF);
    clear();
  }

  void clear() {
    n = 0; nCigar = 0; nSeq = 0;
    cigarOff.set(JAVA_INT, 0, 0);
    seqOff.set(JAVA_LONG, 0, 0L);
  }

  boolean fits(int cigarOps, int bases) {
    return n < maxReads && nCigar + cigarOps <= maxCigar && nSeq + (bases + 1) / 2 <= maxSeqBytes;
  }

  /**
   * cigarWords: BAM encoding (htsjdk BinaryCigarCodec.encode(sam.getCigar())); bases: ASCII read bases as SAMRecord.getReadBases()
   * returns them (already on the reference strand, like sam.getReadString() that the reference walks).
   */
  void add(int alignmentStart, int samFlags, int sourceIndex, int[] cigarWords, byte[] bases) {
    pos.set(JAVA_INT, 4L * n, alignmentStart);
    flag.set(JAVA_SHORT, 2L * n, (short) samFlags);
    src.set(JAVA_SHORT, 2L * n, (short) sourceIndex);
    for (int w : cigarWords) cigar.set(JAVA_INT, 4L * nCigar++, w);
    for (int i = 0; i < bases.length; i += 2) {
      int hi = Npt.code4((char) bases[i]);
      int lo = i + 1 < bases.length ? Npt.code4((char) bases[i + 1]) : 0;
      seq4.set(JAVA_BYTE, nSeq++, (byte) ((hi << 4) | lo));
    }
    n++;
    cigarOff.set(JAVA_INT, 4L * n, (int) nCigar);
    seqOff.set(JAVA_LONG, 8L * n, nSeq);
  }

  /** fills the npt_batch struct and returns it */
  MemorySegment asStruct() {
    struct.fill((byte) 0);
    struct.set(JAVA_LONG, NptLayouts.NPT_BATCH__N, (long) n);
    struct.set(JAVA_INT, NptLayouts.NPT_BATCH__LOCATION, Npt.MEM_HOST);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__POS, pos);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__FLAG, flag);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__SRC, src);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__CIGAR_OFF, cigarOff);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__CIGAR, cigar);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__SEQ_OFF, seqOff);
    struct.set(java.lang.foreign.ValueLayout.ADDRESS, NptLayouts.NPT_BATCH__SEQ4, seq4);
    return struct;
  }

  void free() {
    try {
      for (MemorySegment m : new MemorySegment[] {pos, flag, src, cigarOff, cigar, seqOff, seq4, struct}) Npt.freePinned.invokeExact(m);
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }
}
