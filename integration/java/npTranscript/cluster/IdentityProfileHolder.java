package npTranscript.cluster;

import java.io.IOException;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.List;
import java.util.SortedSet;
import java.util.TreeSet;

import htsjdk.samtools.BinaryCigarCodec;
import htsjdk.samtools.SAMRecord;
import japsa.seq.Sequence;
import npTranscript.run.ViralTranscriptAnalysisCmd2;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

/**
 * Drop-in replacement of npTranscript.cluster.IdentityProfileHolder (reference IdentityProfileHolder.java:106-128, 170-196,
 * 101-104, 152-157): same constructor, identity1(...), printBreakPoints(), getConsensus(), waitOnThreads(...),
 * shutDownExecutor() — the surface ViralTranscriptAnalysisCmd2.errorAnalysis calls (:885-886, :967, :840-843, :989-991,
 * :477).  Records are buffered into structure-of-arrays batches in pinned memory and walked / clustered on the GPU(s) by
 * libnpt.so; at the end of the chromosome the result tables are rendered through the reference's own Outputs writers.
 *
 * Semantics: the --maxThreads=1 order (records are numbered in call order), which is the only order in which the reference's
 * own cluster ids are defined.  -Dnpt.devices=0,1,... selects the GPUs (default "0"); more than one goes through npt_mg_*.
 *
 * Not restated here (stay in the host exactly as before, or are default-off at HEAD): supplementary alignments, 5'/3' rescue,
 * polyA split, leftover fastq, MSA bins, the NW break-point score matrices.
 */
public class IdentityProfileHolder {
  // ---- one native context per JVM run (the reference keeps the same flags in static fields)
  private static MemorySegment ctx = MemorySegment.NULL;
  private static boolean mg = false;
  private static int numSourcesOfCtx = -1;

  final CigarClusters all_clusters;   // kept for Outputs.writeTotalString / annotation access
  final BreakPoints bp;
  final Outputs o;
  final Sequence genome;
  final String chrom_;
  final int chrom_index;
  final String[] type_nmes;
  final int num_sources;
  final Annotation annot;
  final SortedSet<String> geneNames = new TreeSet<String>();

  private final Arena arena = Arena.ofConfined();
  private NptBatch batch;
  private final List<String> readNames = new ArrayList<String>();
  private final List<Double> qvals = new ArrayList<Double>();

  public static void waitOnThreads(int sleep) { /* the library drains in getConsensus(); nothing runs on Java threads */ }

  public static void shutDownExecutor() {
    if (ctx.equals(MemorySegment.NULL)) return;
    try {
      if (mg) Npt.mgDestroy.invokeExact(ctx); else Npt.destroy.invokeExact(ctx);
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
    ctx = MemorySegment.NULL;
  }

  /** npt_config from the static option fields (ViralTranscriptAnalysisCmd2.java:221-386, 496-511) */
  private static synchronized void ensureContext(int numSources) throws Throwable {
    if (!ctx.equals(MemorySegment.NULL)) {
      if (numSources != numSourcesOfCtx) throw new IllegalStateException("number of sources changed");
      return;
    }
    if ((int) Npt.abiVersion.invokeExact() != NptLayouts.NPT_ABI_VERSION) throw new UnsatisfiedLinkError("libnpt ABI version");
    Arena g = Arena.global();
    MemorySegment cfg = g.allocate(NptLayouts.NPT_CONFIG__SIZEOF, 8);
    cfg.fill((byte) 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__ABI_VERSION, NptLayouts.NPT_ABI_VERSION);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__BIN, (int) CigarHash2.round);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__BREAK_THRESH, IdentityProfile1.break_thresh);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__INCLUDE_START, IdentityProfile1.includeStartEnd ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__CORONAVIRUS, TranscriptUtils.coronavirus ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__START_THRESH, TranscriptUtils.startThresh);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__END_THRESH, TranscriptUtils.endThresh);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__ENFORCE_STRAND, Annotation.enforceStrand ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__RECORD_DEPTH, CigarCluster.recordDepthByPosition ? 1 : 0);
    // calcBreaks is true in virus mode and false in host mode (ViralTranscriptAnalysisCmd2.java:341, 370); the per-chromosome
    // condition break_thresh < seqlen (:882) is applied inside the library
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__CALC_BREAKS, TranscriptUtils.coronavirus ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__FIRST_IS_TRANSCRIPTOME, Outputs.firstIsTranscriptome ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__TRACK_BREAK_SUMS, (Outputs.writeGFF || Outputs.writeIsoforms) ? 1 : 0);
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__NUM_SOURCES, numSources);
    for (int s = 0; s < numSources; s++) {
      cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__RNA + 4L * s, ViralTranscriptAnalysisCmd2.RNA[s] ? 1 : 0);
    }
    String[] dev = System.getProperty("npt.devices", "0").split(",");
    cfg.set(JAVA_INT, NptLayouts.NPT_CONFIG__DEVICE, Integer.parseInt(dev[0].trim()));
    MemorySegment out = g.allocate(ADDRESS);
    if (dev.length == 1) {
      int rc = (int) Npt.create.invokeExact(cfg, out);
      if (rc != Npt.OK) Npt.check(rc, MemorySegment.NULL, false);
      mg = false;
    } else {
      MemorySegment devs = g.allocate(4L * dev.length, 4);
      for (int i = 0; i < dev.length; i++) devs.set(JAVA_INT, 4L * i, Integer.parseInt(dev[i].trim()));
      int rc = (int) Npt.mgCreate.invokeExact(cfg, devs, dev.length, out);
      if (rc != Npt.OK) Npt.check(rc, MemorySegment.NULL, true);
      mg = true;
    }
    ctx = out.get(ADDRESS, 0);
    numSourcesOfCtx = numSources;
  }

  public IdentityProfileHolder(ArrayList<Sequence> genomes, Sequence refSeq, Outputs o, String[] in_nmes, boolean calcBreakpoints,
                               int chrom_index, Annotation annot) throws IOException {
    this.genome = refSeq;
    this.chrom_ = refSeq.getName();
    this.chrom_index = chrom_index;
    this.type_nmes = in_nmes;
    this.num_sources = in_nmes.length;
    this.o = o;
    this.annot = annot;
    this.all_clusters = new CigarClusters(refSeq, num_sources, annot);
    this.bp = calcBreakpoints ? new BreakPoints(num_sources, refSeq, genomes) : null;
    try {
      ensureContext(num_sources);
      int G = refSeq.length();
      MemorySegment codes = arena.allocate(G, 1);
      for (int i = 0; i < G; i++) codes.set(JAVA_BYTE, i, Npt.code4(refSeq.charAt(i)));
      MemorySegment an = annotationStruct(annot);
      int rc = mg ? (int) Npt.mgBeginChrom.invokeExact(ctx, chrom_index, codes, (long) G, an)
                  : (int) Npt.beginChrom.invokeExact(ctx, chrom_index, codes, (long) G, an);
      Npt.check(rc, ctx, mg);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IOException(t);
    }
    this.batch = new NptBatch(1 << 20, 256, 2048);
  }

  /** npt_annotation: rows in file order; equal gene strings share the id of their first row (include/npt.h) */
  private MemorySegment annotationStruct(Annotation a) {
    MemorySegment st = arena.allocate(NptLayouts.NPT_ANNOTATION__SIZEOF, 8);
    st.fill((byte) 0);
    if (a instanceof EmptyAnnotation) {
      st.set(JAVA_INT, NptLayouts.NPT_ANNOTATION__KIND, Npt.ANNOT_EMPTY);
      return st;
    }
    boolean gff = a instanceof GFFAnnotation;
    int n = a.start.size();
    MemorySegment start = arena.allocate(4L * n, 4), end = arena.allocate(4L * n, 4), strand = arena.allocate(n, 1);
    MemorySegment nameId = arena.allocate(4L * n, 4), nameHash = gff ? arena.allocate(4L * n, 4) : MemorySegment.NULL;
    java.util.HashMap<String, Integer> first = new java.util.HashMap<String, Integer>();
    for (int i = 0; i < n; i++) {
      start.set(JAVA_INT, 4L * i, a.start.get(i));
      end.set(JAVA_INT, 4L * i, a.end.get(i));
      strand.set(JAVA_BYTE, i, (byte) (a.strand.get(i) ? 1 : 0));
      String g = a.genes.get(i);
      Integer id = first.get(g);
      if (id == null) { id = i; first.put(g, i); }
      nameId.set(JAVA_INT, 4L * i, id);
      if (gff) { int h = g.hashCode(); nameHash.set(JAVA_INT, 4L * i, h ^ (h >>> 16)); }   // HashMap.hash(): decides HashSet iteration order
    }
    st.set(JAVA_INT, NptLayouts.NPT_ANNOTATION__KIND, gff ? Npt.ANNOT_GFF : Npt.ANNOT_CSV);
    st.set(JAVA_INT, NptLayouts.NPT_ANNOTATION__N, n);
    st.set(ADDRESS, NptLayouts.NPT_ANNOTATION__START, start);
    st.set(ADDRESS, NptLayouts.NPT_ANNOTATION__END, end);
    st.set(ADDRESS, NptLayouts.NPT_ANNOTATION__STRAND, strand);
    st.set(ADDRESS, NptLayouts.NPT_ANNOTATION__NAME_ID, nameId);
    st.set(ADDRESS, NptLayouts.NPT_ANNOTATION__NAME_HASH, nameHash);
    return st;
  }

  private void flush() {
    if (batch.n == 0) return;
    try {
      int rc = mg ? (int) Npt.mgSubmit.invokeExact(ctx, batch.asStruct()) : (int) Npt.submit.invokeExact(ctx, batch.asStruct());
      Npt.check(rc, ctx, mg);
      // npt_submit is asynchronous; the batch arrays are handed to the copy engine.  Simplest ownership rule: a fresh
      // batch per submit, the old ones are freed after getConsensus().
      pendingFree.add(batch);
      batch = new NptBatch(1 << 20, 256, 2048);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }
  private final List<NptBatch> pendingFree = new ArrayList<NptBatch>();

  /** reference signature (IdentityProfileHolder.java:170-171); pool / supp are unused on the paths that are live at HEAD */
  public void identity1(Sequence readSeq, SAMRecord sam, int source_index, boolean cluster_reads, String pool, double qval, boolean supp) {
    int[] cig = BinaryCigarCodec.encode(sam.getCigar());
    byte[] bases = sam.getReadBases();
    if (!batch.fits(cig.length, bases.length)) flush();
    batch.add(sam.getAlignmentStart(), sam.getFlags(), source_index, cig, bases);
    readNames.add(sam.getReadName());
    qvals.add(qval);
  }

  /** deferred: the break-point counts arrive with the other tables in getConsensus() and are written there */
  public void printBreakPoints() throws IOException {}

  public void getConsensus() throws IOException {
    flush();
    try {
      Npt.check(mg ? (int) Npt.mgEndChrom.invokeExact(ctx) : (int) Npt.endChrom.invokeExact(ctx), ctx, mg);
      MemorySegment rs = arena.allocate(NptLayouts.NPT_RESULTS__SIZEOF, 8);
      Npt.check(mg ? (int) Npt.mgFetch.invokeExact(ctx, Npt.FETCH_ALL, rs) : (int) Npt.fetch.invokeExact(ctx, Npt.FETCH_ALL, rs), ctx, mg);
      NptResults r = new NptResults(rs, num_sources);
      NptRender.readToCluster(r, this, readNames, qvals);          // Outputs.printRead rows (IdentityProfile1.java:133-142)
      NptRender.breakPoints(r, this);                              // BreakPoints.addBreakPoint counts + printBreakPoints (:102-160)
      NptRender.consensus(r, this);                                // CigarClusters.getConsensus / process1 rows (:138-196, 230-262)
      Npt.check(mg ? (int) Npt.mgRelease.invokeExact(ctx) : (int) Npt.release.invokeExact(ctx), ctx, mg);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IOException(t);
    }
    for (NptBatch b : pendingFree) b.free();
    pendingFree.clear();
    batch.free();
    readNames.clear();
    qvals.clear();
    arena.close();
  }
}
