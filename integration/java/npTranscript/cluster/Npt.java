package npTranscript.cluster;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (Java 22) binding of include/npt.h — the C ABI of libnpt.so, the B200 implementation of what runs behind
 * IdentityProfileHolder.identity1 / printBreakPoints / getConsensus.  One static handle per entry point; struct offsets come
 * from NptLayouts (generated from the header with gcc offsetof, checked by tests/test_java_shim.py).
 *
 * Load path: -Dnpt.lib=/path/to/libnpt.so, else "libnpt.so" through the platform loader (LD_LIBRARY_PATH).
 * Run the JVM with --enable-native-access=ALL-UNNAMED.
 */
final class Npt {
  private Npt() {}

  static final int OK = 0, EINVAL = -1, ENODEVICE = -2, ECUDA = -3, ENOMEM = -4, ECAPACITY = -5, ESTATE = -6;
  static final int ANNOT_CSV = 0, ANNOT_EMPTY = 1, ANNOT_GFF = 2;
  static final int MEM_HOST = 0;
  static final int FETCH_ALL = 7;
  static final int TOK_ST = -1, TOK_END = -2, TOK_PLUS = -3, TOK_MINUS = -4, TOK_NOENDS = -5, TOK_GENESET = -6;
  static final int RF_TYPE_MASK = 0x3, RF_LEADER = 0x4, RF_NEG = 0x8, RF_BADCIGAR = 0x10;

  private static final Linker LINKER = Linker.nativeLinker();
  private static final SymbolLookup LIB =
      SymbolLookup.libraryLookup(System.getProperty("npt.lib", "libnpt.so"), Arena.global());

  private static MethodHandle h(String name, FunctionDescriptor d) {
    return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
  }

  static final MethodHandle abiVersion = h("npt_abi_version", FunctionDescriptor.of(JAVA_INT));
  static final MethodHandle create = h("npt_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle destroy = h("npt_destroy", FunctionDescriptor.ofVoid(ADDRESS));
  static final MethodHandle lastError = h("npt_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle beginChrom =
      h("npt_begin_chrom", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
  static final MethodHandle submit = h("npt_submit", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  // the compact transfer format (npt_packed_batch: 16-bit CIGAR words, 2-bit bases): half the bytes over PCIe
  static final MethodHandle submitPacked = h("npt_submit_packed", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle waitAll = h("npt_wait", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle endChrom = h("npt_end_chrom", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle fetch = h("npt_fetch_results", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle release = h("npt_release_results", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle allocPinned = h("npt_alloc_pinned", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
  static final MethodHandle freePinned = h("npt_free_pinned", FunctionDescriptor.ofVoid(ADDRESS));
  // several GPUs behind one handle (include/npt.h, npt_mg_*): same call sequence, reads are dealt to the devices in
  // contiguous ranges of the submit order and the tables are merged on the devices at the end of the chromosome
  static final MethodHandle mgCreate = h("npt_mg_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle mgDestroy = h("npt_mg_destroy", FunctionDescriptor.ofVoid(ADDRESS));
  static final MethodHandle mgLastError = h("npt_mg_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle mgBeginChrom =
      h("npt_mg_begin_chrom", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
  static final MethodHandle mgSubmit = h("npt_mg_submit", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle mgSubmitPacked = h("npt_mg_submit_packed", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle mgEndChrom = h("npt_mg_end_chrom", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle mgFetch = h("npt_mg_fetch_results", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle mgRelease = h("npt_mg_release_results", FunctionDescriptor.of(JAVA_INT, ADDRESS));

  /** BAM 4-bit code of a reference / read base: "=ACMGRSVTWYHKDBN".indexOf(upper case), anything else 15 (N). */
  static byte code4(char c) {
    int i = "=ACMGRSVTWYHKDBN".indexOf(Character.toUpperCase(c));
    return (byte) (i < 0 ? 15 : i);
  }

  /** throws with the library's message unless rc == NPT_OK */
  static void check(int rc, MemorySegment ctx, boolean mg) {
    if (rc == OK) return;
    String msg;
    try {
      MemorySegment p = (MemorySegment) (mg ? mgLastError : lastError).invokeExact(ctx);
      msg = p.equals(MemorySegment.NULL) ? "" : p.reinterpret(4096).getString(0);
    } catch (Throwable t) {
      msg = t.toString();
    }
    throw new RuntimeException("libnpt error " + rc + ": " + msg);
  }
}
