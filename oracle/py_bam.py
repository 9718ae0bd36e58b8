"""TEST INFRASTRUCTURE ONLY — an independent, pure-Python reading of a BAM file and of the reference's per-record filter
chain (J/run/ViralTranscriptAnalysisCmd2.java:712-830, 936-959), used to check libnptbam.  PARITY UNPINNED: the
reference ships no BAM fixtures and htsjdk is not available here; BAM/BGZF layouts follow the SAM/BAM specification.

BGZF is a series of gzip members, so Python's gzip module reads it as one stream."""
import gzip
import math
import struct


def read_bam(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<I", data, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<I", data, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<I", data, o)
        name = data[o + 4:o + 4 + l_name - 1].decode()
        ln, = struct.unpack_from("<I", data, o + 4 + l_name)
        refs.append((name, ln))
        o += 8 + l_name
    recs = []
    while o < len(data):
        bs, = struct.unpack_from("<I", data, o)
        ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack_from("<iiBBHHHIiii", data, o + 4)
        p = o + 36
        name = data[p:p + l_name - 1].decode()
        p += l_name
        cigar = list(struct.unpack_from(f"<{n_cig}I", data, p))
        p += 4 * n_cig
        seq = data[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        qual = data[p:p + l_seq]
        recs.append(dict(ref_id=ref_id, pos=pos + 1, mapq=mapq, flag=flag, name=name, cigar=cigar, seq4=seq, l_seq=l_seq, qual=qual))
        o += 4 + bs
    return refs, recs


def filter_chain(recs, fail_thresh=0.0, min_mapq=0, max_reads=0, ref_include=None):
    """Yields (record, q1) for the records that reach profile.identity1, plus 'switch' markers are implied by ref_id changes
    of the yielded-or-secondary stream; returns through StopIteration the counters."""
    counts = dict(unmapped=0, excluded=0, lowq=0, short=0, lowmapq=0, secondary=0, stopped=False)
    out = []
    num_reads = 0
    for r in recs:
        if r["flag"] & 0x4:
            counts["unmapped"] += 1
            continue
        if ref_include is not None and not (0 <= r["ref_id"] < len(ref_include) and ref_include[r["ref_id"]]):
            counts["excluded"] += 1
            continue
        q1 = 500.0
        b = r["qual"]
        if len(b) > 0 and b[0] != 0xFF:
            sump = 0.0
            for v in b:
                v = v - 256 if v > 127 else v  # Java bytes are signed
                sump += math.pow(10, -v / 10.0)
            sump = sump / float(len(b))
            q1 = -10.0 * math.log10(sump)
        if q1 < fail_thresh:
            counts["lowq"] += 1
            continue
        if r["l_seq"] <= 1:
            counts["short"] += 1
            continue
        num_reads += 1
        if max_reads > 0 and num_reads > max_reads:
            counts["stopped"] = True
            break
        if r["mapq"] < min_mapq:
            counts["lowmapq"] += 1
            continue
        if r["flag"] & 0x900:
            counts["secondary"] += 1
            out.append((r, q1, False))  # still triggers a chromosome switch in the reference's loop
            continue
        out.append((r, q1, True))
    return out, counts
