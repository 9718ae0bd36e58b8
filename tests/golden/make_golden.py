"""Writes tests/golden/cases.json: small fixed inputs with the outputs of the pure-Python restatement
(oracle/py_oracle.py: Java-style string keys and ids), so that oracle.cpp, the renderers and the CUDA path are all held to
one frozen set of answers.  The reference itself cannot run here (no JVM) and ships no fixtures, so these are
"golden" only relative to the restatement — the KAT cases among them are the hand-derived ones of SURVEY §8c.

    python tests/golden/make_golden.py        # regenerates cases.json (deterministic)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nptranscript_b200 import synth, workloads  # noqa: E402
from oracle import py_oracle  # noqa: E402
from tests.fuzz import gff_case, random_batch  # noqa: E402

OPS = "MIDNSHP=X"
LUT = "=ACMGRSVTWYHKDBN"


def batch_to_reads(b):
    out = []
    for pos, flag, src, cig, bases in py_oracle.unpack_batch(b):
        out.append([int(pos), int(flag), int(src), "".join(f"{ln}{OPS[op] if op < 9 else '?'}" for ln, op in cig), "".join(LUT[c] for c in bases)])
    return out


def cov_digest(d):
    """[entries, sum, position-weighted checksum] of a sparse {(src, pos): count} row set."""
    return [len(d), int(sum(d.values())), int(sum((s + 1) * (p + 7) * v for (s, p), v in d.items()) % ((1 << 61) - 1))]


def run_case(name, kw, genome, annotation, reads, chrom_index):
    fl = py_oracle.Flags(bin=kw.get("bin", 1), break_thresh=kw.get("break_thresh", 1000), include_start=bool(kw.get("include_start", 1)),
                         coronavirus=bool(kw.get("coronavirus", 1)), start_thresh=kw.get("start_thresh", 100), end_thresh=kw.get("end_thresh", 100),
                         enforce_strand=bool(kw.get("enforce_strand", 0)), record_depth=bool(kw.get("record_depth", 0)),
                         num_sources=kw.get("num_sources", 1), rna=tuple(bool(x) for x in kw.get("rna", [1] * 16)),
                         first_is_transcriptome=bool(kw.get("first_is_transcriptome", 0)))
    ann = py_oracle.Annot(annotation.genes, annotation.start, annotation.end, annotation.strand, len(genome), fl, empty=annotation.kind == 1,
                          chrom_index=chrom_index, gff=annotation.kind == 2)
    run = py_oracle.Run(fl, genome, ann, chrom_index)
    b = synth.from_reads([tuple(r) for r in reads])
    for i, rd in enumerate(py_oracle.unpack_batch(b)):
        run.process(*rd, read_name=f"read{i}")
    rows = []
    for row in run.rows:
        rows.append(None if not row["ok"] else dict(cluster=row["cluster"], sub=row["sub"], start_pos=row["start_pos"], end_pos=row["end_pos"],
                                                     read_st=row["read_st"], read_en=row["read_en"], type_ind=row["type_ind"], leader=bool(row["leader"]),
                                                     neg=bool(row["neg"]), breaks=row["breaks"], maps_sum=row["maps_sum"], err_sum=row["err_sum"]))
    clusters = []
    for cl in run.clusters.values():
        full = name.startswith("kats")
        # dense rows hold positions 0..seqlen+1: entries beyond (alignments that run off the reference) are dropped
        inside = lambda d: {(s, p): v for (s, p), v in d.items() if 0 <= p < len(genome) + 2}
        sp = (lambda d: sorted([int(s), int(p), int(v)] for (s, p), v in inside(d).items())) if full else (lambda d: cov_digest(inside(d)))
        clusters.append(dict(id=cl.id, key=cl.key, read_count=cl.read_count,
                             subs=[dict(id=s["id"], key=list(br), count=s["count"], break_sum=s["break_sum"]) for br, s in cl.subs.items()],
                             depth=sp(cl.depth), errors=sp(cl.errors), map_start=sp(cl.map_start), map_end=sp(cl.map_end)))
    return dict(name=name, flags=kw, chrom_index=chrom_index, annotation=dict(kind=annotation.kind, genes=annotation.genes, start=annotation.start,
                                                                              end=annotation.end, strand=[bool(x) for x in annotation.strand]),
                reads=reads, rows=rows, clusters=clusters, total_counts=run.total_counts,
                pairs=sorted([list(k) + [v] for k, v in run.pairs.items()]), spliced=ann.spliced, unspliced=ann.unspliced)


def main():
    w = workloads.SARS2
    g = synth.genome_codes(w)
    gl = g.tolist()
    allA = [1] * w.genome_len
    lut = np.full(16, ord("N"), np.uint8)
    lut[[1, 2, 4, 8]] = [ord(c) for c in "ACGT"]
    gstr = lut[g].tobytes().decode()
    cases = []
    kat_reads = [
        [5, 0, 0, "10S65M28190N1600M30S", "A" * 1705],                       # KAT1
        [5, 0, 0, "65M28190N1600M", "A" * 63 + "CCCCC" + "A" * 1597],        # KAT2: mismatches around the junction
        [100, 0, 0, "5M2D5M", "A" * 10],                                     # KAT3: deletion quirk
        [200, 16, 0, "4=2X4=", "A" * 10],                                    # KAT4: = / X quirk, negative strand
        [29885, 0, 0, "20M", "A" * 20],                                      # KAT5: clamp at the reference end
    ]
    cases.append(run_case("kats_all_A_genome", dict(bin=10, break_thresh=1000, record_depth=1), allA, w.annotation, kat_reads, 0))
    rng = np.random.default_rng(20200301)
    cases.append(run_case("virus_synthetic_40", dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2), gl, w.annotation,
                          batch_to_reads(synth.generate(w, 20200301, 40, num_sources=2)), 0))
    cases.append(run_case("virus_fuzz_strand_T50", dict(bin=100, break_thresh=50, record_depth=1, enforce_strand=1, num_sources=2), gl, w.annotation,
                          batch_to_reads(random_batch(rng, 60, w.genome_len, g, num_sources=2, max_ops=30)), 1))
    cases.append(run_case("virus_fuzz_small_T", dict(bin=7, break_thresh=3, record_depth=1, include_start=0), gl, w.annotation,
                          batch_to_reads(random_batch(rng, 40, w.genome_len, g, max_ops=10)), 0))
    cases.append(run_case("host_empty_fit", dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, num_sources=2, rna=[1, 1], first_is_transcriptome=1),
                          gl, workloads.Annotation.empty(), batch_to_reads(random_batch(rng, 60, w.genome_len, g, num_sources=2, big_n=300)), 4))
    ann, b = gff_case(rng, 80, w.genome_len, gstr, n_genes=12, hash_collide=True)
    cases.append(run_case("host_gff", dict(bin=10, break_thresh=50, coronavirus=0, include_start=0, record_depth=1, num_sources=2, rna=[1, 0]), gl, ann,
                          batch_to_reads(b), 3))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cases.json")
    with open(out, "w") as f:
        json.dump(dict(note="frozen outputs of oracle/py_oracle.py; regenerate with tests/golden/make_golden.py", genome="workloads.SARS2 synthetic genome "
                       "(or all 'A' for the KAT case)", cases=cases), f, separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes", [(c["name"], len(c["reads"]), len(c["clusters"])) for c in cases])


if __name__ == "__main__":
    main()
