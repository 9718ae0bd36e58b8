"""Random CIGAR/read generator for parity tests: every op of MIDNSHP=X, zero-length ops, long ops, long mismatch runs,
alignments running past the genome end, clips, junction-sized N ops, negative strand."""
import numpy as np

from nptranscript_b200.synth import Batch

OPS = "MIDNSHP=X"


def random_batch(rng, n, genome_len, genome_codes, num_sources=1, weird=0.3, max_ops=60, mismatch=0.1, big_n=2000, bad_op=0.0,
                 long_reads=False):
    pos, flag, src, coff, cig, soff, seq = [], [], [], [0], [], [0], bytearray()
    c4 = np.array([1, 2, 4, 8], dtype=np.uint8)
    for _ in range(n):
        p = int(rng.integers(1, genome_len + 1)) if rng.random() < 0.85 else int(rng.integers(max(1, genome_len - 300), genome_len + 1))
        if rng.random() < 0.3:
            p = int(rng.integers(1, 120))
        nops = int(rng.integers(1, max_ops + 1))
        if long_reads and rng.random() < 0.3:
            nops = int(rng.integers(300, 900))
        ops = []
        if rng.random() < 0.5:
            ops.append((int(rng.integers(1, 40)), 4 if rng.random() < 0.8 else 5))
        refpos = p - 1
        bases = []
        for _k in range(nops):
            u = rng.random()
            if u < weird * 0.15:
                op = int(rng.integers(0, 9))
            elif u < 0.55:
                op = 0
            elif u < 0.70:
                op = 1
            elif u < 0.85:
                op = 2
            elif u < 0.90:
                op = 3
            elif u < 0.95:
                op = 7
            else:
                op = 8
            if op == 3:
                ln = int(rng.integers(1, big_n)) if rng.random() < 0.7 else int(rng.integers(1, 40))
            elif op == 0 or op == 7:
                r = rng.random()
                ln = int(rng.integers(1, 30)) if r < 0.8 else int(rng.integers(30, 400)) if r < 0.97 else int(rng.integers(400, 1500))
            else:
                ln = int(rng.geometric(0.5)) if rng.random() < 0.9 else int(rng.integers(1, 80))
            if rng.random() < weird * 0.02:
                ln = 0
            if bad_op and rng.random() < bad_op:
                op = int(rng.integers(9, 16))
            ops.append((ln, op))
        if rng.random() < 0.5:
            ops.append((int(rng.integers(1, 60)), 4))
        # bases: follow the reference on M with noise, sometimes a long mismatch run
        refpos = p - 1
        for (ln, op) in ops:
            if op in (0, 7, 8):
                idx = np.arange(refpos, refpos + ln)
                g = genome_codes[np.clip(idx, 0, genome_len - 1)].copy()
                mm = rng.random(ln) < (mismatch if rng.random() < 0.9 else 0.9)
                alt = c4[rng.integers(0, 4, ln)]
                g[mm] = alt[mm]
                if rng.random() < 0.02:
                    g[:] = 15  # N bases: never equal to ACGT reference
                bases.extend(g.tolist())
                refpos += ln
            elif op in (1, 4):
                bases.extend(c4[rng.integers(0, 4, ln)].tolist())
            elif op in (2, 3):
                refpos += ln
        pos.append(p); flag.append(16 if rng.random() < 0.3 else 0); src.append(int(rng.integers(0, num_sources)))
        cig.extend(((ln << 4) | op) for ln, op in ops)
        coff.append(len(cig))
        if len(bases) % 2:
            bases.append(0)
        b = np.array(bases, dtype=np.uint8)
        seq.extend(((b[0::2] << 4) | b[1::2]).tobytes() if len(b) else b"")
        soff.append(len(seq))
    seq4 = np.frombuffer(bytes(seq) + b"\0" * 16, dtype=np.uint8).copy()[: len(seq)] if len(seq) else np.zeros(0, np.uint8)
    return Batch(np.array(pos, np.int32), np.array(flag, np.uint16), np.array(src, np.uint16), np.array(coff, np.uint32),
                 np.array(cig, np.uint32), np.array(soff, np.uint64), seq4)


def gff_case(rng, n_reads, genome_len, genome_str, n_genes=40, num_sources=2, jitter=6, hash_collide=False):
    """Host-mode case on a stand-in chromosome: a random exon table (several transcripts per gene, overlapping genes, both
    strands, rows in 'file order' = gene by gene) and spliced reads that follow exon chains with jittered ends, skipped and
    novel exons, plus reads between genes.  Returns (Annotation.gff, synth.Batch).  hash_collide: gene names chosen so that
    many land in the same HashSet bucket (insertion-order dependence of the key)."""
    from nptranscript_b200 import synth, workloads
    genes, start, end, strand, chains = [], [], [], [], []
    names = []
    if hash_collide:
        # Java: "Aa".hashCode() == "BB".hashCode(); concatenations of such pairs collide in full
        import itertools
        names = ["".join(p) for p in itertools.product(("Aa", "BB"), repeat=4)][:n_genes]
    while len(names) < n_genes:
        names.append(f"ENSG{int(rng.integers(0, 10**8)):011d}")
    for gi in range(n_genes):
        g0 = int(rng.integers(200, genome_len - 4000))
        fwd = bool(rng.random() < 0.5)
        for _t in range(int(rng.integers(1, 4))):
            p = g0 + int(rng.integers(0, 60))
            chain = []
            for _e in range(int(rng.integers(1, 7))):
                ln = int(rng.integers(30, 300))
                if p + ln >= genome_len - 50:
                    break
                chain.append((p, p + ln - 1))
                genes.append(names[gi]); start.append(p); end.append(p + ln - 1); strand.append(fwd)
                p += ln + int(rng.integers(70, 600))
            if chain:
                chains.append((chain, fwd))
    reads = []
    for _ in range(n_reads):
        src = int(rng.integers(0, num_sources))
        u = rng.random()
        if u < 0.8:
            chain, fwd = chains[int(rng.integers(0, len(chains)))]
            ex = [list(e) for e in chain if rng.random() > 0.15] or [list(chain[0])]
            if rng.random() < 0.2 and len(chains) > 1:  # chimeric: append exons of another gene further right
                other, _f = chains[int(rng.integers(0, len(chains)))]
                ex += [list(e) for e in other if e[0] > ex[-1][1] + 80]
            for e in ex:
                e[0] += int(rng.integers(-jitter, jitter + 1)); e[1] += int(rng.integers(-jitter, jitter + 1))
                if rng.random() < 0.1:
                    e[1] += int(rng.integers(15, 60))  # novel exon end: outside 2*tolerance for small bins
            ex = [e for e in ex if e[1] - e[0] >= 5 and e[0] >= 1 and e[1] < genome_len]
            if not ex:
                continue
            flag = 0 if (fwd if rng.random() < 0.9 else not fwd) else 16
        else:
            p = int(rng.integers(1, genome_len - 700))
            ex = [[p, p + int(rng.integers(20, 300))]]
            if rng.random() < 0.5:
                q = ex[0][1] + int(rng.integers(60, 300))
                ex.append([q, q + int(rng.integers(5, 200))])
            flag = 16 if rng.random() < 0.5 else 0
        cig, bases, last = "", [], None
        for (a, b) in ex:
            if last is not None:
                if a <= last + 1:
                    continue
                cig += f"{a - last - 1}N"
            seg = list(genome_str[a - 1:b])
            for k in range(len(seg)):
                if rng.random() < 0.05:
                    seg[k] = "ACGT"[int(rng.integers(0, 4))]
            cig += f"{len(seg)}M"
            bases += seg
            last = b
        reads.append((ex[0][0], flag, src, cig, "".join(bases)))
    return workloads.Annotation.gff(genes, start, end, strand), synth.from_reads(reads)
