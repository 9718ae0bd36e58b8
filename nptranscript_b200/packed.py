"""Host-side packer for the compact transfer format (npt_packed_batch, include/npt.h): 16-bit CIGAR words with a list of
wide elements, 2-bit bases with a list of exception bytes.  Pure repacking of arrays the caller already holds (numpy, no
compute of the path); the device expands it back bit-identically (csrc/unpack.cuh).  The BAM packer emits the same
format directly from BAM records (include/npt_bam.h)."""
import numpy as np

# seq4 byte (two BAM 4-bit codes) -> 4-bit pair of 2-bit codes, and whether both nibbles are one of A C G T (1 2 4 8)
_CODE = np.zeros(16, np.uint8)
_OK = np.zeros(16, bool)
for _c, _v in ((1, 0), (2, 1), (4, 2), (8, 3)):
    _CODE[_c], _OK[_c] = _v, True
_PAIR = np.array([(_CODE[b >> 4] << 2) | _CODE[b & 15] for b in range(256)], np.uint8)
_PAIR_OK = np.array([_OK[b >> 4] and _OK[b & 15] for b in range(256)], bool)


class PackedBatch:
    def __init__(self, pos, flag, src, cigar_off, cigar16, wide_idx, wide_word, seq_off, seq2, exc_byte, exc_val):
        self.pos, self.flag, self.src, self.cigar_off, self.cigar16 = pos, flag, src, cigar_off, cigar16
        self.wide_idx, self.wide_word, self.seq_off, self.seq2, self.exc_byte, self.exc_val = wide_idx, wide_word, seq_off, seq2, exc_byte, exc_val

    @property
    def n(self):
        return len(self.pos)

    def arrays(self):
        return (self.pos, self.flag, self.src, self.cigar_off, self.cigar16, self.wide_idx, self.wide_word, self.seq_off, self.seq2,
                self.exc_byte, self.exc_val)

    def nbytes(self):
        return sum(a.nbytes for a in self.arrays())


def pack(b, alloc=None):
    """synth.Batch-like (npt_batch arrays) -> PackedBatch.  alloc(nbytes) -> writable buffer (e.g. pinned memory)."""
    def place(a):
        if alloc is None:
            return np.ascontiguousarray(a)
        out = np.frombuffer(alloc(max(1, a.nbytes)), dtype=a.dtype, count=a.size)
        out[...] = a
        return out

    cig = np.asarray(b.cigar, np.uint32)
    wide = (cig >> 4) >= 4095
    c16 = np.where(wide, 0xFFF0 | (cig & 15), cig & 0xFFFF).astype(np.uint16)
    widx = np.nonzero(wide)[0].astype(np.uint32)
    seq4 = np.asarray(b.seq4, np.uint8)
    nseq = int(b.seq_off[-1]) if len(b.seq_off) else 0
    seq4 = seq4[:nseq]
    bad = ~_PAIR_OK[seq4]
    pair = _PAIR[seq4]
    if nseq & 1:
        pair = np.concatenate([pair, np.zeros(1, np.uint8)])
    seq2 = ((pair[0::2] << 4) | pair[1::2]).astype(np.uint8)
    eidx = np.nonzero(bad)[0].astype(np.uint64)
    return PackedBatch(place(np.asarray(b.pos, np.int32)), place(np.asarray(b.flag, np.uint16)), place(np.asarray(b.src, np.uint16)),
                       place(np.asarray(b.cigar_off, np.uint32)), place(c16), place(widx), place(cig[wide]), place(np.asarray(b.seq_off, np.uint64)),
                       place(seq2), place(eidx), place(seq4[bad]))


def unpack(pb):
    """The expansion the device performs, in numpy (tests: pack -> unpack is the identity)."""
    cig = pb.cigar16.astype(np.uint32)
    cig[pb.wide_idx] = pb.wide_word
    nseq = int(pb.seq_off[-1]) if len(pb.seq_off) else 0
    c = np.stack([(pb.seq2 >> 6) & 3, (pb.seq2 >> 4) & 3, (pb.seq2 >> 2) & 3, pb.seq2 & 3], axis=1).astype(np.uint8)
    hot = (1 << c).astype(np.uint8)
    seq4 = np.stack([(hot[:, 0] << 4) | hot[:, 1], (hot[:, 2] << 4) | hot[:, 3]], axis=1).reshape(-1)[:nseq].copy()
    seq4[pb.exc_byte.astype(np.int64)] = pb.exc_val
    return cig, seq4
