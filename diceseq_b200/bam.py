"""Minimal pysam-free BAM access for the `diceseq` drop-in.

The reference reads alignments through pysam (`diceseq/utils/sam_utils.py:7-20,55-64`,
`pysam.idxstats` in `diceseq/diceseq.py:141-147`); pysam is not installable here.  This
module decodes BGZF/BAM with the standard library and numpy and offers exactly the pieces
the read model consumes: `references`, `fetch(chrom, start, end)`,
per read `qname, pos, aend, mapq, rlen, qlen, is_read2, is_reverse, positions`, and the
mapped-read totals of `idxstats`.

The whole file is decoded once and kept in memory, indexed per reference by start position;
that is the right trade for per-gene region queries (the reference re-opens every BAM for
every gene, `run_utils.py:52-57`).  SAM/CRAM are not supported.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

_CIGAR_REF = (True, False, True, True, False, False, False, True, True)      # M I D N S H P = X consume reference
_CIGAR_QRY = (True, True, False, False, True, False, False, True, True)      # consume query
_CIGAR_ALN = (True, False, False, False, False, False, False, True, True)    # aligned (M = X): pysam `positions`


class Read:
    """One alignment record with the attribute names pysam's AlignedSegment uses."""
    __slots__ = ("qname", "flag", "tid", "pos", "aend", "mapq", "rlen", "qlen", "cigar", "_positions")

    def __init__(self, qname, flag, tid, pos, mapq, rlen, cigar):
        self.qname, self.flag, self.tid, self.pos, self.mapq, self.rlen = qname, flag, tid, pos, mapq, rlen
        self.cigar = cigar
        ref = 0
        qlen = 0
        for op, n in cigar:
            if _CIGAR_REF[op]:
                ref += n
            if _CIGAR_QRY[op] and op != 4:
                qlen += n
        self.aend = pos + ref if cigar else None        # one past the last aligned base (0-based)
        self.qlen = qlen                                # query_alignment_length (no soft clips)
        self._positions = None

    @property
    def is_read2(self):
        return bool(self.flag & 0x80)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def positions(self):
        """Reference positions of the aligned bases (pysam get_reference_positions)."""
        if self._positions is None:
            out, ref = [], self.pos
            for op, n in self.cigar:
                if _CIGAR_ALN[op]:
                    out.extend(range(ref, ref + n))
                if _CIGAR_REF[op]:
                    ref += n
            self._positions = out
        return self._positions


class BamFile:
    """In-memory BAM: `references`, `fetch`, `idxstats`."""

    def __init__(self, path):
        self.path = path
        with gzip.open(path, "rb") as f:                 # BGZF = concatenated gzip members
            data = f.read()
        if data[:4] != b"BAM\x01":
            raise ValueError("%s is not a BAM file" % path)
        l_text = struct.unpack_from("<i", data, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from("<i", data, off)[0]
        off += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", data, off)[0]
            off += 4
            self.references.append(data[off:off + l_name - 1].decode())
            off += l_name
            self.lengths.append(struct.unpack_from("<i", data, off)[0])
            off += 4
        self._reads = [[] for _ in range(n_ref)]
        self.n_unplaced = 0
        n = len(data)
        while off + 4 <= n:
            block_size = struct.unpack_from("<i", data, off)[0]
            rec = off + 4
            off = rec + block_size
            tid, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, rec)
            p = rec + 32
            qname = data[p:p + l_read_name - 1].decode()
            p += l_read_name
            cig = struct.unpack_from("<%dI" % n_cigar, data, p) if n_cigar else ()
            cigar = tuple((c & 0xF, c >> 4) for c in cig)
            if tid < 0:
                self.n_unplaced += 1
                continue
            self._reads[tid].append(Read(qname, flag, tid, pos, mapq, l_seq, cigar))
        self._starts = [np.array([r.pos for r in rs], dtype=np.int64) for rs in self._reads]
        self._max_span = [max((r.aend - r.pos for r in rs if r.aend is not None), default=0) for rs in self._reads]

    def fetch(self, chrom, start, end):
        """Reads overlapping the 0-based half-open interval [start, end), in file order
        (pysam AlignmentFile.fetch).  Raises ValueError for an unknown reference."""
        if chrom not in self.references:
            raise ValueError("invalid contig %s" % chrom)
        tid = self.references.index(chrom)
        starts = self._starts[tid]
        lo = int(np.searchsorted(starts, start - self._max_span[tid], side="left"))
        hi = int(np.searchsorted(starts, end, side="left"))
        out = []
        for r in self._reads[tid][lo:hi]:
            if r.is_unmapped or r.aend is None:
                continue
            if r.aend > start and r.pos < end:
                out.append(r)
        return out

    def idxstats(self):
        """[(reference, length, mapped, unmapped)] like `samtools idxstats` (plus the '*' row)."""
        rows = []
        for name, ln, rs in zip(self.references, self.lengths, self._reads):
            mapped = sum(1 for r in rs if not r.is_unmapped)
            rows.append((name, ln, mapped, len(rs) - mapped))
        rows.append(("*", 0, 0, self.n_unplaced))
        return rows

    def mapped_total(self):
        """Sum of the third idxstats column: TOTAL_READ of diceseq.py:141-147."""
        return float(sum(r[2] for r in self.idxstats()))


_cache = {}


def load_samfile(path):
    """Open (once per process) an alignment file; only BAM is supported."""
    if not path.endswith(".bam"):
        raise ValueError("only .bam alignment files are supported without pysam: %s" % path)
    if path not in _cache:
        _cache[path] = BamFile(path)
    return _cache[path]
