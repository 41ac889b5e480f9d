"""Minimal pysam-free BAM access for the `diceseq` drop-in.

The reference reads alignments through pysam (`diceseq/utils/sam_utils.py:7-20,55-64`,
`pysam.idxstats` in `diceseq/diceseq.py:141-147`); pysam is not installable here.  This
module decodes BGZF/BAM with the standard library and numpy and offers exactly the pieces
the read model consumes: `references`, `fetch(chrom, start, end)`,
per read `qname, pos, aend, mapq, rlen, qlen, is_read2, is_reverse, positions`, and the
mapped-read totals of `idxstats`.

With a .bai index next to the file (`IndexedBam`) a region fetch inflates only the BGZF blocks the
index points to -- what pysam's fetch does -- and the idxstats totals come from the index's pseudo-bins;
a small cache of inflated blocks serves neighbouring genes.  Without an index (`BamFile`) the whole file is
decoded once and kept in memory.  Both check that the file is coordinate sorted.  SAM/CRAM are not supported.
"""
from __future__ import annotations

import gzip
import os
import zlib
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


def _parse_header(read):
    """references / lengths from a byte reader positioned at the start of the uncompressed stream."""
    if read(4) != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack("<i", read(4))[0]
    read(l_text)
    n_ref = struct.unpack("<i", read(4))[0]
    refs, lens = [], []
    for _ in range(n_ref):
        l_name = struct.unpack("<i", read(4))[0]
        refs.append(read(l_name)[:-1].decode())
        lens.append(struct.unpack("<i", read(4))[0])
    return refs, lens


def _parse_record(buf, rec):
    """Read object of the alignment record whose core block starts at buf[rec] (after block_size)."""
    tid, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", buf, rec)
    p = rec + 32
    qname = buf[p:p + l_read_name - 1].decode()
    p += l_read_name
    cig = struct.unpack_from("<%dI" % n_cigar, buf, p) if n_cigar else ()
    return Read(qname, flag, tid, pos, mapq, l_seq, tuple((c & 0xF, c >> 4) for c in cig))


class BamFile:
    """In-memory BAM (no index needed): `references`, `fetch`, `idxstats`.  Fine for small files; a file with a
    .bai next to it is opened as `IndexedBam` instead, which inflates only the blocks a region needs."""

    def __init__(self, path):
        self.path = path
        with gzip.open(path, "rb") as f:                 # BGZF = concatenated gzip members
            data = f.read()
        pos_ = [0]

        def read(n):
            out = data[pos_[0]:pos_[0] + n]
            pos_[0] += n
            return out
        try:
            self.references, self.lengths = _parse_header(read)
        except ValueError:
            raise ValueError("%s is not a BAM file" % path)
        off = pos_[0]
        n_ref = len(self.references)
        self._reads = [[] for _ in range(n_ref)]
        self.n_unplaced = 0
        n = len(data)
        last = (-1, -1)
        while off + 4 <= n:
            block_size = struct.unpack_from("<i", data, off)[0]
            rec = off + 4
            off = rec + block_size
            r = _parse_record(data, rec)
            if r.tid < 0:
                self.n_unplaced += 1
                continue
            if (r.tid, r.pos) < last:
                raise ValueError("%s is not sorted by coordinate (needed for region fetches)" % path)
            last = (r.tid, r.pos)
            self._reads[r.tid].append(r)
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


# --------------------------------------------------------------------------------------
# indexed access: BGZF virtual offsets + BAI (SAM specification sections 4.1 and 5.2)
# --------------------------------------------------------------------------------------
class _Bgzf:
    """Random access into a BGZF file by virtual offset (compressed block offset << 16 | offset in block)."""

    def __init__(self, path, cache_blocks=64):
        self.f = open(path, "rb")
        self.cache, self.order, self.cap = {}, [], cache_blocks
        self.coff = self.uoff = 0
        self.block, self.next_coff = b"", 0

    def _load(self, coff):
        if coff in self.cache:
            return self.cache[coff]
        self.f.seek(coff)
        hdr = self.f.read(12)
        if len(hdr) < 12:
            return b"", coff
        if hdr[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at offset %d" % coff)
        xlen = struct.unpack_from("<H", hdr, 10)[0]
        extra = self.f.read(xlen)
        bsize, p = None, 0
        while p + 4 <= xlen:
            si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", extra, p + 4)[0] + 1
            p += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without a BC field at offset %d" % coff)
        payload = self.f.read(bsize - 12 - xlen)
        data = zlib.decompressobj(-15).decompress(payload[:-8])
        out = (data, coff + bsize)
        self.cache[coff] = out
        self.order.append(coff)
        if len(self.order) > self.cap:
            self.cache.pop(self.order.pop(0), None)
        return out

    def seek(self, voff):
        self.coff, self.uoff = voff >> 16, voff & 0xFFFF
        self.block, self.next_coff = self._load(self.coff)

    def tell(self):
        return (self.coff << 16) | self.uoff

    def read(self, n):
        out = []
        while n > 0:
            if self.uoff >= len(self.block):
                if not self.block and self.next_coff == self.coff:
                    break
                self.coff, self.uoff = self.next_coff, 0
                self.block, self.next_coff = self._load(self.coff)
                if not self.block:
                    break
            piece = self.block[self.uoff:self.uoff + n]
            self.uoff += len(piece)
            n -= len(piece)
            out.append(piece)
        return b"".join(out)


def _reg2bins(beg, end):
    """Bins of the UCSC scheme that may hold reads overlapping [beg, end) (SAM spec 5.3)."""
    end -= 1
    bins = [0]
    for shift, first in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(first + (beg >> shift), first + (end >> shift) + 1))
    return bins


class IndexedBam:
    """Coordinate-sorted BAM with a .bai index: a region fetch inflates only the BGZF blocks the index points
    to (what pysam.AlignmentFile.fetch does, sam_utils.py:55-64), idxstats come from the index's pseudo-bins."""

    def __init__(self, path, bai):
        self.path = path
        self.bgzf = _Bgzf(path)
        self.bgzf.seek(0)
        try:
            self.references, self.lengths = _parse_header(self.bgzf.read)
        except ValueError:
            raise ValueError("%s is not a BAM file" % path)
        with open(bai, "rb") as f:
            idx = f.read()
        if idx[:4] != b"BAI\x01":
            raise ValueError("%s is not a BAI index" % bai)
        n_ref = struct.unpack_from("<i", idx, 4)[0]
        if n_ref != len(self.references):
            raise ValueError("%s does not belong to %s" % (bai, path))
        off = 8
        self.bins, self.linear, self.counts = [], [], []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", idx, off)[0]
            off += 4
            bins, counts = {}, None
            for _b in range(n_bin):
                bin_id, n_chunk = struct.unpack_from("<Ii", idx, off)
                off += 8
                chunks = struct.unpack_from("<%dQ" % (2 * n_chunk), idx, off)
                off += 16 * n_chunk
                if bin_id == 37450 and n_chunk == 2:
                    counts = (chunks[2], chunks[3])                      # mapped, unmapped (samtools pseudo-bin)
                else:
                    bins[bin_id] = [(chunks[2 * i], chunks[2 * i + 1]) for i in range(n_chunk)]
            n_intv = struct.unpack_from("<i", idx, off)[0]
            off += 4
            self.linear.append(struct.unpack_from("<%dQ" % n_intv, idx, off))
            off += 8 * n_intv
            self.bins.append(bins)
            self.counts.append(counts)
        self.n_no_coor = struct.unpack_from("<Q", idx, off)[0] if off + 8 <= len(idx) else 0

    def fetch(self, chrom, start, end):
        if chrom not in self.references:
            raise ValueError("invalid contig %s" % chrom)
        tid = self.references.index(chrom)
        start = max(int(start), 0)
        end = int(end)
        if end <= start:
            return []
        lin = self.linear[tid]
        w = start >> 14
        min_off = lin[min(w, len(lin) - 1)] if lin else 0
        chunks = sorted(c for b in _reg2bins(start, end) for c in self.bins[tid].get(b, ()) if c[1] > min_off)
        out, last_pos = [], -1
        done_upto = 0
        for beg, stop in chunks:
            beg = max(beg, min_off, done_upto)
            if beg >= stop:
                continue
            self.bgzf.seek(beg)
            while self.bgzf.tell() < stop:
                head = self.bgzf.read(4)
                if len(head) < 4:
                    break
                body = self.bgzf.read(struct.unpack("<i", head)[0])
                r = _parse_record(body, 0)
                if r.tid != tid or r.pos >= end:
                    done_upto = stop
                    break
                if r.pos < last_pos:
                    raise ValueError("%s is not sorted by coordinate" % self.path)
                last_pos = r.pos
                if not r.is_unmapped and r.aend is not None and r.aend > start:
                    out.append(r)
            done_upto = max(done_upto, self.bgzf.tell() if self.bgzf.tell() < stop else stop)
            if out and out[-1].pos >= end:
                break
        return out

    def idxstats(self):
        if any(c is None for c in self.counts):                            # index without the count pseudo-bins
            return BamFile(self.path).idxstats()
        rows = [(n, ln, int(c[0]), int(c[1])) for n, ln, c in zip(self.references, self.lengths, self.counts)]
        rows.append(("*", 0, 0, int(self.n_no_coor)))
        return rows

    def mapped_total(self):
        return float(sum(r[2] for r in self.idxstats()))


_cache = {}


def load_samfile(path):
    """Open (once per process) an alignment file; only BAM is supported.  With a .bai index next to it
    (file.bam.bai or file.bai) region fetches go through the index, otherwise the whole file is decoded."""
    if not path.endswith(".bam"):
        raise ValueError("only .bam alignment files are supported without pysam: %s" % path)
    if path not in _cache:
        bai = next((b for b in (path + ".bai", path[:-4] + ".bai") if os.path.isfile(b)), None)
        _cache[path] = IndexedBam(path, bai) if bai else BamFile(path)
    return _cache[path]
