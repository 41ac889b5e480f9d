"""Read -> isoform probability matrices (host side).

Vectorised equivalent of what `get_psi` obtains from `TranSplice` / `TranUnits`
(`diceseq/utils/tran_utils.py:72-237,239-322`, `sam_utils.py:23-149`) in the mode the CLI
hard-codes (`mate_mode="single"`, uniform bias, `diceseq/diceseq.py:150-153`): for one gene
and one time point the boolean compatibility matrix `Rmat` (reads x isoforms), the per
read per isoform probability `proU` (NaN where incompatible) and the effective lengths.

Only the uniform model is implemented; bias modes need a genome FASTA and a bias file,
neither of which ships with the reference data (`--bias` other than `unif` is refused).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# reads of a gene region                                     sam_utils.py:23-149
# --------------------------------------------------------------------------------------
def fetch_reads(bam, chrom, start, end, rm_duplicate=True, inner_only=True, mapq_min=0,
                mismatch_max=5, rlen_min=1, is_mated=True):
    """Filtered reads of [start, end) split into mated pairs (reads1/reads2, aligned lists) and
    unmated mates (reads1u/reads2u), with the reference's filters and mate matching."""
    chrom = str(chrom)
    if chrom not in bam.references:
        parts = chrom.split("chr")
        chrom = parts[0] if len(parts) <= 1 else parts[1]
    try:
        reads = bam.fetch(chrom, start, end)
    except ValueError:
        print("Cannot fetch reads in region: %s:%d-%d" % (chrom, start, end))
        reads = []
    q1, q2, r1s, r2s = [], [], [], []
    prev = None
    for r in reads:
        if rm_duplicate and prev is not None and prev.qname == r.qname and prev.positions == r.positions:
            prev = r
            continue
        prev = r
        if inner_only and (r.pos < start or r.aend > end):
            continue
        if r.mapq < mapq_min or r.rlen - len(r.positions) > mismatch_max or r.rlen < rlen_min:
            continue
        (r2s if r.is_read2 else r1s).append(r)
        (q2 if r.is_read2 else q1).append(r.qname)
    # mate suffix: if all first-mate names end with the same character, strip one character from all
    if len(q1) > 0 and all(q1[i][-1] == q1[i + 1][-1] for i in range(len(q1) - 1)):
        q1 = [q[:-1] for q in q1]
        q2 = [q[:-1] for q in q2]
    out = {"reads1": [], "reads2": [], "reads1u": [], "reads2u": []}
    if not is_mated:
        out["reads1u"], out["reads2u"] = r1s, r2s
        return out
    i1 = sorted(range(len(q1)), key=q1.__getitem__)
    i2 = sorted(range(len(q2)), key=q2.__getitem__)
    a = b = 0
    while a < len(i1) and b < len(i2):
        n1, n2 = q1[i1[a]], q2[i2[b]]
        if n1 == n2:
            out["reads1"].append(r1s[i1[a]]); out["reads2"].append(r2s[i2[b]])
            a += 1; b += 1
        elif n1 < n2:
            out["reads1u"].append(r1s[i1[a]]); a += 1
        else:
            out["reads2u"].append(r2s[i2[b]]); b += 1
    out["reads1u"] += [r1s[k] for k in i1[a:]]
    out["reads2u"] += [r2s[k] for k in i2[b:]]
    return out


# --------------------------------------------------------------------------------------
# one transcript                                             tran_utils.py:14-237
# --------------------------------------------------------------------------------------
class _Units:
    def __init__(self, transcript):
        self.strand_plus = transcript.strand in ("+", "1")
        self.exons = np.asarray(transcript.exons, dtype=np.int64)
        lens = self.exons[:, 1] - self.exons[:, 0] + 1
        self.cum = np.concatenate([[0], np.cumsum(lens)])       # transcript offset of each exon start
        self.ulen = int(self.cum[-1])

    def index(self, loc):
        """(offset on the transcript, exon id) per genomic position; exon id -2 / -3 before /
        after the transcript, -1 inside an intron (tran_utils.py:72-85)."""
        loc = np.asarray(loc, dtype=np.int64)
        ex = np.searchsorted(self.exons[:, 0], loc, side="right") - 1      # last exon starting <= loc
        exc = np.clip(ex, 0, len(self.exons) - 1)
        inside = (ex >= 0) & (loc <= self.exons[exc, 1])
        off = np.where(inside, self.cum[exc] + loc - self.exons[exc, 0], -1)
        eid = np.where(inside, exc, -1)
        before, after = loc < self.exons[0, 0], loc > self.exons[-1, 1]
        off = np.where(before, 0, np.where(after, self.ulen - 1, off))
        eid = np.where(before, -2, np.where(after, -3, eid))
        return off, eid

    def read_columns(self, reads, flen_mean, flen_std):
        """Column of Rmat / proU of this transcript for single-end reads (one mate each).

        A forward read on a + transcript (reverse on -) plays the 5' mate, the other the 3'
        mate; either way it is compatible iff both ends fall into exons and the spliced span
        equals the aligned length within 3 bases (tran_utils.py:87-146)."""
        n = len(reads)
        if n == 0:
            return np.ones(0, bool), np.ones(0)
        pos = np.array([r.pos for r in reads], dtype=np.int64)
        last = np.array([r.aend - 1 for r in reads], dtype=np.int64)
        qlen = np.array([r.qlen for r in reads], dtype=np.int64)
        mapq = np.array([r.mapq for r in reads], dtype=np.float64)
        o_a, e_a = self.index(pos)
        o_b, e_b = self.index(last)
        span = np.abs(o_a - o_b) + 1
        ok = (e_a >= 0) & (e_b >= 0) & (span <= qlen + 3) & (span >= qlen - 3)
        prob = 1.0 - 10 ** (0 - mapq / 10.0)
        flen = span.astype(np.float64)
        # fragment-length distribution of this transcript from its compatible reads (:203-215)
        probs = np.zeros(self.ulen)
        fl_ok = flen[ok]
        uniq = np.unique(fl_ok)
        if ok.sum() == 0:
            probs[0] = 1.0
        elif uniq.shape[0] >= 10:
            sd = np.std(fl_ok) if flen_std is None else flen_std
            mu = np.mean(fl_ok) if flen_mean is None else flen_mean
            x = np.arange(self.ulen) + 1
            probs[:] = 1 / (sd * np.sqrt(2 * np.pi)) * np.exp(-1.0 / 2 * ((x - mu) / sd) ** 2)
            if probs.sum() > 0:
                probs /= probs.sum()
        else:
            for v in uniq:
                probs[int(v) - 1] = np.mean(fl_ok == v)
        pro = np.full(n, np.nan)
        fl = fl_ok.astype(np.int64)
        pro[ok] = prob[ok] * probs[fl - 1] / (self.ulen - fl + 1)       # (:233-236)
        return ok, pro


def gene_time_matrices(gene, bams, flen_mean=None, flen_std=None):
    """(Rmat (N,C) bool, proU (N,C) float with NaN, efflen (C,), n_counted) of one gene at one
    time point; `bams` = the replicate files of that time point (run_utils.py:62-92)."""
    r1u, r2u, mated1, mated2 = [], [], [], []
    for bam in bams:
        got = fetch_reads(bam, gene.chrom, gene.start, gene.stop)
        r1u += got["reads1u"]; r2u += got["reads2u"]
        mated1 += got["reads1"]; mated2 += got["reads2"]
    # mate_mode "single": the mated reads are appended after ALL files' unmated ones
    # (tran_utils.py:282-287)
    r1u += mated1; r2u += mated2
    C = len(gene.trans)
    N = len(r1u) + len(r2u)
    Rmat = np.ones((N, C), dtype=bool)
    proU = np.ones((N, C), dtype=float)
    efflen = np.zeros(C)
    for i, tr in enumerate(gene.trans):
        u = _Units(tr)
        for lo, group in ((0, r1u), (len(r1u), r2u)):
            if not group:
                continue
            ok, pro = u.read_columns(group, flen_mean, flen_std)
            Rmat[lo:lo + len(group), i] = ok
            proU[lo:lo + len(group), i] = pro
        efflen[i] = u.ulen          # every length class counts once: efflen_unif == transcript length (:218,:314-322)
    count = int(np.sum(Rmat.sum(axis=1) > 0))
    return Rmat, proU, efflen, count
