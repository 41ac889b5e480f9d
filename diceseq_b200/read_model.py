"""Read -> isoform probability matrices (host side).

Vectorised equivalent of what `get_psi` obtains from `TranSplice` / `TranUnits`
(`diceseq/utils/tran_utils.py:72-237,239-322`, `sam_utils.py:23-149`) in the mode the CLI
hard-codes (`mate_mode="single"`, `diceseq/diceseq.py:150-153`): for one gene and one time point
the boolean compatibility matrix `Rmat` (reads x isoforms), the per read per isoform probability
(NaN where incompatible) and the effective lengths -- `proU` of the uniform model, or `proB` of the
bias modes end5 / end3 / both, where the sequence weights of a bias file (bias.py) at the read's 5' or
3' end and the bias-weighted fragment positions enter (tran_utils.py:38-70,195-198,218-236).
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

    # ---- bias mode: genome sequence and per-locus weights --------------- tran_utils.py:25-70
    def set_sequence(self, fasta, chrom):
        """Exon sequence with 20 more bases on either side (genomic orientation)."""
        seq = fasta.get_seq(chrom, int(self.exons[0, 0]) - 20, int(self.exons[0, 0]) - 1)
        for a, b in self.exons:
            seq += fasta.get_seq(chrom, int(a), int(b))
        seq += fasta.get_seq(chrom, int(self.exons[-1, 1]) + 1, int(self.exons[-1, 1]) + 20)
        self.seq = seq

    def set_bias(self, bias):
        """bias5 / bias3 of every locus from the sequence model (get_psi always asks for mode "seq",
        run_utils.py:73): on the + strand the 5' window is seq[i+12 : i+33] and the 3' window
        seq[i+8 : i+29] reversed (i = locus, the sequence carries 20 leading bases); swapped on -."""
        from .bias import encode_bases
        codes = encode_bases(self.seq)
        if len(codes) != self.ulen + 40:
            raise ValueError("the FASTA does not cover the transcript with 20 bases on either side")
        centre = np.arange(self.ulen) + 20
        fwd5 = bias.seq_bias_track(codes, centre, 5, reverse=not self.strand_plus)
        fwd3 = bias.seq_bias_track(codes, centre, 3, reverse=self.strand_plus)
        self.bias5, self.bias3 = fwd5, fwd3

    def read_columns(self, reads, flen_mean, flen_std, bias_mode="unif"):
        """Column of Rmat / probability of this transcript for single-end reads (one mate each), and the
        transcript's bias-weighted effective length of that read set (0 in the uniform model).

        A forward read on a + transcript (reverse on -) plays the 5' mate, the other the 3'
        mate; either way it is compatible iff both ends fall into exons and the spliced span
        equals the aligned length within 3 bases (tran_utils.py:87-146)."""
        n = len(reads)
        if n == 0:
            return np.ones(0, bool), np.ones(0), 0.0
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
        if bias_mode == "unif":
            pro[ok] = prob[ok] * (probs[fl - 1] / (self.ulen - fl + 1))   # (:233-236)
            return ok, pro, 0.0
        # ---- bias modes (:195-198, :218-236) ----
        # the read's role: forward on + / reverse on - is the 5' mate, whose 5' end weight applies unless the
        # mode is end3; otherwise the 3' mate, whose 3' end weight applies unless the mode is end5
        is_rev = np.array([r.is_reverse for r in reads], dtype=bool)
        five = ~is_rev if self.strand_plus else is_rev
        end5 = np.where(self.strand_plus, o_a, o_b)                      # transcript offset of the 5' / 3' end
        end3 = np.where(self.strand_plus, o_b, o_a)
        w = prob.copy()
        if bias_mode != "end3":
            sel = ok & five
            w[sel] = w[sel] * self.bias5[end5[sel]]
        if bias_mode != "end5":
            sel = ok & ~five
            w[sel] = w[sel] * self.bias3[end3[sel]]
        # bias-weighted number of positions of every fragment length in use, summed in position order
        biasLen = np.zeros(self.ulen)
        efflen_bias = 0.0
        for i in np.nonzero(probs)[0] + 1:
            j = np.arange(self.ulen - i + 1)
            p5, p3 = (j, j + i - 1) if self.strand_plus else (j + i - 1, j)
            if bias_mode == "end5":
                terms = self.bias5[p5]
            elif bias_mode == "end3":
                terms = self.bias3[p3]
            else:
                terms = self.bias5[p5] * self.bias3[p3]
            biasLen[i - 1] = np.cumsum(terms)[-1] if len(terms) else 0.0
            efflen_bias += probs[i - 1] * biasLen[i - 1]
        with np.errstate(all="ignore"):
            pro[ok] = w[ok] * (probs[fl - 1] / biasLen[fl - 1])
        return ok, pro, float(efflen_bias)


def gene_time_matrices(gene, bams, flen_mean=None, flen_std=None, bias_mode="unif", fasta=None, bias=None,
                       with_efflen_bias=False):
    """(Rmat (N,C) bool, prob (N,C) float with NaN, efflen (C,), n_counted) of one gene at one
    time point; `bams` = the replicate files of that time point (run_utils.py:62-92).  prob is proU, or
    proB when bias_mode is end5 / end3 / both (then `fasta` and `bias` = a FastaFile and the BiasFile of
    this time point are needed).  efflen is efflen_unif in every mode, as get_psi uses it (run_utils.py:86);
    with_efflen_bias appends the bias-weighted effective lengths (tran_utils.py:312-322)."""
    if bias_mode not in ("unif", "end5", "end3", "both"):
        raise ValueError("bias_mode must be unif, end5, end3 or both")
    if bias_mode != "unif" and (fasta is None or bias is None):
        raise ValueError("bias mode %s needs a reference FASTA and a bias file" % bias_mode)
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
    pro_all = np.ones((N, C), dtype=float)
    efflen = np.zeros(C)
    efflen_bias = np.zeros(C)
    for i, tr in enumerate(gene.trans):
        u = _Units(tr)
        if bias_mode != "unif":
            u.set_sequence(fasta, tr.chrom)
            u.set_bias(bias)
        for lo, group in ((0, r1u), (len(r1u), r2u)):
            if not group:
                continue
            ok, pro, eb = u.read_columns(group, flen_mean, flen_std, bias_mode)
            Rmat[lo:lo + len(group), i] = ok
            pro_all[lo:lo + len(group), i] = pro
            efflen_bias[i] += ok.sum() * eb
        efflen[i] = u.ulen          # every length class counts once: efflen_unif == transcript length (:218,:314-322)
        n_ok = Rmat[:, i].sum()
        efflen_bias[i] = efflen_bias[i] / n_ok if n_ok > 0 else u.ulen
    count = int(np.sum(Rmat.sum(axis=1) > 0))
    if with_efflen_bias:
        return Rmat, pro_all, efflen, count, efflen_bias
    return Rmat, pro_all, efflen, count


class TranSplice:
    """The reference's per-gene, per-time-point object (tran_utils.py:239-322) on top of the vectorised
    functions above: set_reads (one call per replicate file) / set_sequence / set_bias / get_ready, then
    Rmat, proU or proB, efflen_unif, efflen_bias.  mate_mode "single" only (what the CLI uses)."""

    def __init__(self, gene):
        self.gene = gene
        self._bams, self._fasta, self._bias = [], None, None

    def set_reads(self, samfile):
        self._bams.append(samfile)

    def set_sequence(self, fastafile):
        self._fasta = fastafile

    def set_bias(self, biasfile, mode="seq"):
        if mode not in ("seq", "sequence"):
            raise NotImplementedError("only the sequence model is used by get_psi (run_utils.py:73)")
        self._bias = biasfile

    def get_ready(self, bias_mode="unif", flen_mean=None, flen_std=None, mate_mode="single", auto_min=200):
        if mate_mode != "single":
            raise NotImplementedError('mate_mode "single" is what diceseq.py:152 hard-codes')
        R, P, L, cnt, Lb = gene_time_matrices(self.gene, self._bams, flen_mean, flen_std, bias_mode, self._fasta,
                                              self._bias, with_efflen_bias=True)
        self.Rmat, self.efflen_unif, self.efflen_bias, self.count = R, L, Lb, cnt
        self.proU = P if bias_mode == "unif" else None
        self.proB = P if bias_mode != "unif" else None
