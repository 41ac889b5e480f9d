"""Bias files and genome sequence for the bias-corrected read model (BASELINE config 4).

`BiasFile` reads the parameter file `dice-bias` writes (diceseq/utils/bias_utils.py:21-86: fragment
length line, 5 x 20 position bins, 744 sequence k-mer rows) and turns it into the per-offset weight
tables of the variable-length Markov model; `seq_bias_track` evaluates the 5' / 3' sequence weight of
EVERY locus of a transcript at once (the reference walks loci and offsets in Python,
bias_utils.py:140-157 inside tran_utils.py:38-70).  `FastaFile` is a plain in-memory FASTA reader with
pysam.FastaFile.fetch's coordinates (bias_utils.py:11-19).
"""
from __future__ import annotations

import numpy as np

# variable-length Markov chain of Roberts et al. 2011: context length of each of the 21 offsets
CHAIN_LEN = [1] * 4 + [2] * 3 + [3] * 10 + [2] * 2 + [1] * 2            # bias_utils.py:131
_BASE_CODE = {"A": 0, "T": 1, "G": 2, "C": 3}


class FastaFile:
    """get_seq(ref, start, stop): bases start..stop, 1-based inclusive (pysam fetch(ref, start-1, stop))."""

    def __init__(self, fasta_file):
        self.seqs = {}
        name, parts = None, []
        with open(fasta_file) as f:
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        self.seqs[name] = "".join(parts)
                    name, parts = line[1:].split()[0], []
                else:
                    parts.append(line.strip())
        if name is not None:
            self.seqs[name] = "".join(parts)

    def get_seq(self, qref, start, stop):
        if qref not in self.seqs:
            raise KeyError("sequence %r not in the FASTA file" % qref)
        return self.seqs[qref][max(start - 1, 0):stop]


class BiasFile:
    """Parameters of one bias file.  seq5_prob / seq3_prob: per offset j the weight of each context, in the
    order the file lists them (= the order get_seq_bias looks them up in, bias_utils.py:149-156)."""

    def __init__(self, bias_file):
        with open(bias_file) as f:
            lines = f.readlines()
        head = lines[4].split("\t")
        self.flen_mean, self.flen_std = float(head[0]), float(head[1])
        self.flen_sum1, self.flen_sum2, self.read_num = float(head[2]), float(head[3]), float(head[4])
        self.percentile = np.zeros((5, 2))
        self.pos5_prob = np.zeros((5, 20))
        self.pos3_prob = np.zeros((5, 20))
        for i in range(5, 105):
            a, b = (i - 5) // 20, (i - 5) % 20
            f_ = lines[i].split("\t")
            if b == 0:
                self.percentile[a, :] = f_[0].split("|")[0].split("-")
            with np.errstate(all="ignore"):
                self.pos5_prob[a, b] = max(0, np.float64(f_[1]) / np.float64(f_[3]))
                self.pos3_prob[a, b] = max(0, np.float64(f_[2]) / np.float64(f_[4]))
        self.kmers = {str(j): [] for j in range(21)}
        self.seq5_prob = {str(j): [] for j in range(21)}
        self.seq3_prob = {str(j): [] for j in range(21)}
        for i in range(105, 849):
            f_ = lines[i].split("\t")
            j, kmer = f_[0].split("|")
            with np.errstate(all="ignore"):
                self.seq5_prob[j].append(max(0, np.float64(f_[1]) / np.float64(f_[3])))
                self.seq3_prob[j].append(max(0, np.float64(f_[2]) / np.float64(f_[4])))
            self.kmers[j].append(kmer)
        # dense lookup per offset: code of a context (base-4 digits, first character least significant) -> weight;
        # a context the file does not list (or one holding a character outside ATGC) contributes a factor 1
        self._lut5, self._lut3 = [], []
        for j in range(21):
            L = CHAIN_LEN[j]
            l5, l3 = np.ones(4 ** L), np.ones(4 ** L)
            seen = set()
            for idx, kmer in enumerate(self.kmers[str(j)]):
                if kmer in seen or len(kmer) != L or any(ch not in _BASE_CODE for ch in kmer):
                    continue                                             # list.index finds the first occurrence
                seen.add(kmer)
                code = sum(_BASE_CODE[ch] * 4 ** p for p, ch in enumerate(kmer))
                l5[code], l3[code] = self.seq5_prob[str(j)][idx], self.seq3_prob[str(j)][idx]
            listed = np.zeros(4 ** L, bool)
            for kmer in seen:
                listed[sum(_BASE_CODE[ch] * 4 ** p for p, ch in enumerate(kmer))] = True
            self._lut5.append((l5, listed))
            self._lut3.append((l3, listed))

    def seq_bias_track(self, codes, centres, end_num, reverse):
        """Sequence weight of a 21-base window per locus: window character j is codes[centre - 8 + j]
        (`reverse` False: seq[ipos-8 : ipos+13]) or codes[centre + 8 - j] (True: seq[ipos-12 : ipos+9][::-1]);
        the product runs over j ascending like get_seq_bias (bias_utils.py:140-157).  codes: base codes 0-3 of
        the padded transcript sequence, -1 for any other character."""
        lut = self._lut5 if end_num == 5 else self._lut3
        centres = np.asarray(centres, dtype=np.int64)
        prob = np.ones(len(centres))
        step = -1 if reverse else 1
        first = centres + (8 if reverse else -8)
        for j in range(21):
            L = CHAIN_LEN[j]
            code = np.zeros(len(centres), dtype=np.int64)
            valid = np.ones(len(centres), bool)
            for p in range(L):                                           # context = window[j-L+1 .. j], first char at digit 0
                ch = codes[first + step * (j - L + 1 + p)]
                valid &= ch >= 0
                code += np.where(ch >= 0, ch, 0) * 4 ** p
            table, listed = lut[j]
            use = valid & listed[code]
            prob = np.where(use, prob * table[code], prob)
        return prob


def encode_bases(seq):
    """Base codes A=0 T=1 G=2 C=3 (upper case only, as the reference's k-mer tables), -1 otherwise."""
    raw = np.frombuffer(seq.encode("ascii", "replace"), dtype=np.uint8)
    out = np.full(len(raw), -1, dtype=np.int64)
    for ch, v in _BASE_CODE.items():
        out[raw == ord(ch)] = v
    return out
