"""Per-gene driver pieces of the `diceseq` drop-in: what
`get_psi` does around the sampler (`diceseq/utils/run_utils.py:49-170`), split so that the
matrices of MANY genes are built on the host first and sampled in one GPU batch.

  gene_inputs()        run_utils.py:60-93   R / len / prob lists and fragment counts of a gene
  format_dice_lines()  run_utils.py:134-150 one `.dice` row per transcript
  format_sample_lines  run_utils.py:152-165 the `.sample.gz` block of a gene
  sample_gene_theta2() run_utils.py:106-131 one gene with theta2 sampled (`--thetas x learn`)
"""
from __future__ import annotations

import numpy as np

from .bam import load_samfile
from .read_model import gene_time_matrices


_FASTA_CACHE, _BIAS_CACHE = {}, {}


def _fasta(path):
    from .bias import FastaFile
    if path not in _FASTA_CACHE:
        _FASTA_CACHE[path] = FastaFile(path)
    return _FASTA_CACHE[path]


def _bias(path):
    from .bias import BiasFile
    if path not in _BIAS_CACHE:
        _BIAS_CACHE[path] = BiasFile(path)
    return _BIAS_CACHE[path]


def gene_inputs(gene, sam_list, FLmean=None, FLstd=None, bias_mode="unif", ref_file=None, bias_file=None):
    """R_all, len_iso_all, prob_iso_all (lists over time points) and count (T,) of one gene.
    `sam_list[t]` lists the replicate BAM files of time point t.  Single-mate mode (diceseq.py:152).
    bias_mode unif -> proU; end5 / end3 / both -> proB with the sequence weights of `bias_file` (a list: one
    file for all time points or one per time point) on the genome `ref_file`; the fragment-length model
    takes the bias file's mean / std once a file provides them, and keeps them for the later time points
    (run_utils.py:66-77).  The effective length stays efflen_unif in every mode (run_utils.py:86)."""
    R_all, len_all, prob_all, count = [], [], [], []
    for i, files in enumerate(sam_list):
        bias = None
        if bias_mode != "unif":
            if len(bias_file) == len(sam_list):
                bias = _bias(bias_file[i])
            elif len(bias_file) == 1:
                bias = _bias(bias_file[0])
            if bias is not None:
                if FLmean is None and bias.flen_mean != 0:
                    FLmean = bias.flen_mean
                if FLstd is None and bias.flen_std != 0:
                    FLstd = bias.flen_std
        if bias_mode != "unif" and bias is None:
            raise ValueError("--bias: give one bias file, or one per time point")
        R, P, L, n = gene_time_matrices(gene, [load_samfile(f) for f in files], FLmean, FLstd, bias_mode,
                                        _fasta(ref_file) if bias_mode != "unif" else None, bias)
        R_all.append(R)
        len_all.append(L)
        prob_all.append(P)
        count.append(n)
    return R_all, len_all, prob_all, np.array(count)


def format_dice_lines(gene, X, count, TOTAL_READ, psi_mean, psi_95, lik_marg):
    """Rows `tran_id gene_id logLik transLen (FPKM ratio ratio_lo ratio_hi) x T` of one gene."""
    tran_len = np.array([tr.tranL for tr in gene.trans])
    out = ""
    for c, tr in enumerate(gene.trans):
        line = "%s\t%s\t%.1e\t%d" % (tr.tranID, gene.geneID, lik_marg, tr.tranL)
        for t in range(len(X)):
            fsi = (tran_len * psi_mean[:, t]) / np.sum(tran_len * psi_mean[:, t])
            fpkm = count[t] * fsi[c] * 10 ** 9 / tran_len[c] / TOTAL_READ[t]
            line += "\t%.2e\t%.2e\t%.2e\t%.2e" % (fpkm, psi_mean[c, t], psi_95[c][t, 1], psi_95[c][t, 0])
        out += line + "\n"
    return out


def format_sample_lines(gene, theta2, Y_tail, n_iso):
    """`@gene|transcripts|theta2...` then one line per saved sample: `y_c1t1,y_c2t1;y_c1t2,...`.
    Y_tail: (n_samples, C, T) rows Y_all[int(m/2) : int(m/2)+n] (run_utils.py:156-165)."""
    info = gene.get_gene_info()
    th = np.broadcast_to(np.asarray(theta2, dtype=float), (n_iso - 1,))    # fixed value, or the trace means
    head = ",".join("%.2e" % v for v in th)
    out = "@%s|%s|%s\n" % (info[0], info[-1], head)
    for s in range(Y_tail.shape[0]):
        out += ";".join(",".join("%.2e" % Y_tail[s, c, t] for c in range(Y_tail.shape[1]))
                        for t in range(Y_tail.shape[2])) + "\n"
    return out


def single_isoform_result(T):
    """A gene with one transcript never reaches the sampler: psi = 1 (run_utils.py:99-104; the
    reference then dies on undefined variables at :153 -- SURVEY.md Appendix B #7 -- here the
    rows are simply emitted)."""
    return np.ones((1, T)), [np.ones((T, 2))], 0.0


class Theta2GeneResult:
    """What the output stage needs of a gene sampled with theta2 = None."""
    __slots__ = ("m", "cnt", "psi_mean", "psi_95", "lik_marg", "theta2_mean", "Y_tail")


def sample_gene_theta2(R_all, len_all, prob_all, X, theta1, M, initial, gap, sample_num=0, device=0):
    """`get_psi`'s sampler section for `--thetas x learn` (run_utils.py:106-131): single-time-point
    pre-runs and the joint run, all with theta2 sampled, then the burn-in summaries.  The chains
    are host-driven with the read likelihood on the device (theta2_mode.py), one gene at a time,
    consuming numpy's global random state like the reference.  Raises `LinAlgError` for genes
    with more than two isoforms, like the reference's sampler does in this mode."""
    from .batch import get_CI, harmonic_mean_log
    from .model_GP import Psi_GP_MH
    T, C = len(R_all), R_all[0].shape[1]
    var = 0.1 * np.ones(C - 1)
    Ymean = np.zeros((C, T))
    acc = np.zeros(C - 1)
    th2 = None
    m1 = 0
    for j in range(T):                                           # :110-114
        _psi, Y1, th2, _pro, _lik, _cnt, m1 = Psi_GP_MH(
            R_all[j:j + 1], len_all[j:j + 1], prob_all[j:j + 1], X[j:j + 1], Ymean[:, j:j + 1], var, theta1, None,
            100, 100, 50, device=device)
        acc += np.var(Y1[int(m1 / 4):, :-1, 0], axis=0)
    var = (acc + 0.000001) / T                                   # :115
    theta2_std = np.var(th2[int(m1 / 4):, 0]) / 10.0 + 0.001    # :116 (from the LAST pre-run only)
    Psi, Y, th2, _pro, Lik, cnt, m = Psi_GP_MH(R_all, len_all, prob_all, X, Ymean, var, theta1, None, M, initial,
                                               gap, theta2_std=theta2_std, device=device)
    r = Theta2GeneResult()
    r.m, r.cnt = m, cnt
    b = int(m / 4)
    r.psi_95 = [get_CI(Psi[b:, c, :], 0.95) for c in range(C)]
    r.psi_mean = Psi[b:].mean(axis=0)
    r.lik_marg = harmonic_mean_log(Lik[b:])
    r.theta2_mean = th2[b:].mean(axis=0)
    n_keep = min(int(0.5 * m), sample_num)
    r.Y_tail = Y[int(m / 2):int(m / 2) + n_keep]
    return r
