"""Per-gene driver pieces of the `diceseq` drop-in: what
`get_psi` does around the sampler (`diceseq/utils/run_utils.py:49-170`), split so that the
matrices of MANY genes are built on the host first and sampled in one GPU batch.

  gene_inputs()        run_utils.py:60-93   R / len / prob lists and fragment counts of a gene
  format_dice_lines()  run_utils.py:134-150 one `.dice` row per transcript
  format_sample_lines  run_utils.py:152-165 the `.sample.gz` block of a gene
"""
from __future__ import annotations

import numpy as np

from .bam import load_samfile
from .read_model import gene_time_matrices


def gene_inputs(gene, sam_list, FLmean=None, FLstd=None):
    """R_all, len_iso_all, prob_iso_all (lists over time points) and count (T,) of one gene.
    `sam_list[t]` lists the replicate BAM files of time point t.  Uniform model, single-mate
    mode: the two settings the reference CLI hard-codes (diceseq.py:150-153)."""
    R_all, len_all, prob_all, count = [], [], [], []
    for files in sam_list:
        R, P, L, n = gene_time_matrices(gene, [load_samfile(f) for f in files], FLmean, FLstd)
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
    head = ",".join(["%.2e" % theta2] * (n_iso - 1))
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
