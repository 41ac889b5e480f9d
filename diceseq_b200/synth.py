"""Seeded synthetic genes of the shapes BASELINE.json names.

The generator follows what the reference's read model produces for real data
(diceseq/utils/tran_utils.py:148-237, 277-322): a boolean compatibility matrix R, a
per-read-per-isoform probability `prob = q * phi(flen) / (L_k - flen + 1)` with NaN where
the read is incompatible, and effective lengths `len = L`.  One RandomState per gene,
keyed by (config seed, gene index), so any partition of the genes over GPUs sees the
same data.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (config id, T, isoform rule, reads rule, bias weights)
    "static": dict(cid=2, T=1, G=10000),
    "timeseries": dict(cid=3, T=8, G=10000),
    "bias": dict(cid=4, T=8, G=20000),
    "human": dict(cid=5, T=4, G=60000),
}


def default_theta2(X):
    """diceseq/diceseq.py:169: theta2 covers a third of the duration."""
    return ((np.max(X) - np.min(X) + 0.1) / 3.0) ** 2


def _gene_rng(cid, g):
    return np.random.RandomState((1000 + cid) * 1000003 + int(g) & 0x7FFFFFFF)


def _n_isoforms(cid, rng):
    if cid == 2:
        return 3
    if cid in (3, 4):
        return int(rng.randint(2, 9))
    return int(min(20, 2 + rng.geometric(0.25) - 1))


def _n_reads(cid, rng, T, reads):
    if cid == 5:
        n = np.rint(rng.lognormal(np.log(1000.0), 2.0, size=T))
        return np.clip(n, 10, 1000000).astype(np.int64)
    return np.full(T, reads, dtype=np.int64)


def _gene_header(config, g, reads, T, C):
    """The draws that fix a gene's shape: (rng, cid, T, C, L, psi (C, T), reads per time point)."""
    spec = CONFIGS[config]
    cid = spec["cid"]
    T = spec["T"] if T is None else T
    rng = _gene_rng(cid, g)
    C = _n_isoforms(cid, rng) if C is None else C
    L = rng.randint(500, 4001, size=C).astype(np.float64)
    X = np.arange(T)
    if T == 1:
        psi = rng.dirichlet(np.ones(C))[:, None]
    else:
        d = X[:, None] - X[None, :]
        K = 3.0 * np.exp(-0.5 * d * d / default_theta2(X)) + 1e-9 * np.eye(T)
        y = rng.multivariate_normal(np.zeros(T), K, size=C)
        y[-1] = 0.0
        e = np.exp(y - y.max(axis=0))
        psi = e / e.sum(axis=0)
    n_t = _n_reads(cid, rng, T, reads)
    return rng, cid, T, C, L, psi, n_t


def gene_shape(config, g, reads=1000, T=None, C=None):
    """(isoforms, reads per time point) of gene g without generating its matrices: what the
    read-count-balanced partition over GPUs needs (diceseq_b200/partition.py)."""
    _rng, _cid, _T, C, _L, _psi, n_t = _gene_header(config, g, reads, T, C)
    return C, n_t


def gene_matrices(config, g, reads=1000, T=None, C=None):
    """(R_list, len_list, prob_list) of gene number g of a named config, in the exact form
    `Psi_GP_MH` takes them (lists over time points)."""
    rng, cid, T, C, L, psi, n_t = _gene_header(config, g, reads, T, C)
    R_list, len_list, prob_list = [], [], []
    for t in range(T):
        n = int(n_t[t])
        fsi = L * psi[:, t]
        fsi /= fsi.sum()
        src = rng.choice(C, size=n, p=fsi)
        R = rng.rand(n, C) < 0.5
        R[np.arange(n), src] = True
        flen = np.clip(rng.normal(200.0, 30.0, size=n), 75.0, L.min())
        phi = np.exp(-0.5 * ((flen - 200.0) / 30.0) ** 2) / (30.0 * np.sqrt(2 * np.pi))
        mapq_ok = rng.rand(n) < 0.985
        q = np.where(mapq_ok, 1.0 - 10.0 ** (-25.5), 1.0 - 10.0 ** (-0.1))
        prob = (q * phi)[:, None] / (L[None, :] - flen[:, None] + 1.0)
        if cid == 4:
            prob = prob * rng.lognormal(0.0, 0.5, size=n)[:, None]
        prob[~R] = np.nan
        R_list.append(R)
        len_list.append(L.copy())
        prob_list.append(prob)
    return R_list, len_list, prob_list


def gene_W(config, g, reads=1000, T=None, C=None):
    """(W_list, len_list) with W_t = R_t * prob_t after the reference's NaN cleaning."""
    R, lens, prob = gene_matrices(config, g, reads, T, C)
    W = []
    for t in range(len(R)):
        p = np.where(np.isnan(prob[t]), 0.0, prob[t])
        W.append(R[t] * p)
    return W, lens
