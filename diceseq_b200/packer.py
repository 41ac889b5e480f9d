"""Host-side packing of per-gene read-by-isoform matrices for libdicegp.

`clean_inputs` is the input cleaning Psi_GP_MH does first (diceseq/models/model_GP.py:250-264,
done on the host because it mutates the caller's arrays); `pack_genes` lays the cleaned
W_t = R_t * prob_t (model_GP.py:338) out the way `dicegp_upload_genes` wants it
(include/dicegp.h): per gene, per time point, an isoform-major (C x pad2(N_t)) block.
"""
from __future__ import annotations

import numpy as np


def clean_inputs(R_mat, len_isos, prob_isos):
    """NaN lengths -> 0 (and that isoform's column zeroed), NaN probabilities -> 0, then
    keep only reads that are compatible with some isoform and have some probability mass.
    Works IN PLACE on the arrays and rebinds the list entries, like the reference does."""
    for t, lens in enumerate(len_isos):
        nan_len = np.isnan(lens) if lens.dtype.kind == "f" else np.zeros(lens.shape, bool)
        lens[nan_len] = 0.0
        prob_isos[t][:, nan_len] = 0.0
        R_mat[t][:, nan_len] = False
        if R_mat[t].dtype.kind == "f":
            R_mat[t][np.isnan(R_mat[t])] = False
        if prob_isos[t].dtype.kind == "f":
            prob_isos[t][np.isnan(prob_isos[t])] = 0.0
        alive = (R_mat[t].sum(axis=1) > 0) & (prob_isos[t].sum(axis=1) > 0)
        R_mat[t] = R_mat[t][alive, :]
        prob_isos[t] = prob_isos[t][alive, :]


class PackedGenes:
    """Flat arrays for dicegp_upload_genes."""

    def __init__(self, G, T, n_iso, n_reads, W, len_iso):
        self.G, self.T = int(G), int(T)
        self.n_iso = np.ascontiguousarray(n_iso, dtype=np.int32)
        self.n_reads = np.ascontiguousarray(n_reads, dtype=np.int32)
        self.W = W
        self.len_iso = len_iso

    @property
    def w_bytes(self):
        return self.W.nbytes

    def reads_per_gene(self):
        return self.n_reads.reshape(self.G, self.T).sum(axis=1)

    def subset(self, idx):
        """PackedGenes of the genes listed in idx (used to shard genes over GPUs)."""
        idx = np.asarray(idx, dtype=np.int64)
        npad = self.n_reads.reshape(self.G, self.T).astype(np.int64)
        npad = npad + (npad & 1)
        wsize = self.n_iso.astype(np.int64) * npad.sum(axis=1)
        woff = np.concatenate([[0], np.cumsum(wsize)])
        lsize = self.n_iso.astype(np.int64) * self.T
        loff = np.concatenate([[0], np.cumsum(lsize)])
        W = np.concatenate([self.W[woff[g]:woff[g + 1]] for g in idx]) if len(idx) else np.zeros(0)
        L = np.concatenate([self.len_iso[loff[g]:loff[g + 1]] for g in idx]) if len(idx) else np.zeros(0)
        return PackedGenes(len(idx), self.T, self.n_iso[idx],
                           self.n_reads.reshape(self.G, self.T)[idx].ravel(), W, L)


def pack_gene(W_list, len_list):
    """One gene in the device layout: (flat W, reads per time point, flat len).  Cheap to ship between
    processes: the matrix workers (bench.py, the CLI's -p pool) pack their own genes so that the process that
    feeds the GPU only concatenates (`pack_packed`)."""
    T = len(W_list)
    if len(len_list) != T:
        raise ValueError("every gene must have the same number of time points")
    C = len(len_list[0])
    n = np.empty(T, dtype=np.int32)
    for t in range(T):
        if W_list[t].ndim != 2 or W_list[t].shape[1] != C:
            raise ValueError("W[t] must be (N_t, C)")
        n[t] = W_list[t].shape[0]
    pad = n.astype(np.int64) + (n & 1)
    flat = np.zeros(int(C * pad.sum()))
    o = 0
    for t in range(T):
        blk = flat[o:o + C * int(pad[t])].reshape(C, int(pad[t]))
        blk[:, :n[t]] = W_list[t].T
        o += C * int(pad[t])
    return flat, n, np.concatenate([np.asarray(l, dtype=np.float64) for l in len_list])


def pack_packed(items, pinned=False):
    """PackedGenes from per-gene `pack_gene` outputs (all with the same number of time points)."""
    G = len(items)
    T = len(items[0][1])
    n_iso = np.array([len(it[2]) // T for it in items], dtype=np.int32)
    n_reads = np.empty((G, T), dtype=np.int32)
    for g, it in enumerate(items):
        if len(it[1]) != T:
            raise ValueError("every gene must have the same number of time points")
        n_reads[g] = it[1]
    W = _alloc(int(sum(it[0].size for it in items)), pinned)
    L = _alloc(int(sum(it[2].size for it in items)), pinned)
    wo = lo = 0
    for flat, _n, lens in items:
        W[wo:wo + flat.size] = flat
        wo += flat.size
        L[lo:lo + lens.size] = lens
        lo += lens.size
    return PackedGenes(G, T, n_iso, n_reads.ravel(), W, L)


def pack_genes(genes, pinned=False):
    """genes: list of (W_list, len_list) with W_list[t] an (N_t, C) float array and
    len_list[t] a (C,) array; all genes share T = len(W_list)."""
    return pack_packed([pack_gene(W_list, len_list) for W_list, len_list in genes], pinned)


def _alloc(n, pinned):
    if pinned:
        import torch  # plumbing only: pinned host memory
        if torch.cuda.is_available():
            return torch.empty(max(n, 1), dtype=torch.float64).pin_memory().numpy()[:n]
    return np.empty(n, dtype=np.float64)
