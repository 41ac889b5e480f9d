"""Sharding genes over the GPUs of one box.

Genes are independent units of work (the reference fans them over a process pool,
diceseq/diceseq.py:226-235); there is no exchange step inside a chain.  Each rank
(one process per GPU) samples its own genes and the only communication is the final gather
of the small per-gene summaries.  Streams are keyed by the GLOBAL gene id, so the results
do not depend on the partition.
"""
from __future__ import annotations

import numpy as np


def gene_cost(n_iso, reads_per_gene):
    """Work estimate of one gene: isoforms x reads (bytes of W touched per MH step)."""
    return np.asarray(n_iso, dtype=np.int64) * np.asarray(reads_per_gene, dtype=np.int64)


def lpt_partition(costs, n_bins):
    """Longest-processing-time-first greedy partition: genes by decreasing cost, each into
    the currently lightest bin.  Returns a list of index arrays (ascending inside a bin)."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_bins, dtype=np.int64)
    bins = [[] for _ in range(n_bins)]
    for g in order:
        b = int(np.argmin(load))
        bins[b].append(int(g))
        load[b] += costs[g]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def gather_results(local_ids, local_payload, group=None):
    """Final gather: every rank contributes (gene ids, list of per-gene payloads); rank 0
    gets them back in global gene order (others get None).  Uses torch.distributed
    (NCCL on the GPU box, gloo in CPU tests); a single process just sorts."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = np.argsort(np.asarray(local_ids))
        return [local_payload[i] for i in order]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [None] * world if rank == 0 else None
    dist.gather_object((np.asarray(local_ids).tolist(), local_payload), out, dst=0, group=group)
    if rank != 0:
        return None
    ids, payload = [], []
    for part_ids, part_payload in out:
        ids += part_ids
        payload += part_payload
    order = np.argsort(np.asarray(ids))
    return [payload[i] for i in order]


def gather_arrays(arrays, device=None, group=None):
    """Variable-length gather of small numeric 1-D arrays: every rank passes the same number of arrays
    (lengths may differ between ranks); rank 0 gets, per array, the concatenation over ranks in rank
    order, the others None.  Two collectives: an all_gather of the lengths and one gather of a padded
    float64 buffer (integers up to 2^53 survive the round trip).  CUDA tensors for NCCL (`device`),
    host tensors for gloo; a single process returns its own arrays."""
    import torch
    import torch.distributed as dist
    arrays = [np.ascontiguousarray(a).ravel() for a in arrays]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arrays
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", device) if device is not None and dist.get_backend(group) == "nccl" else torch.device("cpu")
    lens = torch.tensor([a.size for a in arrays], dtype=torch.int64, device=dev)
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    all_lens = torch.stack(all_lens).cpu().numpy()                      # (world, n_arrays)
    width = int(all_lens.sum(axis=1).max())
    buf = np.zeros(max(width, 1))
    o = 0
    for a in arrays:
        buf[o:o + a.size] = a
        o += a.size
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0, group=group)
    if rank != 0:
        return None
    parts = [p.cpu().numpy() for p in parts]
    out = []
    for k, a in enumerate(arrays):
        cols = []
        for r in range(world):
            o = int(all_lens[r, :k].sum())
            cols.append(parts[r][o:o + int(all_lens[r, k])].astype(a.dtype))
        out.append(np.concatenate(cols))
    return out


def gather_batch(res, gene_ids, device=None, group=None):
    """The final gather of the per-gene summaries of a rank's BatchResult (the only communication of a
    sharded run; diceseq/diceseq.py:226-235 collects its pool's results the same way).  Rank 0 gets a dict
    of arrays in GLOBAL gene-id order -- gene_ids, n_iso, m, cnt, state, flags, lik_marg, off (G + 1
    offsets into psi_mean; 2 x into psi_ci), psi_mean, psi_ci -- the other ranks None."""
    n_iso = np.asarray(res.n_iso, dtype=np.int64)
    per_gene = [np.asarray(gene_ids, dtype=np.int64), n_iso, res.m, res.cnt, res.state, res.flags,
                res.lik_marg if res.lik_marg is not None else np.zeros(len(n_iso))]
    flat = [res.psi_mean if res.psi_mean is not None else np.zeros(0),
            res.psi_ci if res.psi_ci is not None else np.zeros(0)]
    got = gather_arrays(per_gene + flat, device=device, group=group)
    if got is None:
        return None
    ids, n_iso, m, cnt, state, flags, lik, psi_mean, psi_ci = got
    T = res.T
    size = n_iso * T
    off_in = np.concatenate([[0], np.cumsum(size)])
    order = np.argsort(ids, kind="stable")
    out = {"gene_ids": ids[order], "n_iso": n_iso[order], "m": m[order], "cnt": cnt[order], "state": state[order],
           "flags": flags[order], "lik_marg": lik[order], "T": T}
    out["off"] = np.concatenate([[0], np.cumsum(size[order])])
    if psi_mean.size:
        if np.array_equal(order, np.arange(len(order))):       # already in gene order (one rank, sorted ids)
            out["psi_mean"], out["psi_ci"] = psi_mean, psi_ci
        else:
            # element i of gene order[k] moves from off_in[order[k]] + i to out["off"][k] + i: one fancy index,
            # no per-gene Python loop (10k genes: 40 ms -> 1 ms)
            so = size[order]
            src = np.repeat(off_in[:-1][order] - out["off"][:-1], so) + np.arange(int(so.sum()))
            out["psi_mean"] = psi_mean[src]
            out["psi_ci"] = psi_ci.reshape(-1, 2)[src].ravel() if psi_ci.size == 2 * psi_mean.size else psi_ci
    return out


class _EmptyBatch:
    """The BatchResult of a rank that got no genes."""

    def __init__(self, T):
        z = np.zeros(0, dtype=np.int32)
        self.T, self.n_iso, self.m, self.cnt, self.state, self.flags = T, z, z, z, z, z
        self.lik_marg = self.psi_mean = self.psi_ci = np.zeros(0)


def sample_sharded(genes, gene_ids, sampler_factory, rank, world, device=None, **run_kw):
    """Partition `genes` (list of (W_list, len_list)) with LPT over `world` ranks, sample this rank's
    share with `sampler_factory().run`, gather on rank 0 in gene order: a dict of arrays (`gather_batch`)
    when the sampler returns a BatchResult, else the list of per-gene payloads (`gather_results`)."""
    from .packer import pack_genes
    n_iso = [len(g[1][0]) for g in genes]
    reads = [sum(w.shape[0] for w in g[0]) for g in genes]
    bins = lpt_partition(gene_cost(n_iso, reads), world)
    mine = bins[rank]
    ids = np.asarray(gene_ids)[mine]
    res = None
    if len(mine):
        packed = pack_genes([genes[i] for i in mine])
        sampler = sampler_factory()
        res = sampler.run(packed, gene_ids=ids, **{k: v for k, v in run_kw.items() if k != "_batch_results"})
    if res is None and run_kw.get("_batch_results"):
        res = _EmptyBatch(len(genes[0][0]))                    # a rank without genes still joins the gather
    run_kw.pop("_batch_results", None)
    if res is not None and hasattr(res, "psi_mean"):
        return gather_batch(res, ids, device=device)
    payload = [res[k] for k in range(len(mine))] if res is not None else []
    return gather_results(ids, payload)
