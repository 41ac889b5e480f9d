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


def sample_sharded(genes, gene_ids, sampler_factory, rank, world, **run_kw):
    """Partition `genes` (list of (W_list, len_list)) with LPT over `world` ranks, sample this
    rank's share with `sampler_factory().run`, gather on rank 0 in gene order."""
    from .packer import pack_genes
    n_iso = [len(g[1][0]) for g in genes]
    reads = [sum(w.shape[0] for w in g[0]) for g in genes]
    bins = lpt_partition(gene_cost(n_iso, reads), world)
    mine = bins[rank]
    payload = []
    if len(mine):
        packed = pack_genes([genes[i] for i in mine])
        sampler = sampler_factory()
        res = sampler.run(packed, gene_ids=np.asarray(gene_ids)[mine], **run_kw)
        payload = [res[k] for k in range(len(mine))]
    return gather_results(np.asarray(gene_ids)[mine], payload)
