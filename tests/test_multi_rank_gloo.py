"""N > 1 path on CPU: world_size-2 gloo processes partition the genes (LPT), each "samples"
its share with a stand-in sampler keyed by the global gene id, rank 0 gathers.  The gathered
result must equal the single-process result (genes are independent; streams are keyed by
gene id, so the partition cannot change anything)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diceseq_b200 import partition  # noqa: E402


class FakeSampler:
    """Deterministic per-gene output from (gene id, W): what a partition-invariant sampler returns."""

    def run(self, packed, gene_ids=None, **kw):
        out, off = [], 0
        npad = packed.n_reads.reshape(packed.G, packed.T).astype(np.int64)
        npad = npad + (npad & 1)
        for g in range(packed.G):
            n = int(packed.n_iso[g] * npad[g].sum())
            out.append((int(gene_ids[g]), float(packed.W[off:off + n].sum()), int(packed.n_iso[g])))
            off += n
        return out


def _genes(n):
    rng = np.random.RandomState(3)
    genes = []
    for g in range(n):
        C = int(rng.randint(2, 6))
        T = 3
        W = [rng.rand(int(rng.randint(5, 60)), C) for _ in range(T)]
        genes.append((W, [np.arange(1, C + 1, dtype=float)] * T))
    return genes


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genes = _genes(23)
    ids = np.arange(100, 123)
    res = partition.sample_sharded(genes, ids, FakeSampler, rank, world)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_partition_balances_and_covers():
    costs = np.array([100, 1, 1, 1, 50, 50, 2, 2, 97, 3])
    bins = partition.lpt_partition(costs, 3)
    allidx = np.sort(np.concatenate(bins))
    assert np.array_equal(allidx, np.arange(len(costs)))
    loads = [costs[b].sum() for b in bins]
    assert max(loads) - min(loads) <= 5
    one = partition.lpt_partition(costs, 1)
    assert np.array_equal(one[0], np.arange(len(costs)))


def test_world2_gloo_gather_equals_single_process():
    genes = _genes(23)
    ids = np.arange(100, 123)
    single = partition.sample_sharded(genes, ids, FakeSampler, 0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == single
    assert [g[0] for g in got] == list(ids)


class FakeBatch:
    """BatchResult-shaped output keyed by the global gene id (what gather_batch ships)."""

    def __init__(self, packed, gene_ids):
        G, T = packed.G, packed.T
        self.T, self.n_iso = T, packed.n_iso
        ids = np.asarray(gene_ids)
        self.m = (ids * 7 % 1000).astype(np.int32)
        self.cnt = (ids * 3 % 100).astype(np.int32)
        self.state = (ids % 3).astype(np.int32)
        self.flags = np.zeros(G, dtype=np.int32)
        self.lik_marg = -ids.astype(float) / 3
        self.psi_mean = np.concatenate([np.full(int(c) * T, float(i)) + np.arange(int(c) * T) for i, c in zip(ids, packed.n_iso)])
        self.psi_ci = np.repeat(self.psi_mean, 2) + 0.5


class FakeBatchSampler:
    def run(self, packed, gene_ids=None, **kw):
        return FakeBatch(packed, gene_ids)


def _worker_batch(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = partition.sample_sharded(_genes(23), np.arange(100, 123), FakeBatchSampler, rank, world, _batch_results=True)
    if rank == 0:
        q.put({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_gather_batch_world2_equals_single_process():
    """The tensor gather of the summaries (gather_batch: one all_gather of lengths + one padded gather)
    returns, on rank 0, exactly what a single process computes, in global gene order."""
    single = partition.sample_sharded(_genes(23), np.arange(100, 123), FakeBatchSampler, 0, 1, _batch_results=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + 7
    procs = [ctx.Process(target=_worker_batch, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(got) == set(single)
    for k, v in single.items():
        assert np.array_equal(np.asarray(got[k]), np.asarray(v)), k
    assert list(got["gene_ids"]) == list(range(100, 123))
