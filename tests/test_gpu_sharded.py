"""The north-star split on real hardware: ONE gene set, LPT-partitioned over the ranks, every rank
samples its share with the real BatchSampler, rank 0 gathers the summaries (diceseq/diceseq.py:226-235 is
what this replaces).  Two gloo ranks share cuda:0 here (the test box has one GPU; bench.py / the CLI use
one GPU per rank with NCCL).  Streams are keyed by the global gene id, so the gathered summaries must be
bit-identical to the single-process run."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

KW = dict(M=1500, initial=600, gap=100, seed=11)
N_GENES = 24


def _genes():
    from diceseq_b200 import synth
    shapes = [(2, 300), (3, 1000), (4, 64), (8, 500), (5, 999), (6, 120)]
    return [synth.gene_W("timeseries", 700 + g, reads=shapes[g % 6][1], T=4, C=shapes[g % 6][0]) for g in range(N_GENES)]


def _factory():
    from diceseq_b200.batch import BatchSampler
    s = BatchSampler(0)
    s.ctx.set_trace_budget(2 << 30)
    return s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diceseq_b200 import partition
    res = partition.sample_sharded(_genes(), np.arange(700, 700 + N_GENES), _factory, rank, world, device=0,
                                   _batch_results=True, **KW)
    if rank == 0:
        q.put({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_way_partition_gives_bit_identical_summaries():
    from diceseq_b200 import partition
    single = partition.sample_sharded(_genes(), np.arange(700, 700 + N_GENES), _factory, 0, 1, _batch_results=True, **KW)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert list(got["gene_ids"]) == list(range(700, 700 + N_GENES))
    assert (np.asarray(single["m"]) > 0).all()
    for k, v in single.items():
        assert np.array_equal(np.asarray(got[k]), np.asarray(v)), k
