"""Batched sampler: what `get_psi` does around Psi_GP_MH (diceseq/utils/run_utils.py:106-131),
for many genes at once on one GPU.

Per gene the reference runs T single-time-point pre-runs (M=100, var=0.1) to set the jump
variance, then the joint chain, then burn-in summaries.  Here all genes' pre-runs are one
set of G*T chains, the jump variances are reduced on the device
(`dicegp_main_chains_from_prerun`), the G main chains run to Geweke convergence and the
burn-in summaries are reduced on the device too (`dicegp_posterior_summary`).

Random streams: `stream="device"` (default) uses the library's counter-based Philox
stream keyed by (seed, global gene id, chain kind, step) -- results do not depend on how
genes are spread over GPUs.  `stream="host"` generates the same kind of increments on
the host (numpy Philox) and ships them in `gap`-sized chunks, retiring converged chains
between chunks; it exists because the reference generates its stream on the host and is
the path the MT19937 replay in `Psi_GP_MH` uses.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .model_GP import GP_K, check_steps, proposal_factor

_LOG_2PI = np.log(2 * np.pi)


def get_CI(data, percent=0.95):
    """[upper, lower] order statistics per column with the reference's index rule
    (run_utils.py:14-22; `sorted[-0]` is the minimum when there are fewer than 40 rows)."""
    if data.ndim == 0:
        data = data.reshape(-1, 1)
    srt = np.sort(data, axis=0)
    k = int(data.shape[0] * (1 - percent) / 2)
    return np.stack([srt[-k], srt[k]], axis=1)


def harmonic_mean_log(data):
    """log of the harmonic mean of exp(data) (run_utils.py:24-30)."""
    lo = np.min(data, axis=0)
    return np.log(len(data)) + lo - np.log(np.sum(np.exp(lo - data), axis=0))


class GeneResult:
    """What get_psi keeps of one gene's run (run_utils.py:123-131)."""
    __slots__ = ("m", "cnt", "state", "flags", "psi_mean", "psi_95", "lik_marg", "var", "trace")

    def __init__(self):
        self.psi_mean = self.psi_95 = self.lik_marg = self.trace = None


class BatchResult:
    """Flat per-batch arrays; indexing gives a GeneResult."""

    def __init__(self, n_iso, T, status, var):
        self.n_iso, self.T = n_iso, T
        self.m, self.cnt = status["m"], status["cnt"]
        self.state, self.flags, self.n_rows = status["state"], status["flags"], status["n_rows"]
        self.var = var
        self.psi_mean = self.psi_ci = self.lik_marg = self.off = None
        self.traces = None

    def __len__(self):
        return len(self.n_iso)

    def __getitem__(self, g):
        if isinstance(g, slice):
            return [self[i] for i in range(*g.indices(len(self)))]
        if g < 0:
            g += len(self)
        if g < 0 or g >= len(self):
            raise IndexError(g)
        r = GeneResult()
        C = int(self.n_iso[g])
        r.m, r.cnt, r.state, r.flags = int(self.m[g]), int(self.cnt[g]), int(self.state[g]), int(self.flags[g])
        r.var = self.var[g, :C - 1].copy()
        if self.psi_mean is not None:
            a, b = int(self.off[g]), int(self.off[g + 1])
            r.psi_mean = self.psi_mean[a:b].reshape(C, self.T)
            ci = self.psi_ci[2 * a:2 * b].reshape(C, self.T, 2)
            r.psi_95 = [ci[c] for c in range(C)]
            r.lik_marg = float(self.lik_marg[g])
        if self.traces is not None:
            r.trace = self.traces[g]
        return r

    def __iter__(self):
        return (self[g] for g in range(len(self)))

    def d2h_bytes(self):
        n = sum(a.nbytes for a in (self.m, self.cnt, self.state, self.flags, self.n_rows))
        if self.psi_mean is not None:
            n += self.psi_mean.nbytes + self.psi_ci.nbytes + self.lik_marg.nbytes
        return n + self.var.nbytes


class BatchSampler:
    """One GPU.  `run` = `upload` + `sample`."""

    def __init__(self, device=0):
        self.ctx = _lib.Context(device)
        self.stats = {}
        self.packed = None

    def close(self):
        self.ctx.close()

    # -- host stream ---------------------------------------------------------------------
    def _run_host(self, rng, js_rows, pf, T_chain, ini, gap):
        """Advance all unfinished chains to the end with host-generated streams.
        js_rows[i]: jump scales (C-1,) of chain i; pf: (Tc,Tc) proposal factor."""
        ctx = self.ctx
        n = len(ctx.chain_shapes)
        active = np.arange(n)
        M = ctx.chain_shapes[0][2]
        start = 0
        for end in check_steps(M, ini, gap):
            while start < end and active.size:
                S = min(end - start, 512)
                blocks, offs, o = [], [], 0
                for i in active:
                    Cm = len(js_rows[i])
                    z = rng.standard_normal((S, Cm, T_chain))
                    d = (z @ pf) * np.sqrt(js_rows[i])[None, :, None]
                    blocks.append(d.ravel())
                    offs.append(o)
                    o += d.size
                delta = np.concatenate(blocks) if blocks else np.zeros(0)
                u = rng.random((active.size, S))
                ctx.run_host_stream(active, S, delta, offs, u.ravel())
                self._acc_stats("main" if T_chain > 1 or M != 100 else "prerun")
                start += S
            st = ctx.chain_status(active)
            active = active[st["state"] == _lib.CHAIN_RUNNING]
            if not active.size:
                break

    def _acc_stats(self, phase):
        s = self.ctx.last_run_stats()
        s.update(self.ctx.work_counters(reset=True))     # what the kernels of this run counted themselves
        for k, v in s.items():
            self.stats[k] = self.stats.get(k, 0) + v
            self.stats[phase + "_" + k] = self.stats.get(phase + "_" + k, 0) + v

    # -- the pipeline ----------------------------------------------------------------------
    def upload(self, packed):
        """Host -> device copy of the packed gene batch (dicegp_upload_genes)."""
        if not (packed.n_iso > 1).all():
            raise ValueError("single-isoform genes never reach the sampler (run_utils.py:99-104); drop them first")
        self.ctx.upload_genes(packed)
        self.packed = packed

    def sample(self, X=None, theta1=3.0, theta2=None, M=20000, initial=1000, gap=100, seed=0,
               gene_ids=None, stream="device", keep_trace=False, summarize=True):
        """Pre-runs -> jump variance -> joint chains -> summaries for the uploaded genes."""
        import time as _time
        _t0 = _time.perf_counter()
        ctx, packed = self.ctx, self.packed
        G, T = packed.G, packed.T
        X = np.arange(T) if X is None else np.asarray(X)
        if theta2 is None:
            theta2 = ((np.max(X) - np.min(X) + 0.1) / 3.0) ** 2       # diceseq.py:169
        gene_ids = np.arange(G, dtype=np.int64) if gene_ids is None else np.asarray(gene_ids, dtype=np.int64)
        self.stats = {}
        wall = {}                                         # host wall time of each stage (ms)

        # 1. pre-runs (run_utils.py:107-115)
        _t = _time.perf_counter()
        ctx.set_prerun_chains(theta1, 0.1, 100, 100, 50)
        wall["set_prerun"] = (_time.perf_counter() - _t) * 1e3
        if stream == "device":
            _t = _time.perf_counter()
            sid = (gene_ids[:, None] * (T + 1) + np.arange(T)[None, :]).ravel()
            ctx.run_device_stream(seed, sid)
            self._acc_stats("prerun")
            wall["run_prerun"] = (_time.perf_counter() - _t) * 1e3
        else:
            rng = np.random.Generator(np.random.Philox(key=seed))
            js_rows = []
            for g in range(G):
                C = int(packed.n_iso[g])
                js_rows += [np.full(C - 1, 0.1 * 5 / (1 * C * theta1))] * T
            self._run_host(rng, js_rows, np.array([[np.sqrt(theta1)]]), 1, 100, 50)

        # 2. main chains from the pre-run variances (device reduction, run_utils.py:115-116)
        K = GP_K(X, [theta1, theta2])
        det = np.linalg.det(K)
        if det <= 0:
            raise TypeError("GP kernel has a non-positive determinant")
        kinv = np.linalg.inv(K)
        prior_const = -(T * _LOG_2PI + np.log(det)) / 2.0
        pf = proposal_factor(K)
        _t = _time.perf_counter()
        var = ctx.main_chains_from_prerun(M, initial, gap, theta1, kinv, pf, prior_const)
        wall["set_main"] = (_time.perf_counter() - _t) * 1e3

        # 3. main run (run_utils.py:120-122)
        if stream == "device":
            _t = _time.perf_counter()
            ctx.run_device_stream(seed, gene_ids * (T + 1) + T)
            self._acc_stats("main")
            wall["run_main"] = (_time.perf_counter() - _t) * 1e3
        else:
            js_rows = [var[g, :int(packed.n_iso[g]) - 1] * 5 / (T * int(packed.n_iso[g]) * theta1) for g in range(G)]
            self._run_host(rng, js_rows, pf, T, initial, gap)

        # 4. results: burn-in summaries are reduced on the device (run_utils.py:125-129)
        _t = _time.perf_counter()
        st = ctx.chain_status()
        wall["status"] = (_time.perf_counter() - _t) * 1e3
        if (st["state"] == _lib.CHAIN_NOMEM).any():
            raise MemoryError("%d chains ran out of trace memory; raise Context.set_trace_budget or "
                              "use smaller batches" % int((st["state"] == _lib.CHAIN_NOMEM).sum()))
        res = BatchResult(packed.n_iso, T, st, var)
        if summarize:
            _t1 = _time.perf_counter()
            res.psi_mean, res.psi_ci, res.lik_marg, res.off = ctx.posterior_summary(flat=True)
            self.stats["summary_ms"] = (_time.perf_counter() - _t1) * 1e3
        if keep_trace:
            res.traces = [ctx.download_trace(g, int(st["n_rows"][g])) for g in range(G)]
        self.stats["sample_ms"] = (_time.perf_counter() - _t0) * 1e3
        self.stats["wall_ms"] = wall
        return res

    def run(self, packed, **kw):
        self.upload(packed)
        return self.sample(**kw)


class PipelinedSampler:
    """A stream of gene batches on one GPU with the host->device copy of batch k+1 hidden behind
    the sampling of batch k: two `BatchSampler`s (two C-ABI contexts, each with its own device
    buffers and streams) alternate; the copy of the next batch is issued from a helper thread on
    the idle context's stream while the busy one runs its kernels (the copy engine and the SMs
    work concurrently; ctypes drops the GIL inside the library).  Results are identical to
    running the batches one after another through a single `BatchSampler`."""

    def __init__(self, device=0, trace_budget_bytes=0):
        self.samplers = [BatchSampler(device), BatchSampler(device)]
        self.current = None
        self.last_upload_ms = 0.0
        if not trace_budget_bytes:
            # a context sizes its trace pool from the memory that is free when its chains are
            # installed; two contexts share the GPU, so each gets a fixed 30 % of what is free now
            import torch
            trace_budget_bytes = int(0.3 * torch.cuda.mem_get_info(device)[0])
        for s in self.samplers:
            s.ctx.set_trace_budget(int(trace_budget_bytes))

    def close(self):
        for s in self.samplers:
            s.close()

    def run_stream(self, batches):
        """batches: iterable of (PackedGenes, sample-kwargs).  Yields (BatchResult, stats) in order."""
        import threading
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        self.samplers[0].upload(cur[0])
        k = 0
        while cur is not None:
            nxt = next(it, None)
            th, err = None, []
            if nxt is not None:
                def _up(s=self.samplers[(k + 1) % 2], packed=nxt[0]):
                    try:
                        import time as _time
                        _t = _time.perf_counter()
                        s.upload(packed)
                        self.last_upload_ms = (_time.perf_counter() - _t) * 1e3
                    except BaseException as e:          # re-raised in the caller's thread below
                        err.append(e)
                th = threading.Thread(target=_up)
                th.start()
            s = self.samplers[k % 2]
            self.current = s                              # the context that holds this batch's traces
            try:
                res = s.sample(**cur[1])
            finally:
                if th is not None:
                    th.join()
            if err:
                raise err[0]
            yield res, dict(s.stats)
            cur = nxt
            k += 1
