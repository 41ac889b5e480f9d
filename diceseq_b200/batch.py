"""Batched sampler: what `get_psi` does around Psi_GP_MH (diceseq/utils/run_utils.py:106-131),
for many genes at once on one GPU.

Per gene the reference runs T single-time-point pre-runs (M=100, var=0.1) to set the jump
variance, then the joint chain, then burn-in summaries.  Here all genes' pre-runs are one
set of G*T chains, the jump variances are reduced on the device
(`dicegp_main_chains_from_prerun`), and the G main chains run to Geweke convergence.

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
    __slots__ = ("m", "cnt", "state", "flags", "psi_mean", "psi_95", "lik_marg", "var", "trace")

    def __init__(self):
        self.trace = None


class BatchSampler:
    """One GPU; `run` samples every gene of a PackedGenes batch."""

    def __init__(self, device=0):
        self.ctx = _lib.Context(device)
        self.stats = {}

    def close(self):
        self.ctx.close()

    # -- chain construction ----------------------------------------------------------------
    def _prerun_chains(self, packed, theta1, var0=0.1, M=100, initial=100, gap=50):
        """G*T chains of one time point each (run_utils.py:111-114); chain g*T+j."""
        K1 = np.array([[float(theta1)]])                 # GP_K of a single point
        kinv1 = np.linalg.inv(K1)
        prior_const = -(_LOG_2PI + np.log(np.linalg.det(K1))) / 2.0
        pf1 = proposal_factor(K1)
        pool = [kinv1.ravel(), pf1.ravel()]
        off = 2
        per_C = {}
        for C in np.unique(packed.n_iso):
            C = int(C)
            js = np.full(max(C - 1, 0), var0 * 5 / (1 * C * theta1))
            jc = np.array([-(_LOG_2PI + np.log(np.linalg.det(K1 * var0 * 5 / (1 * C * theta1)))) / 2.0] * max(C - 1, 0))
            per_C[C] = (off, off + len(js))
            pool += [js, jc]
            off += 2 * len(js)
        descs = []
        for g in range(packed.G):
            C = int(packed.n_iso[g])
            jso, jco = per_C[C]
            for j in range(packed.T):
                descs.append(_lib.ChainDesc(gene=g, t0=j, Tc=1, M=M, initial=initial, gap=gap,
                                            kinv_off=0, prop_factor_off=1, jump_scale_off=jso,
                                            jump_const_off=jco, y0_off=-1, prior_const=float(prior_const)))
        return descs, np.concatenate(pool)

    # -- host stream ---------------------------------------------------------------------
    def _run_host(self, rng, js_rows, pf, T_chain):
        """Advance all unfinished chains to the end with host-generated streams.
        js_rows[i]: jump scales (C-1,) of chain i; pf: (Tc,Tc) proposal factor."""
        ctx = self.ctx
        n = len(ctx.chain_shapes)
        active = np.arange(n)
        M = ctx.chain_shapes[0][2]
        ini, gap = self._ini_gap
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
                self._acc_stats()
                start += S
            st = ctx.chain_status(active)
            active = active[st["state"] == _lib.CHAIN_RUNNING]
            if not active.size:
                break

    def _acc_stats(self):
        s = self.ctx.last_run_stats()
        for k, v in s.items():
            self.stats[k] = self.stats.get(k, 0) + v

    # -- the pipeline ----------------------------------------------------------------------
    def run(self, packed, X=None, theta1=3.0, theta2=None, M=20000, initial=1000, gap=100,
            seed=0, gene_ids=None, stream="device", keep_trace=False, summarize=True):
        ctx = self.ctx
        G, T = packed.G, packed.T
        X = np.arange(T) if X is None else np.asarray(X)
        if theta2 is None:
            theta2 = ((np.max(X) - np.min(X) + 0.1) / 3.0) ** 2       # diceseq.py:169
        gene_ids = np.arange(G, dtype=np.int64) if gene_ids is None else np.asarray(gene_ids, dtype=np.int64)
        self.stats = {}
        multi = packed.n_iso > 1
        if not multi.all():
            raise ValueError("single-isoform genes never reach the sampler (run_utils.py:99-104); drop them first")
        ctx.upload_genes(packed)

        # 1. pre-runs
        descs, pool = self._prerun_chains(packed, theta1)
        ctx.set_chains(descs, pool)
        if stream == "device":
            sid = (gene_ids[:, None] * (T + 1) + np.arange(T)[None, :]).ravel()
            ctx.run_device_stream(seed, sid)
            self._acc_stats()
        else:
            rng = np.random.Generator(np.random.Philox(key=seed))
            self._ini_gap = (100, 50)
            js_rows = []
            for g in range(G):
                C = int(packed.n_iso[g])
                js_rows += [np.full(C - 1, 0.1 * 5 / (1 * C * theta1))] * T
            self._run_host(rng, js_rows, proposal_factor(np.array([[float(theta1)]])), 1)

        # 2. main chains from the pre-run variances (device reduction)
        K = GP_K(X, [theta1, theta2])
        det = np.linalg.det(K)
        if det <= 0:
            raise TypeError("GP kernel has a non-positive determinant")
        kinv = np.linalg.inv(K)
        prior_const = -(T * _LOG_2PI + np.log(det)) / 2.0
        pf = proposal_factor(K)
        var = ctx.main_chains_from_prerun(M, initial, gap, theta1, kinv, pf, prior_const)

        # 3. main run
        if stream == "device":
            ctx.run_device_stream(seed, gene_ids * (T + 1) + T)
            self._acc_stats()
        else:
            self._ini_gap = (initial, gap)
            js_rows = [var[g, :int(packed.n_iso[g]) - 1] * 5 / (T * int(packed.n_iso[g]) * theta1) for g in range(G)]
            self._run_host(rng, js_rows, pf, T)

        # 4. results: burn-in summaries are reduced on the device (run_utils.py:125-129)
        st = ctx.chain_status()
        if (st["state"] == _lib.CHAIN_NOMEM).any():
            raise MemoryError("%d chains ran out of trace memory; raise Context.set_trace_budget or "
                              "use smaller batches" % int((st["state"] == _lib.CHAIN_NOMEM).sum()))
        if summarize:
            means, cis, lik = ctx.posterior_summary()
        out = []
        for g in range(G):
            r = GeneResult()
            r.m, r.cnt, r.state, r.flags = int(st["m"][g]), int(st["cnt"][g]), int(st["state"][g]), int(st["flags"][g])
            r.var = var[g, :int(packed.n_iso[g]) - 1].copy()
            if summarize:
                r.psi_mean = means[g]
                r.psi_95 = [cis[g][c] for c in range(cis[g].shape[0])]
                r.lik_marg = float(lik[g])
            if keep_trace:
                r.trace = ctx.download_trace(g, int(st["n_rows"][g]))
            out.append(r)
        return out
