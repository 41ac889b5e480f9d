"""Batched pipeline on the GPU: device-generated streams, properties that hold at any size."""
import numpy as np
import pytest

from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
from oracle import psi_gp_mh_oracle as orc

pytestmark = pytest.mark.gpu


def _batch(config, genes, **kw):
    return pack_genes([synth.gene_W(config, g, **kw) for g in genes])


def _loglik_numpy(W_list, lens, Y):
    """sum_t sum_n log(W_t f_t) recomputed on the host from a trace row."""
    tot = 0.0
    for t, W in enumerate(W_list):
        e = np.exp(Y[:, t])
        psi = e / e.sum()
        f = lens[t] * psi
        f = f / f.sum()
        tot += np.log(W @ f).sum()
    return tot


@pytest.fixture(scope="module")
def sampler():
    s = BatchSampler(0)
    yield s
    s.close()


def test_trace_rows_are_self_consistent_full_size(sampler):
    """Size-independent property at BASELINE shapes (1000 reads x 8 time points, C in 2..8):
    every stored Lik_all[m] equals the likelihood of the stored Y_all[m], psi = softmax(y),
    Pro - Lik = GP prior of y, and the Geweke rule holds at the returned m."""
    genes = list(range(40, 52))
    packed = _batch("timeseries", genes)
    res = sampler.run(packed, M=3000, initial=1000, gap=100, seed=5, gene_ids=genes, keep_trace=True)
    T = 8
    X = np.arange(T)
    K = orc.gp_kernel(X, [3.0, synth.default_theta2(X)])
    for g, r in zip(genes, res):
        W_list, lens = synth.gene_W("timeseries", g)
        Psi, Y, Pro, Lik = r.trace
        assert Psi.shape[0] == (r.m if r.state == 1 else 3000)
        for row in (0, len(Lik) // 2, len(Lik) - 1):
            ll = _loglik_numpy(W_list, lens, Y[row])
            assert abs(ll - Lik[row]) <= 1e-12 * abs(ll)
            e = np.exp(Y[row])
            assert np.allclose(Psi[row], e / e.sum(axis=0), rtol=1e-13, atol=0)
            prior = sum(orc.log_mvn_pdf(Y[row, c], np.zeros(T), K) for c in range(Y.shape[1] - 1))
            assert abs((Pro[row] - Lik[row]) - prior) <= 1e-9 * max(1.0, abs(prior))
        assert (Y[:, -1, :] == 0).all()
        if r.state == 1:
            for c in range(Psi.shape[1] - 1):
                for t in range(T):
                    z = orc.geweke_z(Psi[:r.m, c, t])
                    assert z is not None and z <= 2
        assert 0 < r.cnt <= len(Lik)


def test_device_stream_is_deterministic_and_partition_invariant(sampler):
    genes = [60, 61, 62, 63, 64, 65]
    a = sampler.run(_batch("static", genes), M=1500, seed=9, gene_ids=genes, keep_trace=True)
    b = sampler.run(_batch("static", genes[3:]), M=1500, seed=9, gene_ids=genes[3:], keep_trace=True)
    for ra, rb in zip(a[3:], b):
        assert ra.m == rb.m and ra.cnt == rb.cnt
        assert np.array_equal(ra.trace[2], rb.trace[2])
    c = sampler.run(_batch("static", genes), M=1500, seed=10, gene_ids=genes, keep_trace=True)
    assert any(not np.array_equal(x.trace[2][:50], y.trace[2][:50]) for x, y in zip(a, c))


def test_posterior_matches_oracle_within_mc_error(sampler):
    """Independent streams: posterior mean psi of the GPU chain vs the oracle's (MT19937)
    on the same genes.  Tolerance 0.05 absolute on psi (CI half-widths are 0.05-0.1)."""
    genes = [70, 71, 72, 73]
    packed = _batch("timeseries", genes, reads=200, T=3)
    res = sampler.run(packed, M=4000, initial=1000, gap=100, seed=3, gene_ids=genes)
    X = np.arange(3)
    for g, r in zip(genes, res):
        R, L, P = synth.gene_matrices("timeseries", g, reads=200, T=3)
        with np.errstate(all="ignore"):
            out, var, _ = orc.prerun_and_main(R, L, P, X, 3.0, float(synth.default_theta2(X)), 4000, 1000, 100,
                                              rng=np.random.RandomState(g))
        Psi, Y, th2, Pro, Lik, cnt, m = out
        ref_mean = Psi[int(m / 4):].mean(axis=0)
        assert np.abs(ref_mean - r.psi_mean).max() < 0.05, (g, ref_mean, r.psi_mean)
        assert 0.2 < r.var.mean() / var.mean() < 5.0


def test_host_stream_batch_runs_and_is_consistent(sampler):
    genes = [80, 81, 82]
    packed = _batch("timeseries", genes, reads=150, T=4)
    res = sampler.run(packed, M=2000, initial=1000, gap=100, seed=4, gene_ids=genes, stream="host", keep_trace=True)
    for g, r in zip(genes, res):
        W_list, lens = synth.gene_W("timeseries", g, reads=150, T=4)
        Psi, Y, Pro, Lik = r.trace
        ll = _loglik_numpy(W_list, lens, Y[-1])
        assert abs(ll - Lik[-1]) <= 1e-12 * abs(ll)
        assert r.state in (1, 2) and r.cnt > 0


def test_device_summary_matches_numpy(sampler):
    """Posterior summary kernel vs the oracle's restatement of run_utils.py:125-131 on the
    downloaded trace: mean (sequential axis-0 sum) and the order-statistic interval must be
    bit-exact, the harmonic-mean log-likelihood within 1e-13 relative."""
    genes = [90, 91, 92, 93, 94]
    packed = _batch("timeseries", genes, reads=120, T=3)
    res = sampler.run(packed, M=1700, initial=1000, gap=100, seed=11, gene_ids=genes, keep_trace=True)
    for r in res:
        Psi, Y, Pro, Lik = r.trace
        psi_mean, psi_95, lik_marg, _ = orc.posterior_summary(Psi, Y, Lik, r.m)
        assert np.array_equal(r.psi_mean, psi_mean)
        for c in range(Psi.shape[1]):
            assert np.array_equal(r.psi_95[c], psi_95[c])
        assert abs(r.lik_marg - lik_marg) <= 1e-13 * abs(lik_marg)


def test_results_do_not_depend_on_the_pass_plan(sampler, monkeypatch):
    """Which kernel advances a chain and where the passes are cut must not show in the results: same counter-based
    stream, same arithmetic per step.  Default plan (small batch: general kernel to step 1024, then wide rounds)
    against general kernel to 4096 + staged cluster passes, and against a very early hand-over."""
    genes = list(range(700, 760))
    packed = _batch("timeseries", genes, reads=150, T=4)
    ref = sampler.run(packed, seed=17, gene_ids=genes)
    for env in ({"DICEGP_CAP0": "4096"}, {"DICEGP_CAP0": "4096", "DICEGP_WIDE_CS": "2"}, {"DICEGP_CAP0": "200"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = sampler.run(packed, seed=17, gene_ids=genes)
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(got.m, ref.m) and np.array_equal(got.cnt, ref.cnt), env
        assert np.array_equal(got.psi_mean, ref.psi_mean), env
        assert np.all(np.abs(got.lik_marg - ref.lik_marg) <= 1e-12 * np.abs(ref.lik_marg)), env


def test_summary_of_long_chains_split_over_blocks(sampler):
    """Chains of 7k-28k rows take the size classes that run eight blocks per chain (each a share of the
    columns, the last one the log-likelihood too): same bit-exact summaries."""
    genes = [95, 96]
    packed = _batch("timeseries", genes, reads=80, T=3)
    res = sampler.run(packed, M=9000, initial=9000, gap=1000, seed=5, gene_ids=genes, keep_trace=True)
    for r in res:
        Psi, Y, Pro, Lik = r.trace
        assert len(Lik) == 9000
        psi_mean, psi_95, lik_marg, _ = orc.posterior_summary(Psi, Y, Lik, r.m)
        assert np.array_equal(r.psi_mean, psi_mean)
        for c in range(Psi.shape[1]):
            assert np.array_equal(r.psi_95[c], psi_95[c])
        assert abs(r.lik_marg - lik_marg) <= 1e-13 * abs(lik_marg)


def test_summary_of_chains_longer_than_shared_memory():
    """`--mcmc 0 50000 ...`: a chain that keeps more rows after the burn-in than one column of keys fits
    shared memory (about 28.5k) is summarised through a global scratch row: same bit-exact mean /
    order statistics (the reference accepts any max_run)."""
    genes = [97, 98]
    packed = _batch("static", genes, reads=60)
    s = BatchSampler(0)
    try:
        res = s.run(packed, M=50000, initial=50000, gap=1000, seed=3, gene_ids=genes, keep_trace=True)
        for r in res:
            Psi, Y, Pro, Lik = r.trace
            assert r.m == 49999 and len(Lik) == 50000
            psi_mean, psi_95, lik_marg, _ = orc.posterior_summary(Psi, Y, Lik, r.m)
            assert np.array_equal(r.psi_mean, psi_mean)
            for c in range(Psi.shape[1]):
                assert np.array_equal(r.psi_95[c], psi_95[c])
            assert abs(r.lik_marg - lik_marg) <= 1e-13 * abs(lik_marg)
    finally:
        s.close()


def test_trace_pool_grows_and_reports_exhaustion():
    """Chains longer than their first page allocate pages on the device; a pool that is too
    small stops the chain with CHAIN_NOMEM instead of corrupting memory."""
    from diceseq_b200 import _lib
    genes = [95, 96]
    packed = _batch("static", genes, reads=100)
    s = BatchSampler(0)
    try:
        res = s.run(packed, M=5000, initial=4000, gap=500, seed=1, gene_ids=genes, keep_trace=True)
        for g, r in zip(genes, res):
            W_list, lens = synth.gene_W("static", g, reads=100)
            Psi, Y, Pro, Lik = r.trace
            assert len(Lik) >= 4000
            for row in (0, 1023, 1024, 2048, len(Lik) - 1):
                ll = _loglik_numpy(W_list, lens, Y[row])
                assert abs(ll - Lik[row]) <= 1e-12 * abs(ll)
        s.ctx.set_trace_budget(2 * 1024 * 8 * 8)       # room for the first pages only
        with pytest.raises(MemoryError):
            s.run(packed, M=5000, initial=4000, gap=500, seed=1, gene_ids=genes)
    finally:
        s.close()


def test_pipelined_stream_equals_one_batch_at_a_time():
    """PipelinedSampler (copy of batch k+1 hidden behind the sampling of batch k, two contexts)
    returns bit-identical results to a single BatchSampler fed the same batches in turn."""
    from diceseq_b200.batch import PipelinedSampler
    batches = [(_batch("timeseries", list(range(60 + 6 * b, 66 + 6 * b))), dict(M=2000, seed=b, gene_ids=np.arange(6) + 6 * b))
               for b in range(3)]
    pipe = PipelinedSampler(0)
    got = [r for r, _st in pipe.run_stream(iter(batches))]
    pipe.close()
    one = BatchSampler(0)
    for (packed, kw), r in zip(batches, got):
        ref = one.run(packed, **kw)
        assert np.array_equal(ref.m, r.m) and np.array_equal(ref.cnt, r.cnt)
        assert np.array_equal(ref.psi_mean, r.psi_mean) and np.array_equal(ref.psi_ci, r.psi_ci)
        assert np.array_equal(ref.lik_marg, r.lik_marg)
    one.close()


def test_human_scale_batch_is_self_consistent(sampler):
    """BASELINE config 5 through the batched device-stream path: ragged heavy-tailed genes of
    every residency tier in one batch; the last trace row of every gene re-evaluates to its
    stored likelihood, and converged genes satisfy the Geweke rule at the returned m."""
    genes = list(range(0, 40))
    packed = _batch("human", genes)
    res = sampler.run(packed, M=2500, initial=1000, gap=100, seed=21, gene_ids=genes, keep_trace=True)
    n_conv = 0
    for g, r in zip(genes, res):
        W_list, lens = synth.gene_W("human", g)
        Psi, Y, Pro, Lik = r.trace
        assert r.state in (1, 2) and 0 < r.cnt <= len(Lik)
        assert Psi.shape[0] == (r.m if r.state == 1 else 2500)
        for row in (0, len(Lik) - 1):
            ll = _loglik_numpy(W_list, lens, Y[row])
            assert abs(ll - Lik[row]) <= 1e-12 * abs(ll), (g, row, ll, Lik[row])
        if r.state == 1:
            n_conv += 1
            for c in range(Psi.shape[1] - 1):
                for t in range(Psi.shape[2]):
                    z = orc.geweke_z(Psi[:r.m, c, t])
                    assert z is not None and z <= 2
    assert n_conv > 0


def test_warp_per_chain_prerun_kernel_equals_general_kernel(sampler, monkeypatch):
    """The pre-runs (single time point, no convergence check) take the warp-per-chain kernel;
    DICEGP_PRERUN_KERNEL=0 sends them through the general chain kernel.  Same stream, same
    arithmetic per step: identical accept decisions, hence bit-identical jump variances and an
    identical main run.  Includes odd read counts, a 12-isoform gene (run-time isoform count) and
    ragged human-scale genes (those too large for shared memory fall back to the general kernel)."""
    genes = [synth.gene_W("timeseries", g, reads=r, T=3) for g, r in zip(range(100, 112), [999, 1000, 1, 37, 501] * 3)]
    genes += [synth.gene_W("timeseries", 112, reads=333, T=3, C=12)]
    for g in (0, 1, 2, 3):                      # gene 0: 642 / 9948 / 6272 reads x 6 isoforms -> two chains fall back
        W, lens = synth.gene_W("human", g)
        genes.append((W[:3], lens[:3]))
    packed = pack_genes(genes)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DICEGP_PRERUN_KERNEL", mode)
        res = sampler.run(packed, M=1300, initial=1000, gap=100, seed=17)
        out[mode] = (res.var.copy(), res.m.copy(), res.cnt.copy(), res.psi_mean.copy())
        n_launch = sampler.stats["prerun_launches"]
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.array_equal(out["0"][1], out["1"][1]) and np.array_equal(out["0"][2], out["1"][2])
    assert np.array_equal(out["0"][3], out["1"][3])
    assert n_launch > 0


def test_warp_per_chain_kernel_runs_static_mode_like_the_general_kernel(sampler, monkeypatch):
    """Static mode (one time point): the joint chains are single-time-point chains too and take the
    warp-per-chain kernel, with convergence checks and trace pages growing on the device.  Against
    the general kernel (DICEGP_PRERUN_KERNEL=0): same m, cnt and Y / Psi traces bit for bit,
    Pro / Lik within 1e-13 (the likelihood is summed in a different order)."""
    genes = list(range(200, 216))
    packed = _batch("static", genes)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DICEGP_PRERUN_KERNEL", mode)
        res = sampler.run(packed, M=2600, initial=1500, gap=100, seed=23, gene_ids=genes, keep_trace=True)
        out[mode] = [(r.m, r.cnt, r.state, r.trace) for r in res]
    long_ones = 0
    for a, b in zip(out["0"], out["1"]):
        assert a[:3] == b[:3]
        assert np.array_equal(a[3][0], b[3][0]) and np.array_equal(a[3][1], b[3][1])
        for k in (2, 3):
            assert np.allclose(a[3][k], b[3][k], rtol=1e-13, atol=0)
        long_ones += a[0] > 1024                      # crossed a trace page
    assert long_ones > 0
