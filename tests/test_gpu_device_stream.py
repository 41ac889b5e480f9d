"""Step parity of the DEVICE-stream mode -- the path bench.py times -- against the oracle.

`dicegp_run_device_stream` draws its proposals / uniforms from a counter-based Philox stream on
the device, so the oracle cannot be fed "the same seed".  `dicegp_dump_stream` writes that very
stream out (same device functions, same order of operations as the chain kernels); replaying
the dump through the oracle (`replay=`) must give the same accept decisions (identical m, cnt)
and traces within 1e-12 relative, for
  * `dicegp_prerun_kernel<C>` (one warp per single-time-point chain: the pre-runs of get_psi,
    diceseq/utils/run_utils.py:107-114, and every chain of static mode),
  * `dicegp_prerun_var_kernel` (jump variance, run_utils.py:115-116: BIT-exact against numpy on
    the downloaded pre-run traces),
  * `dicegp_chain_kernel` / `dicegp_tchain_kernel` in device-stream mode (joint chains, all tiers).
The dump itself is pinned by an independent numpy restatement of Philox4x32-10 + Box-Muller.
"""
import numpy as np
import pytest

from diceseq_b200 import _lib, synth
from diceseq_b200.model_GP import GP_K, proposal_factor
from diceseq_b200.packer import clean_inputs, pack_genes
from oracle import psi_gp_mh_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-12
THETA1 = 3.0


def rel_close(a, b, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    with np.errstate(invalid="ignore"):
        same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
        ok = same_inf | (np.isnan(a) & np.isnan(b)) | \
            (np.abs(a - b) <= RTOL * np.maximum(np.abs(a), np.abs(b)) + 1e-300)
    assert ok.all(), "%s: %d of %d entries differ, first at %d: %r vs %r" % (
        what, (~ok).sum(), ok.size, np.argmin(ok.ravel()), a.ravel()[np.argmin(ok.ravel())], b.ravel()[np.argmin(ok.ravel())])


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


# ----------------------------------------------------------------------------------------
# independent host restatement of the stream (Salmon et al. 2011, Philox4x32-10)
# ----------------------------------------------------------------------------------------
def philox4x32_10(c, k):
    c = [np.uint64(x) for x in c]
    k = [np.uint64(x) for x in k]
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & mask]
        k = [(k[0] + np.uint64(0x9E3779B9)) & mask, (k[1] + np.uint64(0xBB67AE85)) & mask]
    return [int(x) for x in c]


def host_stream(seed, sid, step, n_normals):
    """z[0..n_normals) and u of one step, as dicegp_kernels.cuh documents them: block (step, pair i)
    gives z[2i] (cosine branch) and z[2i+1] (sine branch); block (step, 0xFFFFFFFF) the uniform."""
    key = (seed & 0xFFFFFFFF, seed >> 32)
    z = np.empty(2 * ((n_normals + 1) // 2))
    for i in range((n_normals + 1) // 2):
        o = philox4x32_10((step, i, sid & 0xFFFFFFFF, sid >> 32), key)
        u1 = ((o[0] >> 5) * 67108864.0 + (o[1] >> 6) + 0.5) / 9007199254740992.0
        u2 = ((o[2] >> 5) * 67108864.0 + (o[3] >> 6) + 0.5) / 9007199254740992.0
        r = np.sqrt(-2.0 * np.log(u1))
        z[2 * i], z[2 * i + 1] = r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)
    o = philox4x32_10((step, 0xFFFFFFFF, sid & 0xFFFFFFFF, sid >> 32), key)
    u = ((o[0] >> 5) * 67108864.0 + (o[1] >> 6)) / 9007199254740992.0
    return z[:n_normals], u


def test_dumped_stream_matches_host_philox_replica(ctx):
    """The uniforms are integers scaled by 2^-53: bit-exact.  The normals go through log / sqrt /
    sincospi on the device and np.log / np.cos on the host: equal to rounding (1e-13 of the scale)."""
    T, C = 3, 4
    packed = pack_genes([synth.gene_W("timeseries", 5, reads=40, T=T, C=C)])
    ctx.upload_genes(packed)
    X = np.arange(T)
    K = GP_K(X, [THETA1, synth.default_theta2(X)])
    pf = proposal_factor(K)
    js = np.array([0.02, 0.5, 1.7])
    pool = np.concatenate([np.linalg.inv(K).ravel(), pf.ravel(), js, np.zeros(C - 1)])
    d = _lib.ChainDesc(gene=0, t0=0, Tc=T, M=50, initial=50, gap=10, kinv_off=0, prop_factor_off=T * T,
                       jump_scale_off=2 * T * T, jump_const_off=2 * T * T + C - 1, y0_off=-1, prior_const=0.0)
    ctx.set_chains([d], pool)
    seed, sid = 0x1234567890ABCDEF, (7 << 32) + 12345
    delta, u = ctx.dump_stream(0, seed, sid, 9, step0=4090)
    for s in range(9):
        z, uh = host_stream(seed, sid, 4090 + s, (C - 1) * T)
        assert u[s] == uh
        ref = np.sqrt(js)[:, None] * (z.reshape(C - 1, T) @ pf)
        assert np.allclose(delta[s], ref, rtol=1e-12, atol=1e-13 * np.abs(ref).max())


# ----------------------------------------------------------------------------------------
# the batched pipeline, stage by stage, with every chain's stream dumped
# ----------------------------------------------------------------------------------------
def run_pipeline(ctx, inputs, gene_ids, seed, M, initial, gap, X=None):
    """inputs: list of (R_list, len_list, prob_list) as Psi_GP_MH takes them.  Runs what
    BatchSampler.sample runs (pre-runs -> device variance -> joint chains), keeping the traces and
    the dumped streams of both stages."""
    T = len(inputs[0][0])
    X = np.arange(T) if X is None else X
    theta2 = float(synth.default_theta2(X))
    cleaned = []
    for R, L, P in inputs:
        R, L, P = [r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P]
        clean_inputs(R, L, P)
        cleaned.append(([r * p for r, p in zip(R, P)], L))
    packed = pack_genes(cleaned)
    G = packed.G
    gene_ids = np.asarray(gene_ids, dtype=np.int64)
    ctx.upload_genes(packed)
    ctx.set_prerun_chains(THETA1, 0.1, 100, 100, 50)
    sid = (gene_ids[:, None] * (T + 1) + np.arange(T)[None, :]).ravel()
    pre_dump = [ctx.dump_stream(i, seed, int(sid[i]), 100) for i in range(G * T)]
    ctx.run_device_stream(seed, sid)
    pre_st = ctx.chain_status()
    pre_tr = [ctx.download_trace(i, int(pre_st["n_rows"][i])) for i in range(G * T)]
    K = GP_K(X, [THETA1, theta2])
    kinv, det = np.linalg.inv(K), np.linalg.det(K)
    prior_const = -(T * np.log(2 * np.pi) + np.log(det)) / 2.0
    var = ctx.main_chains_from_prerun(M, initial, gap, THETA1, kinv, proposal_factor(K), prior_const)
    sid2 = gene_ids * (T + 1) + T
    main_dump = [ctx.dump_stream(g, seed, int(sid2[g]), M) for g in range(G)]
    ctx.run_device_stream(seed, sid2)
    st = ctx.chain_status()
    tr = [ctx.download_trace(g, int(st["n_rows"][g])) for g in range(G)]
    return dict(T=T, X=X, theta2=theta2, pre_dump=pre_dump, pre_st=pre_st, pre_tr=pre_tr, var=var,
                main_dump=main_dump, st=st, tr=tr, n_iso=packed.n_iso)


def check_chain(got_trace, got_m, got_cnt, ref, what):
    Psi, Y, _th2, Pro, Lik, cnt, m = ref
    assert got_m == m, (what, "m", got_m, m)
    assert got_cnt == cnt, (what, "cnt", got_cnt, cnt)
    rel_close(got_trace[2], Pro, what + " Pro_all")
    rel_close(got_trace[3], Lik, what + " Lik_all")
    rel_close(got_trace[0], Psi, what + " Psi_all")
    rel_close(got_trace[1], Y, what + " Y_all")


def check_preruns(out, inputs):
    T = out["T"]
    for g, (R, L, P) in enumerate(inputs):
        C = int(out["n_iso"][g])
        for j in range(T):
            i = g * T + j
            with np.errstate(all="ignore"):
                ref = orc.psi_gp_mh([R[j].copy()], [L[j].copy()], [P[j].copy()], out["X"][j:j + 1], np.zeros((C, 1)),
                                    0.1 * np.ones(C - 1), THETA1, out["theta2"], 100, 100, 50, replay=out["pre_dump"][i])
            check_chain(out["pre_tr"][i], int(out["pre_st"]["m"][i]), int(out["pre_st"]["cnt"][i]), ref,
                        "pre-run gene %d time %d" % (g, j))


def check_variance_bit_exact(out):
    """var[c] = (sum_j np.var(Y_j[int(m_j/4):, :-1, 0], axis=0) + 1e-6) / T   (run_utils.py:113-116)
    from the DEVICE's own pre-run traces: numpy's summation order is reproduced on the device."""
    T = out["T"]
    for g in range(len(out["n_iso"])):
        C = int(out["n_iso"][g])
        acc = np.zeros(C - 1)
        for j in range(T):
            i = g * T + j
            Y, m = out["pre_tr"][i][1], int(out["pre_st"]["m"][i])
            acc += np.var(Y[int(m / 4):, :-1, 0], axis=0)
        want = (acc + 0.000001) / T
        assert np.array_equal(out["var"][g, :C - 1], want), (g, out["var"][g, :C - 1], want)


def check_main(out, inputs, M, initial, gap):
    for g, (R, L, P) in enumerate(inputs):
        C = int(out["n_iso"][g])
        with np.errstate(all="ignore"):
            ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], out["X"],
                                np.zeros((C, out["T"])), out["var"][g, :C - 1].copy(), THETA1, out["theta2"], M, initial, gap,
                                replay=out["main_dump"][g])
        check_chain(out["tr"][g], int(out["st"]["m"][g]), int(out["st"]["cnt"][g]), ref, "joint chain gene %d" % g)


def zero_row_gene(T, C, n=21):
    """R and prob non-zero in different columns of one read -> that row of W
    is zero -> log 0 = -inf, the chain never moves; the row survives the input cleaning."""
    rng = np.random.RandomState(99)
    R, L, P = [], [], []
    for _t in range(T):
        r = np.ones((n, C), bool)
        p = rng.rand(n, C) * 1e-3
        r[4, :] = [True] + [False] * (C - 1)
        p[4, :] = [0.0] + [1e-3] * (C - 1)
        R.append(r), L.append(np.arange(1, C + 1) * 700.0), P.append(p)
    return R, L, P


def test_warp_per_chain_prerun_kernel_step_parity_and_variance(ctx, monkeypatch):
    """dicegp_prerun_kernel<C> for C in {2, 3, 5, 8} (compiled) and 12 (run-time C), odd read counts,
    a 1-read time point and a -inf row; then the device's jump variance, bit-exact."""
    monkeypatch.setenv("DICEGP_PRERUN_KERNEL", "1")
    inputs = [synth.gene_matrices("timeseries", 300 + k, reads=r, T=3, C=C)
              for k, (C, r) in enumerate([(2, 999), (3, 1000), (5, 501), (8, 37), (12, 333), (2, 1), (4, 64)])]
    inputs.append(zero_row_gene(3, 3))
    out = run_pipeline(ctx, inputs, np.arange(300, 300 + len(inputs)), seed=17, M=300, initial=200, gap=100)
    assert (out["pre_st"]["m"] == 99).all() and (out["pre_st"]["state"] == _lib.CHAIN_EXHAUSTED).all()
    check_preruns(out, inputs)
    check_variance_bit_exact(out)
    assert not np.isfinite(out["tr"][-1][3]).any()            # the -inf gene never leaves -inf
    check_main(out, inputs, 300, 200, 100)


def test_general_kernel_prerun_step_parity_and_variance(ctx, monkeypatch):
    """The same pre-runs through the general chain kernel (DICEGP_PRERUN_KERNEL=0)."""
    monkeypatch.setenv("DICEGP_PRERUN_KERNEL", "0")
    inputs = [synth.gene_matrices("timeseries", 320 + k, reads=r, T=2, C=C)
              for k, (C, r) in enumerate([(2, 501), (3, 64), (6, 200), (12, 99)])]
    out = run_pipeline(ctx, inputs, np.arange(320, 320 + len(inputs)), seed=5, M=250, initial=200, gap=50)
    check_preruns(out, inputs)
    check_variance_bit_exact(out)
    check_main(out, inputs, 250, 200, 50)


@pytest.mark.parametrize("kernel", ["1", "0"])
def test_static_mode_device_stream_step_parity(ctx, monkeypatch, kernel):
    """BASELINE config 2 (one time point): the joint chains are single-time-point chains too, run by
    the warp-per-chain kernel with convergence checks and trace pages growing on the device."""
    monkeypatch.setenv("DICEGP_PRERUN_KERNEL", kernel)
    genes = [410, 411, 412, 413]
    inputs = [synth.gene_matrices("static", g) for g in genes]
    out = run_pipeline(ctx, inputs, genes, seed=23, M=2300, initial=1100, gap=100)
    check_preruns(out, inputs)
    check_variance_bit_exact(out)
    check_main(out, inputs, 2300, 1100, 100)
    assert (out["st"]["m"] > 1024).any()                       # crossed a trace page


@pytest.mark.parametrize("C", [2, 3, 4, 5, 6, 7, 8])
def test_timeseries_device_stream_step_parity(ctx, C):
    """BASELINE config 3 shapes (8 time points x 1000 reads, C = 2..8): the joint chain in
    device-stream mode, one convergence check or more, against the oracle on the dumped stream."""
    genes = [500 + C]
    inputs = [synth.gene_matrices("timeseries", g, C=C) for g in genes]
    out = run_pipeline(ctx, inputs, genes, seed=31 + C, M=1250, initial=1000, gap=100)
    check_variance_bit_exact(out)
    check_main(out, inputs, 1250, 1000, 100)


def test_straggler_pass_resumes_device_stream_bit_identically(ctx, monkeypatch):
    """Chains cut into passes (throughput shapes, then latency shapes) continue the same counter-based
    stream: same parity as an uncut run."""
    monkeypatch.setenv("DICEGP_STRAGGLER_STEPS", "450")
    genes = [520, 521, 522]
    inputs = [synth.gene_matrices("timeseries", g, reads=300, T=4, C=C) for g, C in zip(genes, (3, 5, 8))]
    out = run_pipeline(ctx, inputs, genes, seed=41, M=1150, initial=1000, gap=50)
    check_main(out, inputs, 1150, 1000, 50)


@pytest.mark.parametrize("cs", ["8", "2"])
def test_last_stragglers_on_clusters_step_parity(ctx, monkeypatch, cs):
    """Three passes: throughput shapes, one chain per SM (wide-round kernel), then the wide-round kernel on
    thread-block clusters (every chain split over `cs` SMs, partial likelihoods exchanged through DSMEM,
    lead block writes the trace and runs the Geweke rule): same step parity with the oracle."""
    monkeypatch.setenv("DICEGP_STRAGGLER_STEPS", "450,700")
    monkeypatch.setenv("DICEGP_WIDE_CS", cs)
    genes = [540, 541, 542, 543]
    inputs = [synth.gene_matrices("timeseries", g, reads=300, T=4, C=C) for g, C in zip(genes, (2, 3, 5, 8))]
    out = run_pipeline(ctx, inputs, genes, seed=47, M=1250, initial=1000, gap=50)
    check_main(out, inputs, 1250, 1000, 50)


def test_cluster_passes_with_a_minus_inf_gene(ctx, monkeypatch):
    """The real-log redo path on a cluster: a gene whose likelihood is -inf makes every
    round exchange its partial sums twice through DSMEM; the chain never moves and stops at the first check,
    the ordinary gene beside it keeps step parity."""
    monkeypatch.setenv("DICEGP_STRAGGLER_STEPS", "150,250")
    monkeypatch.setenv("DICEGP_WIDE_CS", "4")
    genes = [550, 551]
    inputs = [zero_row_gene(3, 3), synth.gene_matrices("timeseries", 551, reads=200, T=3, C=4)]
    out = run_pipeline(ctx, inputs, genes, seed=53, M=1150, initial=1000, gap=50)
    check_main(out, inputs, 1150, 1000, 50)
    assert int(out["st"]["cnt"][0]) == 0


def test_large_genes_on_wide_cluster_rounds_step_parity(ctx, monkeypatch):
    """Straggler passes of large genes (W >= 1 MB: the cluster tiers): 16-candidate rounds on a cluster of 2 blocks
    (wide-round kernel, compiled and run-time isoform counts) after a first pass of the general cluster kernel."""
    monkeypatch.setenv("DICEGP_STRAGGLER_STEPS", "120")
    genes = [560, 561]
    inputs = [synth.gene_matrices("timeseries", 560, reads=20000, T=2, C=4),
              synth.gene_matrices("timeseries", 561, reads=9000, T=2, C=12)]
    out = run_pipeline(ctx, inputs, genes, seed=59, M=330, initial=300, gap=30)
    check_main(out, inputs, 330, 300, 30)


def test_cluster_tier_device_stream_step_parity(ctx):
    """A gene whose W (1.3 MB) takes the thread-block-cluster tier, device-stream mode."""
    genes = [530]
    inputs = [synth.gene_matrices("timeseries", 530, reads=20000, T=2, C=4)]
    out = run_pipeline(ctx, inputs, genes, seed=43, M=260, initial=200, gap=50)
    check_variance_bit_exact(out)
    check_main(out, inputs, 260, 200, 50)
