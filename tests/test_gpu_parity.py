"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden
vectors.  Needs a B200: every test is marked gpu.

Tolerance (BASELINE.json north_star): per-step log-probability / log-likelihood within
1e-12 RELATIVE in fp64 and identical accept decisions when fed the identical proposal
stream; psi / y traces within 1e-12.  m and cnt are integers and must be equal.
"""
import os

import numpy as np
import pytest

from conftest import load_golden
from diceseq_b200 import Psi_GP_MH, synth
from oracle import psi_gp_mh_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_close(a, b, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    with np.errstate(invalid="ignore"):
        same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
        ok = same_inf | (np.isnan(a) & np.isnan(b)) | \
            (np.abs(a - b) <= RTOL * np.maximum(np.abs(a), np.abs(b)) + 1e-300)
    if not ok.all():
        i = np.argmin(ok.ravel())
        raise AssertionError("%s: first mismatch at flat index %d: %r vs %r (%d bad of %d)" % (
            what, i, a.ravel()[i], b.ravel()[i], (~ok).sum(), ok.size))


def compare(got, ref):
    assert got[6] == ref[6], ("m", got[6], ref[6])
    assert got[5] == ref[5], ("cnt", got[5], ref[5])
    rel_close(got[3], ref[3], "Pro_all")
    rel_close(got[4], ref[4], "Lik_all")
    rel_close(got[0], ref[0], "Psi_all")
    rel_close(got[1], ref[1], "Y_all")
    rel_close(got[2], ref[2], "theta2_all")


GOLDEN = ["static_c3", "joint_t3_c2", "ts_t8_c4", "prerun_c5", "c9_y0", "edge_nan", "zero_row"]


@pytest.mark.parametrize("name", GOLDEN)
def test_wrapper_matches_reference_golden(name):
    R, L, P, kw, z = load_golden(name)
    got = Psi_GP_MH(R, L, P, **kw)
    ref = (z["Psi_all"], z["Y_all"], z["theta2_all"], z["Pro_all"], z["Lik_all"], int(z["cnt"]), int(z["m"]))
    compare(got, ref)


def test_wrapper_mutates_inputs_like_reference():
    R, L, P, kw, _z = load_golden("edge_nan")
    R2, L2, P2 = [r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P]
    Psi_GP_MH(R, L, P, **kw)
    orc.clean_inputs(R2, L2, P2)
    for t in range(len(R)):
        assert np.array_equal(R[t], R2[t]) and np.array_equal(P[t], P2[t]) and np.array_equal(L[t], L2[t])


def test_global_rng_state_consumed_like_reference():
    R, L, P, kw, _z = load_golden("static_c3")
    Psi_GP_MH(R, L, P, **kw)
    after_gpu = np.random.random_sample()
    R, L, P, kw, _z = load_golden("static_c3")
    orc.psi_gp_mh(R, L, P, **kw)
    assert after_gpu == np.random.random_sample()


# (config, gene, reads, T, C, M, initial, gap): shapes that land in each launch class
# W bytes = 8*C*T*reads: 24 KB (8/SM), 48 KB (4/SM), 96 KB (2/SM), 192 KB (1/SM), 512 KB (streamed)
SHAPES = [
    ("static", 21, 1000, 1, 3, 1300, 1000, 100),
    ("timeseries", 22, 1000, 3, 2, 1200, 1000, 100),
    ("timeseries", 23, 999, 4, 3, 1100, 1000, 100),
    ("timeseries", 24, 1000, 8, 3, 1100, 1000, 100),
    ("timeseries", 25, 1000, 8, 8, 300, 200, 100),
    ("timeseries", 26, 1501, 2, 12, 300, 200, 100),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "T%dC%dN%d" % (s[3], s[4], s[2]))
def test_step_parity_with_oracle(shape):
    cfg, g, reads, T, C, M, ini, gap = shape
    R, L, P = synth.gene_matrices(cfg, g, reads=reads, T=T, C=C)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=M, initial=ini, gap=gap,
              randomS=100 + g)
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    compare(got, ref)


@pytest.mark.parametrize("g", [20, 56, 94, 146, 149, 184])
def test_low_acceptance_rounds_across_trace_pages(g):
    """Wide proposals (acceptance ~0.15) put the kernel into its rejection-chain rounds, which
    commit 1-3 steps each; the run crosses the 1024-row trace pages at every alignment and
    the convergence checks read those rows back."""
    rs = np.random.RandomState(g)
    T = int(rs.choice([1, 3, 4, 8])); C = int(rs.choice([2, 3, 4, 6, 8])); reads = int(rs.choice([200, 999, 2000]))
    R, L, P = synth.gene_matrices("timeseries", g, reads=reads, T=T, C=C)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=2500, initial=1000, gap=100,
              randomS=7 + g, var=0.15 * np.ones(C - 1))
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    assert ref[5] < 0.35 * ref[6]
    compare(got, ref)


def test_pipeline_prerun_variance_and_main(golden):
    """get_psi's sampler logic (pre-runs -> var -> joint run) with the product sampler
    plugged into the oracle's driver restatement reproduces the reference run."""
    for name in ("pipeline_t3_c3", "pipeline_t1_c2"):
        R, L, P, _kw, z = load_golden(name)
        np.random.seed(int(z["seed"]))
        res, var, _ = orc.prerun_and_main(R, L, P, z["X"], float(z["theta1"]), float(z["theta2"]),
                                          int(z["M"]), int(z["initial"]), int(z["gap"]),
                                          sampler=lambda *a, **k: Psi_GP_MH(*a, **k))
        rel_close(var, z["var"], "var")
        assert res[6] == int(z["m"]) and res[5] == int(z["cnt"])
        rel_close(res[3], z["Pro_all"], "Pro_all")


def test_no_tma_staging_gives_identical_bits():
    """DICEGP_NO_TMA=1 stages W with plain loads; same arithmetic, same bits."""
    from diceseq_b200 import model_GP
    R, L, P, kw, _z = load_golden("joint_t3_c2")
    a = Psi_GP_MH(R, L, P, **kw)
    old = model_GP._contexts.pop(0, None)
    os.environ["DICEGP_NO_TMA"] = "1"
    try:
        R, L, P, kw, _z = load_golden("joint_t3_c2")
        b = Psi_GP_MH(R, L, P, **kw)
    finally:
        del os.environ["DICEGP_NO_TMA"]
        model_GP._contexts.pop(0).close()
        if old is not None:
            model_GP._contexts[0] = old
    assert a[6] == b[6] and a[5] == b[5]
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[0], b[0])


def test_small_chunks_equal_large_chunks():
    """State save/restore between launches does not change a bit."""
    R, L, P, kw, _z = load_golden("joint_t3_c2")
    a = Psi_GP_MH(R, L, P, **kw)
    R, L, P, kw, _z = load_golden("joint_t3_c2")
    b = Psi_GP_MH(R, L, P, max_chunk=37, **kw)
    assert a[6] == b[6] and a[5] == b[5]
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[1], b[1])


def test_large_gene_takes_cluster_tier_and_matches_oracle():
    """Human-scale tail (BASELINE config 5): 20 isoforms, 60k reads at one of 2 time points
    (W = 10 MB -> cluster of 4 blocks by default), generic-C kernel with plain global loads."""
    rng = np.random.RandomState(77)
    R, L, P = synth.gene_matrices("timeseries", 27, reads=60000, T=1, C=20)
    R2, L2, P2 = synth.gene_matrices("timeseries", 28, reads=3001, T=1, C=20)
    R, L, P = R + R2, L + L2, P + P2
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(2))), M=60, initial=40, gap=10,
              var=0.002 * np.ones(19), randomS=5)
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    compare(got, ref)
    assert got[5] > 0


@pytest.mark.parametrize("g", [3, 4, 5])
def test_bias_weighted_config_step_parity(g):
    """BASELINE config 4: per-read bias weights folded into prob (bias_utils.py:140-173 feeds
    tran_utils.py:181-186); isoform count as the config draws it, 8 time points."""
    R, L, P = synth.gene_matrices("bias", g, reads=400)
    T = len(R)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=500, initial=300, gap=100,
              randomS=400 + g)
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    compare(got, ref)


@pytest.mark.parametrize("g", [0, 1, 6, 10])
def test_human_scale_config_step_parity(g):
    """BASELINE config 5: 4 time points with heavy-tailed, ragged read counts (46 ... 96k reads
    per time point in these genes) and 2-10 isoforms, exactly as the config draws them."""
    R, L, P = synth.gene_matrices("human", g)
    T = len(R)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=260, initial=200, gap=50,
              randomS=500 + g)
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    compare(got, ref)


def test_theta2_sampling_mode_matches_reference_golden():
    """`theta2=None` (model_GP.py:317-326): host-driven steps, read likelihood on the device.
    Same m / cnt / theta2 trace as the unmodified reference, traces within 1e-12."""
    R, L, P, kw, z = load_golden("learn_c2")
    got = Psi_GP_MH(R, L, P, **kw)
    ref = (z["Psi_all"], z["Y_all"], z["theta2_all"], z["Pro_all"], z["Lik_all"], int(z["cnt"]), int(z["m"]))
    compare(got, ref)


def test_theta2_sampling_mode_fails_like_reference_for_three_isoforms():
    """The reference leaves cov_try[:, :, 1:] at zero in this mode and dies in np.linalg.inv
    (SURVEY.md Appendix B); same exception here."""
    R, L, P = synth.gene_matrices("timeseries", 8, reads=50, T=2, C=3)
    with pytest.raises(np.linalg.LinAlgError):
        Psi_GP_MH(R, L, P, theta1=3.0, theta2=None, M=50, initial=30, gap=10, randomS=1)


def test_eval_loglik_matches_numpy_on_ragged_large_gene():
    """dicegp_eval_loglik on a human-scale gene (ragged read counts, several blocks per time
    point) and on a state with a zero likelihood row (-inf like numpy)."""
    from diceseq_b200.model_GP import get_context
    from diceseq_b200.packer import pack_genes
    W, lens = synth.gene_W("human", 0)
    ctx = get_context(0)
    ctx.upload_genes(pack_genes([(W, lens)]))
    rng = np.random.RandomState(3)
    C, T = W[0].shape[1], len(W)
    for _ in range(3):
        Y = rng.normal(0, 1.5, size=(C, T))
        Y[-1] = 0
        got = ctx.eval_loglik(0, Y)
        for t in range(T):
            e = np.exp(Y[:, t]); psi = e / e.sum()
            f = lens[t] * psi / np.sum(lens[t] * psi)
            ref = np.log(W[t] @ f).sum()
            assert abs(got[t] - ref) <= 1e-12 * abs(ref)
    # a sub-range of the time points
    got = ctx.eval_loglik(0, Y[:, 1:3], t0=1)
    full = ctx.eval_loglik(0, Y)
    assert np.allclose(got, full[1:3], rtol=1e-13, atol=0)      # (block partition may differ)
    Y[:, 0] = [800.0] + [0.0] * (C - 1)         # exp overflows: nan like the reference's softmax
    with np.errstate(all="ignore"):
        assert not np.isfinite(ctx.eval_loglik(0, Y)[0])
