"""Drop-in replacement of `diceseq.models.model_GP` for the sampler path.

`Psi_GP_MH` keeps the reference's signature and return tuple
(diceseq/models/model_GP.py:188-190,395) but runs the Metropolis-Hastings loop on a B200
through libdicegp (include/dicegp.h).  The host keeps what the reference also does on
the host side of the loop: input cleaning (in place), the GP kernel and its numpy
inverse / determinant, and -- because the reference consumes numpy's global legacy
MT19937 state -- the proposal / uniform stream, generated in the reference's exact
interleaving and shipped to the device in chunks that end on the
convergence-check steps.  There is no CPU fallback for the loop itself.  With
`theta2=None` the stream depends on the chain state; that mode is driven step by step from
the host with the read likelihood evaluated on the device (theta2_mode.py).

`normal_pdf`, `GP_K`, `Geweke_Z` are the small public helpers the reference exports next
to the sampler (diceseq/__init__.py:15); they stay host functions.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .packer import clean_inputs, pack_genes

_LOG_2PI = np.log(2 * np.pi)


# ----------------------------------------------------------------------------------------
# public helpers (host)
# ----------------------------------------------------------------------------------------
def GP_K(X, theta):
    """Squared-exponential covariance theta[0]*exp(-(xi-xj)^2 / (2*theta[1])) of the points
    X (reference: model_GP.py:132-154).  Evaluated entry by entry on numpy scalars so the
    matrix has the reference's bits."""
    n = len(X)
    K = np.empty((n, n))
    for i in range(n):
        for j in range(n):
            K[i, j] = theta[0] * np.exp(-0.5 * (X[i] - X[j]) ** 2 / theta[1])
    return K


def normal_pdf(x, mu, cov, log=True):
    """(log) density of a Gaussian; x may be one point (K,) or N points (N,K); scalar mu
    means N scalar points (reference: model_GP.py:55-94).  Returns None when the
    covariance has a non-positive determinant, after printing the reference's warning."""
    mu_arr = np.array(mu)
    if mu_arr.ndim == 0:
        pts = np.array(x).reshape(-1, 1)
    elif np.array(x).ndim <= 1:
        pts = np.array(x).reshape(1, -1)
    else:
        pts = np.array(x)
    pts = pts - mu_arr
    cov = np.array(cov)
    if cov.ndim < 2:
        cov = cov.reshape(-1, 1)
    inv = np.linalg.inv(cov)
    det = np.linalg.det(cov)
    if det <= 0:
        print("Warning: the det of covariance is not positive!")
        return None
    n, k = pts.shape
    const = -(k * _LOG_2PI + np.log(det)) / 2.0
    out = np.array([const - np.dot(np.dot(p, inv), p) / 2.0 for p in pts])
    if not log:
        out = np.exp(out)
    return out[0] if n == 1 else out


def Geweke_Z(X, first=0.1, last=0.5):
    """Geweke convergence Z score of a 1-D trace: |mean(head) - mean(tail)| over the root of
    the summed population variances; None when that root is 0 (model_GP.py:156-185)."""
    n = X.shape[0]
    head, tail = X[:int(first * n)], X[int(last * n):]
    spread = np.sqrt(np.var(head) + np.var(tail))
    if spread == 0:
        return None
    return abs(head.mean() - tail.mean()) / spread


# ----------------------------------------------------------------------------------------
# host-side constants of one chain
# ----------------------------------------------------------------------------------------
def proposal_factor(cov):
    """A such that standard_normal(T) @ A has covariance `cov`, built the way numpy's
    legacy multivariate_normal does (svd), so that a replayed MT19937 stream reproduces
    the reference's proposals bit for bit."""
    _u, s, v = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * v


class ChainConstants:
    """Everything dicegp_chain_desc needs for one Psi_GP_MH call with fixed theta2."""

    def __init__(self, X, C, var, theta1, theta2):
        T = len(X)
        self.T, self.C = T, C
        self.K = GP_K(X, [theta1, theta2])
        self.kinv = np.linalg.inv(self.K)
        det = np.linalg.det(self.K)
        if det <= 0:
            print("Warning: the det of covariance is not positive!")
            # the reference then fails with `TypeError` on `P_now += None` (model_GP.py:288)
            raise TypeError("GP kernel has a non-positive determinant (normal_pdf returned None)")
        self.prior_const = -(T * _LOG_2PI + np.log(det)) / 2.0
        self.prop_factor_K = proposal_factor(self.K)
        self.jump_scale = np.empty(max(C - 1, 0))
        self.jump_const = np.empty(max(C - 1, 0))
        self.factors = []
        for c in range(C - 1):
            cov_jmp = self.K * var[c] * 5 / (T * C * theta1)          # model_GP.py:328
            jdet = np.linalg.det(cov_jmp)
            if jdet <= 0:
                print("Warning: the det of covariance is not positive!")
                raise TypeError("jump covariance has a non-positive determinant")
            self.jump_scale[c] = var[c] * 5 / (T * C * theta1)
            self.jump_const[c] = -(T * _LOG_2PI + np.log(jdet)) / 2.0
            self.factors.append(proposal_factor(cov_jmp))


def legacy_stream_chunk(n_steps, factors, T, rng=np.random):
    """n_steps of the reference's random stream from numpy's legacy generator: per step
    (C-1) draws of standard_normal(T), each turned into a proposal increment z @ A_c like
    multivariate_normal does, then one uniform (model_GP.py:329,351)."""
    Cm = len(factors)
    delta = np.empty((n_steps, Cm, T))
    u = np.empty(n_steps)
    for s in range(n_steps):
        for c in range(Cm):
            z = rng.standard_normal(T)
            delta[s, c, :] = np.dot(z.reshape(1, -1), factors[c])[0]
        u[s] = rng.random_sample()
    return delta, u


def check_steps(M, initial, gap):
    """Chunk ends: one past every step at which the reference tests convergence."""
    ends = []
    m = initial if initial % gap == 0 else initial + (gap - initial % gap)
    while m < M:
        ends.append(m + 1)
        m += gap
    if not ends or ends[-1] != M:
        ends.append(M)
    return ends


_contexts = {}


def get_context(device=0):
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = _lib.Context(device)
    return ctx


# ----------------------------------------------------------------------------------------
# the sampler
# ----------------------------------------------------------------------------------------
def Psi_GP_MH(R_mat, len_isos, prob_isos, X=None, Ymean=None, var=None, theta1=3.0,
              theta2=None, M=20000, initial=1000, gap=500, randomS=None, theta2_std=1.0,
              theta2_low=0.00001, theta2_up=100, device=0, max_chunk=2048):
    """Isoform proportions of C isoforms at T time points by MH sampling under a GP prior.

    Arguments and return value as in the reference (model_GP.py:188-244):
    returns (Psi_all (m,C,T), Y_all (m,C,T), theta2_all (m,C-1), Pro_all (m,), Lik_all (m,),
    cnt, m).  `R_mat`, `len_isos`, `prob_isos` are cleaned in place, numpy's global random
    state is consumed exactly as the reference consumes it (and reseeded when `randomS`
    is given).  Extra keyword arguments: `device` (GPU index), `max_chunk` (stream steps
    per host->device transfer).
    """
    T = len(len_isos)
    C = len(len_isos[0])
    if X is None:
        X = np.arange(T)
    if Ymean is None:
        Ymean = np.zeros((C, T))
    if randomS is not None:
        np.random.seed(randomS)
    clean_inputs(R_mat, len_isos, prob_isos)
    if var is None:
        var = 0.05 * np.ones(C - 1)
    M, initial, gap = int(M), int(initial), int(gap)
    if theta2 is None:
        # theta2 is sampled too (model_GP.py:317-326): state-dependent stream, host-driven
        # steps with the read likelihood on the device -- see theta2_mode.py
        from .theta2_mode import sample_with_theta2
        W_list = [R_mat[t] * prob_isos[t] for t in range(T)]
        ctx = get_context(device)
        ctx.upload_genes(pack_genes([(W_list, [np.asarray(l, dtype=np.float64) for l in len_isos])]))
        return sample_with_theta2(ctx, T, C, X, Ymean, var, theta1, M, initial, gap, theta2_std,
                                  theta2_low, theta2_up)

    consts = ChainConstants(X, C, var, theta1, theta2)
    W_list = [R_mat[t] * prob_isos[t] for t in range(T)]
    packed = pack_genes([(W_list, [np.asarray(l, dtype=np.float64) for l in len_isos])])
    ctx = get_context(device)
    ctx.upload_genes(packed)

    TT, Cm = T * T, C - 1
    pool = np.concatenate([consts.kinv.ravel(), consts.prop_factor_K.ravel(), consts.jump_scale,
                           consts.jump_const, np.asarray(Ymean, dtype=np.float64).ravel()])
    d = _lib.ChainDesc(gene=0, t0=0, Tc=T, M=M, initial=initial, gap=gap, kinv_off=0,
                       prop_factor_off=TT, jump_scale_off=2 * TT, jump_const_off=2 * TT + Cm,
                       y0_off=2 * TT + 2 * Cm, prior_const=float(consts.prior_const))
    ctx.set_chains([d], pool)

    start = 0
    status = None
    for end in check_steps(M, initial, gap):
        while start < end:
            n = min(end - start, max_chunk)
            delta, u = legacy_stream_chunk(n, consts.factors, T)
            ctx.run_host_stream([0], n, delta.ravel(), [0], u)
            start += n
        status = ctx.chain_status([0])
        if status["state"][0] != _lib.CHAIN_RUNNING:
            break
    m = int(status["m"][0])
    n_rows = int(status["n_rows"][0])
    Psi_all, Y_all, Pro_all, Lik_all = ctx.download_trace(0, n_rows)
    theta2_all = np.full((n_rows, Cm), float(theta2))
    return Psi_all, Y_all, theta2_all, Pro_all, Lik_all, int(status["cnt"][0]), m
