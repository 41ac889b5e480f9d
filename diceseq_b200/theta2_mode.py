"""`Psi_GP_MH(theta2=None)`: the chain that also samples the GP length scale theta2
(reference: diceseq/models/model_GP.py:317-326, reached with `--thetas x learn`).

Why this mode is driven from the host, one step at a time:
  * theta2 is proposed by a truncated random walk implemented as a REJECTION LOOP on
    `np.random.normal` (model_GP.py:318-320), so the number of normals a step consumes
    from numpy's legacy MT19937 state depends on the current theta2, i.e. on the accept
    history -- the stream cannot be generated ahead of the chain like in fixed-theta2 mode;
  * the proposal covariance K(theta2_try)*s changes every step and numpy's
    `multivariate_normal` factorises it with a LAPACK SVD (model_GP.py:329); bit-level
    step parity needs exactly those singular vectors.
So the host keeps what the reference does with numpy/LAPACK (theta2 walk, GP kernel, SVD
proposal, prior / proposal densities, accept test, Geweke rule) and the B200 evaluates what
scales with the data: the read likelihood of every proposed state
(`dicegp_eval_loglik`, model_GP.py:335-345).  There is no CPU likelihood in this module.

The reference only survives this branch for C = 2 (for C > 2 `cov_try[:, :, c >= 1]` stays
zero and `np.linalg.inv` raises `LinAlgError: Singular matrix`);
the same exception comes out of here, from the same call.
"""
from __future__ import annotations

import numpy as np

# Abramowitz & Stegun 7.1.26 coefficients (the reference's `erf`, model_GP.py:7-24)
_AS_P = 0.3275911
_AS_A = (0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429)


def erf_as(x):
    """erf by A&S 7.1.26 (max error 1.5e-7), evaluated in the reference's operation order."""
    sign = 1 if x >= 0 else -1
    x = abs(x)
    a1, a2, a3, a4, a5 = _AS_A
    t = 1.0 / (1.0 + _AS_P * x)
    poly = (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t
    return sign * (1.0 - poly * np.exp(-x * x))


def trun_normal_logpdf(x, mu, sigma, low, up):
    """log density at x of N(mu, sigma^2) truncated to [low, up] (model_GP.py:96-130)."""
    x, low, up = x - mu, low - mu, up - mu
    dens = np.exp(-0.5 * (x / sigma) ** 2) / (sigma * np.sqrt(2 * np.pi))
    cdf_low = (1 + erf_as(low / sigma / np.sqrt(2))) / 2.0
    cdf_up = (1 + erf_as(up / sigma / np.sqrt(2))) / 2.0
    return np.log(dens / abs(cdf_up - cdf_low))


class _State:
    """Everything `accept` copies (model_GP.py:353-360)."""
    __slots__ = ("P", "L", "Y", "psi", "theta2")

    def __init__(self, P, L, Y, psi, theta2):
        self.P, self.L, self.Y, self.psi, self.theta2 = P, L, Y, psi, theta2


def _softmax_cols(Y):
    psi = np.empty_like(Y)
    for t in range(Y.shape[1]):
        e = np.exp(Y[:, t])
        psi[:, t] = e / np.sum(e)
    return psi


def sample_with_theta2(ctx, T, C, X, Y_start, var, theta1, M, initial, gap, theta2_std, theta2_low, theta2_up):
    """The MH loop of `Psi_GP_MH` in theta2-sampling mode for gene 0 of the batch uploaded to
    `ctx` (a `_lib.Context`).  Consumes numpy's global legacy random state call by call like
    the reference does.  Returns the reference's 7-tuple."""
    from .model_GP import GP_K, Geweke_Z, normal_pdf

    zero = np.zeros(T)
    Cm = C - 1

    def device_lik(Y):
        """P and L both add the per-time-point sums in ascending t (model_GP.py:344-345)."""
        return ctx.eval_loglik(0, Y)

    # ---- initial state (model_GP.py:270-292)
    th0 = 0.1 * (np.max(T) - np.min(T) + 0.001) ** 2           # :274 (T is the int, not X)
    Y0 = np.asarray(Y_start, dtype=np.float64) + 0.0
    P0, L0 = 0, 0
    for c in range(Cm):
        P0 += normal_pdf(Y0[c, :], zero, GP_K(X, [theta1, th0]))
    for lt in device_lik(Y0):
        P0 += lt
        L0 += lt
    now = _State(P0, L0, Y0, _softmax_cols(Y0), th0)

    Psi_all = np.zeros((M, C, T))
    Y_all = np.zeros((M, C, T))
    theta2_all = np.zeros((M, Cm))
    Pro_all = np.zeros(M)
    Lik_all = np.zeros(M)
    cov_try = np.zeros((T, T, Cm))          # only slice 0 is ever filled (:321)
    cnt = 0
    m = -1
    for m in range(M):
        P_try, L_try, Q_now, Q_try = 0, 0, 0, 0
        Y_try = np.zeros((C, T))
        th_try = now.theta2
        for c in range(Cm):
            if c == 0:                                         # :317-326
                th_try = np.random.normal(now.theta2, theta2_std)
                while th_try < theta2_low or th_try > theta2_up:
                    th_try = np.random.normal(now.theta2, theta2_std)
                cov_try[:, :, 0] = GP_K(X, [theta1, th_try])
                Q_now += trun_normal_logpdf(now.theta2, th_try, theta2_std, theta2_low, theta2_up)
                Q_try += trun_normal_logpdf(th_try, now.theta2, theta2_std, theta2_low, theta2_up)
            cov_jmp = cov_try[:, :, c] * var[c] * 5 / (T * C * theta1)            # :328
            Y_try[c, :] = np.random.multivariate_normal(now.Y[c, :], cov_jmp)    # :329
            Q_now += normal_pdf(now.Y[c, :], Y_try[c, :], cov_jmp)               # :330
            Q_try += normal_pdf(Y_try[c, :], now.Y[c, :], cov_jmp)               # :331
            P_try += normal_pdf(Y_try[c, :], zero, cov_try[:, :, c])             # :332
        for lt in device_lik(Y_try):                           # :334-345 on the GPU
            P_try += lt
            L_try += lt

        alpha = np.exp(min(P_try + Q_now - now.P - Q_try, 0))  # :348
        if np.random.rand(1) < alpha:                          # :351
            cnt += 1
            now = _State(P_try + 0.0, L_try + 0.0, Y_try, _softmax_cols(Y_try), th_try)

        Pro_all[m] = now.P                                     # :362-366
        Lik_all[m] = now.L
        Y_all[m] = now.Y
        Psi_all[m] = now.psi
        theta2_all[m, :] = now.theta2

        if m >= initial and m % gap == 0:                      # :369-391
            if _converged(Geweke_Z, Psi_all, theta2_all, m, Cm, T):
                return Psi_all[:m], Y_all[:m], theta2_all[:m], Pro_all[:m], Lik_all[:m], cnt, m
    return Psi_all, Y_all, theta2_all, Pro_all, Lik_all, cnt, m


def _converged(geweke, Psi_all, theta2_all, m, Cm, T):
    """Every psi trace and the theta2 trace of every sampled isoform has Z <= 2."""
    for c in range(Cm):
        for t in range(T):
            z = geweke(Psi_all[:m, c, t])
            if z is None or z > 2:
                return False
        z = geweke(theta2_all[:m, c])
        if z is None or z > 2:
            return False
    return True
