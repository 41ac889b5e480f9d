"""
ORACLE -- test infrastructure, NOT product code.

CPU (numpy) restatement of DICEseq's Metropolis-Hastings Gaussian-process isoform
sampler and of the sampler-related maths of its per-gene driver.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this file; the product (`diceseq_b200/`) never does.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference
(`/root/reference/diceseq/models/model_GP.py`, loaded by path because the package
`__init__` needs pysam/matplotlib) in the build container, runs it with fixed seeds and
stores inputs + outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
this file against those vectors (bit-exact on the machine that generated them).

Every function cites the reference lines it restates (paths relative to
/root/reference/).  The arithmetic is kept call-for-call identical to the reference
where the numpy call decides the bits (strided column views into `np.exp`, `np.dot`
on a (1,T)x(T,T) pair for the proposal, left-to-right evaluation of the MH exponent),
but loop-invariant linear algebra (inverse / determinant / SVD of matrices that do not
change between steps) is hoisted out of the step loop: numpy is deterministic, so the
hoisted value has the same bits as the per-step recomputation.  `faithful_cost=True`
re-does that linear algebra every step exactly like the reference does, so that a
timing of this port costs what the reference costs (used by bench.py's reference arm).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "clean_inputs", "gp_kernel", "log_mvn_pdf", "log_trunc_normal_pdf", "erf_as",
    "geweke_z", "softmax_len_reweight", "read_loglik", "proposal_factor",
    "LegacyStream", "StepRecord", "psi_gp_mh", "prerun_and_main", "credible_interval",
    "harmonic_mean_log", "posterior_summary",
]


# --------------------------------------------------------------------------------------
# small pdf helpers
# --------------------------------------------------------------------------------------
def erf_as(x):
    """Abramowitz & Stegun 7.1.26 erf, |err| <= 1.5e-7.  diceseq/models/model_GP.py:7-24."""
    sgn = 1 if x >= 0 else -1
    x = abs(x)
    a1, a2, a3, a4, a5 = 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429
    p = 0.3275911
    t = 1.0 / (1.0 + p * x)
    y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * np.exp(-x * x)
    return sgn * y


def log_trunc_normal_pdf(x, mu, sigma, a, b):
    """log density of N(mu, sigma) truncated to [a, b].  model_GP.py:96-130 (log=True)."""
    x = x - mu
    a = a - mu
    b = b - mu
    pdf = np.exp(-0.5 * (x / sigma) ** 2) / (sigma * np.sqrt(2 * np.pi))
    cdf_a = (1 + erf_as(a / sigma / np.sqrt(2))) / 2.0
    cdf_b = (1 + erf_as(b / sigma / np.sqrt(2))) / 2.0
    pdf = pdf / abs(cdf_b - cdf_a)
    return np.log(pdf)


class _MvnConst:
    """inverse and log-normaliser of one covariance; model_GP.py:81-89."""
    __slots__ = ("inv", "part1", "ok")

    def __init__(self, cov):
        cov = np.asarray(cov)
        if cov.ndim < 2:
            cov = cov.reshape(-1, 1)
        self.inv = np.linalg.inv(cov)                      # :83
        det = np.linalg.det(cov)                           # :84
        self.ok = bool(det > 0)                            # :85-87 (reference returns None)
        k = cov.shape[0]
        self.part1 = -(k * np.log(2 * np.pi) + np.log(det)) / 2.0 if self.ok else None  # :89

    def logpdf_centered(self, d):
        """d = x - mu (1-D).  model_GP.py:91."""
        if not self.ok:
            return None
        return self.part1 - np.dot(np.dot(d, self.inv), d) / 2.0


def log_mvn_pdf(x, mu, cov):
    """log N(x; mu, cov) for one 1-D x.  model_GP.py:55-94 (the N==1, log=True branch that
    Psi_GP_MH uses).  Returns None when det(cov) <= 0 like the reference (:85-87)."""
    d = np.array(x).reshape(1, -1) - np.array(mu)          # :77-79
    return _MvnConst(cov).logpdf_centered(d[0, :])


def gp_kernel(X, theta):
    """Squared-exponential kernel K[i,j] = theta0 * exp(-0.5 (Xi-Xj)^2 / theta1).
    model_GP.py:132-154 (scalar double loop kept: np.exp on a scalar and on an array may
    round differently)."""
    n = len(X)
    K = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            K[i, j] = theta[0] * np.exp(-0.5 * (X[i] - X[j]) ** 2 / theta[1])
    return K


def geweke_z(series, first=0.1, last=0.5):
    """Geweke Z score of one trace; None if both windows have zero variance.
    model_GP.py:156-185."""
    n = series.shape[0]
    head = series[:int(first * n)]
    tail = series[int(last * n):]
    den = np.sqrt(np.var(head) + np.var(tail))
    if den == 0:
        return None
    return abs(head.mean() - tail.mean()) / den


# --------------------------------------------------------------------------------------
# per-step pieces
# --------------------------------------------------------------------------------------
def clean_inputs(R_mat, len_isos, prob_isos):
    """In-place input cleaning, model_GP.py:250-264.  Mutates the arrays and rebinds the
    list slots exactly like the reference (callers see the filtered matrices)."""
    for t in range(len(len_isos)):
        bad_len = (len_isos[t] != len_isos[t])              # NaN lengths, :251
        len_isos[t][bad_len] = 0.0
        prob_isos[t][:, bad_len] = 0.0
        R_mat[t][:, bad_len] = False
        R_mat[t][np.where(R_mat[t] != R_mat[t])] = False     # :256-257
        prob_isos[t][np.where(prob_isos[t] != prob_isos[t])] = 0.0   # :259-260
        keep = (R_mat[t].sum(axis=1) > 0) * (prob_isos[t].sum(axis=1) > 0)   # :262
        R_mat[t] = R_mat[t][keep, :]
        prob_isos[t] = prob_isos[t][keep, :]


def softmax_len_reweight(Y, len_isos, psi_out, fsi_out, t):
    """psi[:,t] = softmax(Y[:,t]) without max-subtraction; fsi = len*psi / sum.
    model_GP.py:281-282 and :335-337.  Works on column views like the reference."""
    psi_out[:, t] = np.exp(Y[:, t]) / np.sum(np.exp(Y[:, t]))
    fsi_out[:, t] = len_isos[t] * psi_out[:, t] / np.sum(len_isos[t] * psi_out[:, t])


def read_loglik(W_t, fsi_col):
    """sum_n log(sum_k W[n,k] f[k]) for one time point.  model_GP.py:338,344."""
    return np.log(np.dot(W_t, fsi_col)).sum()


def proposal_factor(cov_jmp, psd_check=False):
    """Matrix A with multivariate_normal(mean, cov) == standard_normal(T) @ A + mean
    (numpy legacy: svd-based).  SURVEY Appendix C; numpy mtrand.multivariate_normal.
    `psd_check` repeats numpy's check_valid='warn' test (cost only; result unused)."""
    _u, s, v = np.linalg.svd(cov_jmp)
    if psd_check:
        np.allclose(np.dot(v.T * s, v), cov_jmp, rtol=1e-8, atol=1e-8)
    return np.sqrt(s)[:, None] * v


class LegacyStream:
    """Draws from a legacy MT19937 RandomState in the interleaving the reference uses:
    per step (C-1) x standard_normal(T) then one random_sample (model_GP.py:329,351)."""

    def __init__(self, rng=None):
        # rng None -> numpy's global legacy state, which is what the reference consumes
        self.rng = np.random.mtrand._rand if rng is None else rng

    def seed(self, s):
        self.rng.seed(s)

    def normals(self, n):
        return self.rng.standard_normal(n)

    def normal(self, loc, scale):
        return self.rng.normal(loc, scale)

    def uniform(self):
        return self.rng.random_sample()


class StepRecord:
    """Optional per-step recorder for parity tests (delta = proposal minus current)."""

    def __init__(self):
        self.delta, self.u, self.accept, self.P_try, self.L_try = [], [], [], [], []

    def arrays(self):
        return (np.array(self.delta), np.array(self.u), np.array(self.accept, dtype=bool),
                np.array(self.P_try), np.array(self.L_try))


def psi_gp_mh(R_mat, len_isos, prob_isos, X=None, Ymean=None, var=None, theta1=3.0,
              theta2=None, M=20000, initial=1000, gap=500, randomS=None, theta2_std=1.0,
              theta2_low=0.00001, theta2_up=100, *, rng=None, record=None,
              faithful_cost=False, replay=None):
    """MH sampler with GP prior, restating model_GP.py:188-395.

    Same arguments / return tuple as the reference:
    (Psi_all (m,C,T), Y_all (m,C,T), theta2_all (m,C-1), Pro_all (m,), Lik_all (m,), cnt, m).
    Extra keyword-only arguments: `rng` (a RandomState; default the global legacy state),
    `record` (a StepRecord), `faithful_cost` (redo inv/det/svd every step like the
    reference; same bits, reference-like cost), `replay` = (delta (M, C-1, T), u (M,)): take the
    proposal increments (the value of `np.dot(z, A)` in numpy's multivariate_normal, :329) and the
    uniforms (:351) from these arrays instead of drawing them -- used to step-check a stream that
    was generated elsewhere (the device's Philox stream); fixed-theta2 mode only.
    """
    T = len(len_isos)
    C = len(len_isos[0])
    if X is None:
        X = np.arange(T)
    if Ymean is None:
        Ymean = np.zeros((C, T))
    stream = LegacyStream(rng)
    if randomS is not None:
        stream.seed(randomS)                                   # :249
    clean_inputs(R_mat, len_isos, prob_isos)                   # :250-264

    learn = theta2 is None
    if var is None:
        var = 0.05 * np.ones(C - 1)                            # :267-268
    theta_now = np.zeros((C - 1, 2))
    theta_now[:, 0] = theta1
    if not learn:
        theta_now[:, 1] = theta2
    else:
        theta_now[:, 1] = 0.1 * (np.max(T) - np.min(T) + 0.001) ** 2   # :274 (quirk: T is an int)
    Y_now = Ymean + 0.0
    zero_mean = np.zeros((C, T))                               # :276 prior mean forced to 0

    W = [R_mat[t] * prob_isos[t] for t in range(T)]           # re-materialised per step in ref (:338)

    psi_now = np.zeros((C, T))
    fsi_now = np.zeros((C, T))
    for t in range(T):
        softmax_len_reweight(Y_now, len_isos, psi_now, fsi_now, t)     # :280-282

    P_now, L_now = 0, 0
    cov_now = np.zeros((T, T, C - 1))
    for c in range(C - 1):
        cov_now[:, :, c] = gp_kernel(X, theta_now[c, :])       # :287
        P_now += log_mvn_pdf(Y_now[c, :], zero_mean[c, :], cov_now[:, :, c])   # :288
    for t in range(T):
        lt = read_loglik(W[t], fsi_now[:, t])                  # :291-292 (evaluated twice in ref)
        P_now += lt
        L_now += lt

    Y_try = np.zeros((C, T))
    Y_all = np.zeros((M, C, T))
    psi_try = np.zeros((C, T))
    fsi_try = np.zeros((C, T))
    Psi_all = np.zeros((M, C, T))
    cov_try = np.zeros((T, T, C - 1))
    theta_try = np.zeros((C - 1, 2))
    theta2_all = np.zeros((M, C - 1))
    theta_try[:, 0] = theta1
    prior_c = jump_c = fac_c = None
    if not learn:
        theta_try[:, 1] = theta2
        cov_try[:, :, :] = gp_kernel(X, theta_try[0, :]).reshape(T, T, 1)      # :306
        # hoisted loop invariants (fixed-theta2 mode): same bits as per-step recomputation
        prior_c = [_MvnConst(cov_try[:, :, c]) for c in range(C - 1)]
        jump_cov = [cov_try[:, :, c] * var[c] * 5 / (T * C * theta1) for c in range(C - 1)]   # :328
        jump_c = [_MvnConst(jc) for jc in jump_cov]
        fac_c = [proposal_factor(jc) for jc in jump_cov]

    cnt = 0
    Pro_all = np.zeros(M)
    Lik_all = np.zeros(M)
    m = -1
    for m in range(M):
        P_try, L_try, Q_now, Q_try = 0, 0, 0, 0
        for c in range(C - 1):
            if learn and c == 0:                               # :317-326
                theta_try[:, 1] = stream.normal(theta_now[c, 1], theta2_std)
                while theta_try[c, 1] < theta2_low or theta_try[c, 1] > theta2_up:
                    theta_try[:, 1] = stream.normal(theta_now[c, 1], theta2_std)
                cov_try[:, :, c] = gp_kernel(X, theta_try[c, :])
                Q_now += log_trunc_normal_pdf(theta_now[c, 1], theta_try[c, 1],
                                              theta2_std, theta2_low, theta2_up)
                Q_try += log_trunc_normal_pdf(theta_try[c, 1], theta_now[c, 1],
                                              theta2_std, theta2_low, theta2_up)
            if learn or faithful_cost:
                cov_jmp = cov_try[:, :, c] * var[c] * 5 / (T * C * theta1)     # :328
                A = proposal_factor(cov_jmp, psd_check=faithful_cost)
                if faithful_cost:
                    # the reference does inv+det three times per isoform per step (:330-332)
                    _MvnConst(cov_jmp)
                jc = _MvnConst(cov_jmp)
                pc = _MvnConst(cov_try[:, :, c])
            else:
                A, jc, pc = fac_c[c], jump_c[c], prior_c[c]
            if replay is not None:
                step = np.array(replay[0][m, c], dtype=float).reshape(1, -1)
            else:
                z = stream.normals(T)
                step = np.dot(z.reshape(1, -1), A)             # numpy multivariate_normal body
            if record is not None:
                record.delta.append(step[0].copy())
            step += Y_now[c, :]
            Y_try[c, :] = step[0]                              # :329
            d = np.array(Y_now[c, :]).reshape(1, -1) - np.array(Y_try[c, :])
            Q_now += jc.logpdf_centered(d[0, :])               # :330
            d = np.array(Y_try[c, :]).reshape(1, -1) - np.array(Y_now[c, :])
            Q_try += jc.logpdf_centered(d[0, :])               # :331
            d = np.array(Y_try[c, :]).reshape(1, -1) - np.array(zero_mean[c, :])
            P_try += pc.logpdf_centered(d[0, :])               # :332

        for t in range(T):
            softmax_len_reweight(Y_try, len_isos, psi_try, fsi_try, t)   # :335-337
            if faithful_cost:                                  # ref: R*prob temp + log-sum twice
                read_loglik(R_mat[t] * prob_isos[t], fsi_try[:, t])
                lt = read_loglik(R_mat[t] * prob_isos[t], fsi_try[:, t])
            else:
                lt = read_loglik(W[t], fsi_try[:, t])          # :338,:344-345
            P_try += lt
            L_try += lt

        alpha = np.exp(min(P_try + Q_now - P_now - Q_try, 0))  # :348
        u = replay[1][m] if replay is not None else stream.uniform()   # :351 (np.random.rand(1))
        accepted = bool(u < alpha)
        if record is not None:
            record.u.append(u)
            record.accept.append(accepted)
            record.P_try.append(P_try)
            record.L_try.append(L_try)
        if accepted:                                           # :351-360
            cnt += 1
            P_now = P_try + 0.0
            L_now = L_try + 0.0
            Y_now = Y_try + 0.0
            cov_now = cov_try + 0.0
            psi_now = psi_try + 0.0
            fsi_now = fsi_try + 0.0
            theta_now = theta_try + 0.0

        Pro_all[m] = P_now                                     # :362-366
        Lik_all[m] = L_now
        Y_all[m, :, :] = Y_now
        Psi_all[m, :, :] = psi_now
        theta2_all[m, :] = theta_now[:, 1]

        if m >= initial and m % gap == 0:                      # :369-391
            conv = 1
            for c in range(C - 1):
                for t in range(T):
                    Z = geweke_z(Psi_all[:m, c, t])
                    if Z is None or Z > 2:
                        conv = 0
                        break
                if learn:
                    Z = geweke_z(theta2_all[:m, c])
                    if Z is None or Z > 2:
                        conv = 0
                        break
                if conv == 0:
                    break
            if conv == 1:
                Pro_all = Pro_all[:m, ]
                Lik_all = Lik_all[:m, ]
                Y_all = Y_all[:m, :, :]
                Psi_all = Psi_all[:m, :, :]
                theta2_all = theta2_all[:m, :]
                break
    return Psi_all, Y_all, theta2_all, Pro_all, Lik_all, cnt, m          # :395


# --------------------------------------------------------------------------------------
# caller-side maths (diceseq/utils/run_utils.py)
# --------------------------------------------------------------------------------------
def credible_interval(data, percent=0.95):
    """Per-column [upper, lower] order statistics, run_utils.py:14-22 (note the
    `temp[-0]` quirk when fewer than 40 rows: the 'upper' bound is then the minimum)."""
    if len(data.shape) == 0:
        data = data.reshape(-1, 1)
    out = np.zeros((data.shape[1], 2))
    idx = int(data.shape[0] * (1 - percent) / 2)
    for k in range(data.shape[1]):
        srt = np.sort(data[:, k])
        out[k, :] = [srt[-idx], srt[idx]]
    return out


def harmonic_mean_log(data):
    """log harmonic mean of exp(data), run_utils.py:24-30 (log_in=log_out=True)."""
    rv = np.log(len(data)) + np.min(data, axis=0)
    rv = rv - np.log(np.sum(np.exp(np.min(data, axis=0) - data), axis=0))
    return rv


def prerun_and_main(R_all, len_iso_all, prob_iso_all, X, theta1, theta2, M, initial, gap,
                    *, rng=None, faithful_cost=False, sampler=None):
    """The sampler part of get_psi, run_utils.py:106-122: T single-time-point pre-runs
    (M=100, initial=100, gap=50) that set the jump variance, then the joint run.
    Returns (main result tuple, var, theta2_std)."""
    run = sampler or (lambda *a, **k: psi_gp_mh(*a, rng=rng, faithful_cost=faithful_cost, **k))
    C = R_all[0].shape[1]
    var = 0.1 * np.ones(C - 1)                                  # :107
    Ymean = np.zeros((C, len(X)))                               # :108
    _var = np.zeros(C - 1)
    for j in range(len(R_all)):                                 # :111-115
        _Psi, _Y, _theta2, _Pro, _Lik, _cnt, _m = run(
            R_all[j:j + 1], len_iso_all[j:j + 1], prob_iso_all[j:j + 1], X[j:j + 1],
            Ymean[:, j:j + 1], var, theta1, theta2, 100, 100, 50)
        _var += np.var(_Y[int(_m / 4):, :-1, 0], axis=0)
    var = (_var + 0.000001) / len(R_all)                        # :116
    theta2_std = np.var(_theta2[int(_m / 4):, 0]) / 10.0 + 0.001   # :117
    res = run(R_all, len_iso_all, prob_iso_all, X, Ymean, var, theta1, theta2, M, initial,
              gap, theta2_std=theta2_std)                       # :120-122
    return res, var, theta2_std


def posterior_summary(Psi, Y, Lik, m, sample_num=0):
    """Burn-in m/4 summaries of get_psi, run_utils.py:125-131."""
    b = int(m / 4)
    psi_95 = [credible_interval(Psi[b:, c, :], 0.95) for c in range(Psi.shape[1])]
    psi_mean = Psi[b:, :, :].mean(axis=0)
    lik_marg = harmonic_mean_log(Lik[b:])
    sample_all = Y[-sample_num:, :-1, :]
    return psi_mean, psi_95, lik_marg, sample_all
