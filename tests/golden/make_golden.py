"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py

Loads /root/reference/diceseq/models/model_GP.py by file path (the package __init__ pulls in
pysam / matplotlib, which are not installed), feeds it seeded synthetic inputs and stores
inputs + outputs as tests/golden/<case>.npz.  The `pipeline_*` cases wrap the reference
sampler in the sampler part of get_psi (diceseq/utils/run_utils.py:106-131: T pre-runs ->
jump variance -> joint run -> burn-in summaries), which cannot be imported itself (pysam).
/root/reference does not exist on the GPU box; only the .npz files travel.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from diceseq_b200 import synth  # noqa: E402

REF = "/root/reference/diceseq/models/model_GP.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_model_GP", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def copy_inputs(R, L, P):
    return [r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P]


def pack_inputs(R, L, P):
    out = {"T": len(R)}
    for t in range(len(R)):
        out["R%d" % t], out["len%d" % t], out["prob%d" % t] = R[t], L[t], P[t]
    return out


def edge_inputs(rng, T=2, C=3):
    """NaN length, NaN probabilities, a dead read, odd and empty time points."""
    R, L, P = [], [], []
    for t, n in zip(range(T), [37, 0][:T]):
        r = rng.rand(n, C) < 0.6
        p = rng.rand(n, C) * 1e-3
        p[~r] = np.nan
        if n:
            r[3, :] = False                      # dead read: dropped by the cleaning
            p[5, :] = np.nan                     # no probability mass: dropped
            r[7, :] = True
        lens = rng.randint(500, 3000, size=C).astype(float)
        R.append(r), L.append(lens), P.append(p)
    L[0][1] = np.nan                              # NaN length -> isoform 1 switched off at t=0
    return R, L, P


def zero_row_inputs(rng, C=2):
    """Appendix B #9: R and prob non-zero in different columns -> W row of zeros -> log 0."""
    n = 20
    r = np.ones((n, C), bool)
    p = rng.rand(n, C) * 1e-3
    r[4, :] = [True, False]
    p[4, :] = [0.0, 1e-3]
    return [r], [np.array([1000.0, 2000.0])], [p]


def run_case(ref, name, R, L, P, kwargs):
    Rc, Lc, Pc = copy_inputs(R, L, P)
    with np.errstate(all="ignore"):
        Psi, Y, th2, Pro, Lik, cnt, m = ref.Psi_GP_MH(Rc, Lc, Pc, **kwargs)
    out = pack_inputs(R, L, P)
    for k, v in kwargs.items():
        out["arg_" + k] = np.array(np.nan if v is None else v)
    out.update(Psi_all=Psi, Y_all=Y, theta2_all=th2, Pro_all=Pro, Lik_all=Lik, cnt=cnt, m=m)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-16s m=%d cnt=%d rows=%d  P[-1]=%.6f" % (name, m, cnt, len(Pro), Pro[-1]))


def run_pipeline(ref, name, R, L, P, X, theta1, theta2, M, initial, gap, seed):
    """run_utils.py:106-131 around the reference sampler (global RNG seeded once)."""
    Rc, Lc, Pc = copy_inputs(R, L, P)
    np.random.seed(seed)
    C, T = Rc[0].shape[1], len(Rc)
    var = 0.1 * np.ones(C - 1)
    Ymean = np.zeros((C, T))
    _var = np.zeros(C - 1)
    with np.errstate(all="ignore"):
        for j in range(T):
            _Psi, _Y, _t2, _Pro, _Lik, _cnt, _m = ref.Psi_GP_MH(
                Rc[j:j + 1], Lc[j:j + 1], Pc[j:j + 1], X[j:j + 1], Ymean[:, j:j + 1], var,
                theta1, theta2, 100, 100, 50)
            _var += np.var(_Y[int(_m / 4):, :-1, 0], axis=0)
        var = (_var + 0.000001) / T
        _Psi, _Y, _t2, _Pro, _Lik, _cnt, _m = ref.Psi_GP_MH(
            Rc, Lc, Pc, X, Ymean, var, theta1, theta2, M, initial, gap)
    b = int(_m / 4)
    psi_mean = _Psi[b:, :, :].mean(axis=0)
    out = pack_inputs(R, L, P)
    out.update(X=X, theta1=theta1, theta2=theta2, M=M, initial=initial, gap=gap, seed=seed,
               var=var, Psi_all=_Psi, Y_all=_Y, Pro_all=_Pro, Lik_all=_Lik, cnt=_cnt, m=_m,
               psi_mean=psi_mean)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-16s m=%d cnt=%d var=%s" % (name, _m, _cnt, np.array2string(var, precision=4)))


def main():
    ref = load_reference()
    X3 = np.array([1.5, 2.5, 5.0])
    # 1. static mode, config-2 shape (fewer reads)
    R, L, P = synth.gene_matrices("static", 0, reads=301)
    run_case(ref, "static_c3", R, L, P, dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(1))),
                                             M=1500, initial=1000, gap=100, randomS=11))
    # 2. joint model like demo.sh (3 time points, 2 isoforms, uneven times)
    R, L, P = synth.gene_matrices("timeseries", 1, reads=200, T=3, C=2)
    run_case(ref, "joint_t3_c2", R, L, P, dict(X=X3, theta1=3.0, theta2=float(synth.default_theta2(X3)),
                                               M=2000, initial=1000, gap=100, randomS=12))
    # 3. time series, 8 points, 4 isoforms
    R, L, P = synth.gene_matrices("timeseries", 2, reads=120, T=8, C=4)
    run_case(ref, "ts_t8_c4", R, L, P, dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(8))),
                                            M=1300, initial=1000, gap=100, randomS=13))
    # 4. a pre-run (never checks convergence: m = 99)
    R, L, P = synth.gene_matrices("timeseries", 3, reads=150, T=1, C=5)
    run_case(ref, "prerun_c5", R, L, P, dict(X=np.array([2]), Ymean=np.zeros((5, 1)), var=0.1 * np.ones(4),
                                             theta1=3.0, theta2=float(synth.default_theta2(np.arange(8))),
                                             M=100, initial=100, gap=50, randomS=14))
    # 5. nine isoforms (generic-C kernel path), non-zero start value
    R, L, P = synth.gene_matrices("timeseries", 4, reads=91, T=2, C=9)
    Y0 = np.random.RandomState(5).randn(9, 2) * 0.3
    run_case(ref, "c9_y0", R, L, P, dict(Ymean=Y0, var=0.3 * np.ones(8), theta1=3.0, theta2=0.7,
                                         M=500, initial=200, gap=100, randomS=15))
    # 6. dirty inputs
    R, L, P = edge_inputs(np.random.RandomState(6))
    run_case(ref, "edge_nan", R, L, P, dict(theta1=3.0, theta2=0.5, M=300, initial=100, gap=50, randomS=16))
    # 7. -inf likelihood: chain never moves, never converges
    R, L, P = zero_row_inputs(np.random.RandomState(7))
    run_case(ref, "zero_row", R, L, P, dict(theta1=3.0, theta2=0.1, M=160, initial=100, gap=50, randomS=17))
    # 8. theta2 sampled (works for C == 2 only in the reference)
    R, L, P = synth.gene_matrices("timeseries", 8, reads=80, T=3, C=2)
    run_case(ref, "learn_c2", R, L, P, dict(theta1=3.0, theta2=None, M=600, initial=300, gap=100,
                                            randomS=18, theta2_std=0.5))
    # 9./10. get_psi's sampler pipeline
    R, L, P = synth.gene_matrices("timeseries", 9, reads=100, T=3, C=3)
    run_pipeline(ref, "pipeline_t3_c3", R, L, P, X3, 3.0, float(synth.default_theta2(X3)), 1500, 1000, 100, 19)
    R, L, P = synth.gene_matrices("static", 10, reads=200, T=1, C=2)
    X1 = np.arange(1)
    run_pipeline(ref, "pipeline_t1_c2", R, L, P, X1, 3.0, float(synth.default_theta2(X1)), 1400, 1000, 100, 20)


if __name__ == "__main__":
    main()
