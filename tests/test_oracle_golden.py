"""The oracle (oracle/psi_gp_mh_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import psi_gp_mh_oracle as orc

CASES = ["static_c3", "joint_t3_c2", "ts_t8_c4", "prerun_c5", "c9_y0", "edge_nan", "zero_row",
         "learn_c2"]
# The vectors were generated in the build container; on another CPU model OpenBLAS may
# pick other kernels, so allow rounding-level drift instead of demanding identical bits.
RTOL = 1e-12


def _close(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_inf | (np.abs(a - b) <= RTOL * np.maximum(np.abs(a), np.abs(b)) + 1e-300) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), "max rel diff %g" % np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    R, L, P, kw, z = load_golden(name)
    with np.errstate(all="ignore"):
        Psi, Y, th2, Pro, Lik, cnt, m = orc.psi_gp_mh(R, L, P, **kw)
    assert m == int(z["m"]) and cnt == int(z["cnt"])
    _close(Pro, z["Pro_all"])
    _close(Lik, z["Lik_all"])
    _close(Psi, z["Psi_all"])
    _close(Y, z["Y_all"])
    _close(th2, z["theta2_all"])


def test_oracle_faithful_cost_mode_same_numbers():
    R, L, P, kw, z = load_golden("joint_t3_c2")
    Psi, Y, th2, Pro, Lik, cnt, m = orc.psi_gp_mh(R, L, P, faithful_cost=True, **kw)
    assert m == int(z["m"]) and cnt == int(z["cnt"])
    _close(Pro, z["Pro_all"])


def test_oracle_cleaning_mutates_like_reference():
    R, L, P, kw, z = load_golden("edge_nan")
    orc.clean_inputs(R, L, P)
    assert L[0][1] == 0.0 and not R[0][:, 1].any() and (P[0][:, 1] == 0).all()
    assert R[0].shape[0] == P[0].shape[0] < z["R0"].shape[0]
    assert R[1].shape[0] == 0
    assert not np.isnan(P[0]).any()


@pytest.mark.parametrize("name", ["pipeline_t3_c3", "pipeline_t1_c2"])
def test_oracle_pipeline(name):
    R, L, P, _kw, z = load_golden(name)
    np.random.seed(int(z["seed"]))
    with np.errstate(all="ignore"):
        res, var, _t2s = orc.prerun_and_main(R, L, P, z["X"], float(z["theta1"]), float(z["theta2"]),
                                             int(z["M"]), int(z["initial"]), int(z["gap"]))
    Psi, Y, th2, Pro, Lik, cnt, m = res
    _close(var, z["var"])
    assert m == int(z["m"]) and cnt == int(z["cnt"])
    _close(Pro, z["Pro_all"])
    psi_mean, psi_95, lik_marg, _s = orc.posterior_summary(Psi, Y, Lik, m)
    _close(psi_mean, z["psi_mean"])
    assert len(psi_95) == Psi.shape[1] and np.isfinite(lik_marg)


def test_recorder_replays():
    """delta / u recorded by the oracle reproduce the accepted trajectory."""
    R, L, P, kw, z = load_golden("joint_t3_c2")
    rec = orc.StepRecord()
    Psi, Y, th2, Pro, Lik, cnt, m = orc.psi_gp_mh(R, L, P, record=rec, **kw)
    delta, u, acc, P_try, L_try = rec.arrays()
    assert delta.shape == (m + 1, 3) and acc.sum() == cnt
    y = np.zeros(3)
    for s in range(m):
        if acc[s]:
            y = delta[s] + y
        np.testing.assert_array_equal(Y[s, 0, :], y)


def test_replay_mode_reproduces_a_recorded_run():
    """`replay=(delta, u)` feeds psi_gp_mh a stream generated elsewhere (tests/test_gpu_device_stream.py
    replays the device's Philox stream); replaying the oracle's own recorded stream must give back the
    very same run, without touching any random state."""
    R, L, P, kw, z = load_golden("ts_t8_c4")
    rec = orc.StepRecord()
    a = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], record=rec, **kw)
    delta, u, _acc, _pt, _lt = rec.arrays()
    C, T = a[0].shape[1], a[0].shape[2]
    delta = delta.reshape(-1, C - 1, T)
    state = np.random.get_state()
    kw2 = dict(kw, randomS=None)
    b = orc.psi_gp_mh(R, L, P, replay=(delta, u), **kw2)
    assert np.array_equal(np.random.get_state()[1], state[1])
    assert a[5] == b[5] and a[6] == b[6]
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
