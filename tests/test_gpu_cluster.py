"""Cluster tier (large genes: one chain over a thread-block cluster, partial likelihoods
exchanged through distributed shared memory).  The tier threshold is lowered through
DICEGP_WIDE_MIN_BYTES in a subprocess so that ordinary test genes take the cluster path."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(ROOT, "tests", "helpers", "cluster_parity.py")


def _run(env_extra, out):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, HELPER, out], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return np.load(out)


def test_cluster_tier_step_parity_and_same_results_as_single_block(tmp_path):
    wide = _run({"DICEGP_WIDE_MIN_BYTES": "20000"}, str(tmp_path / "wide.npz"))
    plain = _run({}, str(tmp_path / "plain.npz"))
    # same device stream, same arithmetic per step: identical chains whether or not a cluster ran them,
    # up to the reduction order of the partial likelihoods (1e-12 relative on every stored log-likelihood)
    assert np.array_equal(wide["m"], plain["m"]) and np.array_equal(wide["cnt"], plain["cnt"])
    assert np.all(np.abs(wide["lik"] - plain["lik"]) <= 1e-12 * np.abs(plain["lik"]))
