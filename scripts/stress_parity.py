"""Host-stream parity stress: many genes, low acceptance (exercises the rejection-chain rounds),
GPU wrapper against the oracle under the identical MT19937 stream.  Prints mismatching genes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np
from diceseq_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
VAR = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5


def case(g):
    rs = np.random.RandomState(g)
    T = int(rs.choice([1, 3, 4, 8])); C = int(rs.choice([2, 3, 4, 6, 8])); reads = int(rs.choice([200, 999, 2000]))
    R, L, P = synth.gene_matrices("timeseries", g, reads=reads, T=T, C=C)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=M, initial=1000, gap=100,
              randomS=7 + g, var=VAR * np.ones(C - 1))
    return R, L, P, kw


def oracle(g):
    from oracle import psi_gp_mh_oracle as orc
    R, L, P, kw = case(g)
    with np.errstate(all="ignore"):
        r = orc.psi_gp_mh(R, L, P, **kw)
    return r[6], r[5], r[3][:r[6]]


if __name__ == "__main__":
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        refs = pool.map(oracle, range(N))
    from diceseq_b200 import Psi_GP_MH
    bad = 0
    accs = []
    for g in range(N):
        R, L, P, kw = case(g)
        got = Psi_GP_MH(R, L, P, **kw)
        m, cnt, pro = refs[g]
        accs.append(cnt / m)
        ok = got[6] == m and got[5] == cnt and np.allclose(got[3][:m], pro, rtol=1e-12, atol=0)
        if not ok:
            bad += 1
            d = np.nonzero(~np.isclose(got[3][:min(m, got[6])], pro[:min(m, got[6])], rtol=1e-12, atol=0))[0]
            print("MISMATCH gene", g, "shape", R[0].shape, len(R), "m", got[6], m, "cnt", got[5], cnt,
                  "first step", d[:1])
    print("genes", N, "bad", bad, "mean acceptance %.3f min %.3f" % (np.mean(accs), np.min(accs)))
