"""Random single-time-point genes (odd / tiny / empty read counts, 2-12 isoforms, zero rows that give
-inf likelihoods) through the warp-per-chain kernel and the general kernel: m, cnt, state, flags and
the posterior means must agree bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.RandomState(5)
genes = []
for g in range(n):
    C = int(rng.randint(2, 13))
    reads = int(rng.choice([1, 2, 3, 31, 64, 65, 255, 999, 1000, 1501, 3000]))
    W, lens = synth.gene_W("static", 1000 + g, reads=reads, C=C)
    if g % 17 == 0 and reads > 3:
        W[0][1, :] = 0.0                         # a read no isoform explains: log(0) = -inf (model_GP.py:344)
    genes.append((W, lens))
packed = pack_genes(genes)
s = BatchSampler(0)
out = {}
for mode in ("0", "1"):
    os.environ["DICEGP_PRERUN_KERNEL"] = mode
    res = s.run(packed, M=1500, initial=300, gap=50, seed=9)
    out[mode] = (res.m.copy(), res.cnt.copy(), res.state.copy(), res.flags.copy(), res.psi_mean.copy(), res.var.copy())
    print("mode %s: main %.2f ms, converged %.3f, flagged %d" % (mode, s.stats["main_kernel_ms"], (res.state == 1).mean(), (res.flags != 0).sum()))
names = ("m", "cnt", "state", "flags", "psi_mean", "var")
for k, nm in enumerate(names):
    a, b = out["0"][k], out["1"][k]
    print("%-8s identical: %s" % (nm, np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)))
