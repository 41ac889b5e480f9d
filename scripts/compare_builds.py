"""Same batch, same seed, two library builds (DICEGP_LIB): dump results for comparison."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
G = 400
packed = pack_genes([synth.gene_W("timeseries", g) for g in range(G)])
s = BatchSampler(0)
a = s.run(packed, seed=3)
np.savez(sys.argv[1], m=a.m, cnt=a.cnt, psi=a.psi_mean, lik=a.lik_marg)
print("steps", s.stats["steps"], "acc", (a.cnt / a.m).mean())
tr = {}
for k in (12, 194, 377, 5):
    n = int(min(a.m[k], 1100))
    psi, Y, pro, lik = s.ctx.download_trace(k, n, first_row=0)
    tr["Y%d" % k] = Y; tr["P%d" % k] = pro
np.savez(sys.argv[1].replace(".npz", "_tr.npz"), **tr)
