"""Run in a subprocess with DICEGP_WIDE_MIN_BYTES set: step parity of the cluster tier (one
chain over 2/4/8 thread blocks with a DSMEM exchange) against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diceseq_b200 import Psi_GP_MH, synth  # noqa: E402
from diceseq_b200.batch import BatchSampler  # noqa: E402
from diceseq_b200.packer import pack_genes  # noqa: E402
from oracle import psi_gp_mh_oracle as orc  # noqa: E402

# W bytes: 24 KB (cluster of 2), 96 KB (4), 512 KB (8), 12 isoforms (generic kernel, plain loads)
SHAPES = [("static", 31, 1000, 1, 3, 1300, 1000, 100), ("timeseries", 32, 1000, 4, 3, 1100, 1000, 100),
          ("timeseries", 33, 1000, 8, 8, 300, 200, 100), ("timeseries", 34, 1501, 2, 12, 300, 200, 100)]
for cfg, g, reads, T, C, M, ini, gap in SHAPES:
    R, L, P = synth.gene_matrices(cfg, g, reads=reads, T=T, C=C)
    kw = dict(theta1=3.0, theta2=float(synth.default_theta2(np.arange(T))), M=M, initial=ini, gap=gap, randomS=100 + g)
    with np.errstate(all="ignore"):
        ref = orc.psi_gp_mh([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P], **kw)
    got = Psi_GP_MH(R, L, P, **kw)
    assert got[6] == ref[6] and got[5] == ref[5], (cfg, T, C, got[5:], ref[5:])
    for a, b in ((got[3], ref[3]), (got[4], ref[4]), (got[0], ref[0]), (got[1], ref[1])):
        assert np.all(np.abs(a - b) <= 1e-12 * np.maximum(np.abs(a), np.abs(b)) + 1e-300), (cfg, T, C)
    print("ok", cfg, T, C, "m", got[6], "cnt", got[5])

# device stream + persistent queue on clusters: results equal the non-cluster run of the same seed
genes = list(range(40, 46))
packed = pack_genes([synth.gene_W("timeseries", g, reads=300, T=4) for g in genes])
s = BatchSampler(0)
res = s.run(packed, M=2500, initial=1000, gap=100, seed=5, gene_ids=genes, keep_trace=True)
np.savez(sys.argv[1], m=np.array([r.m for r in res]), cnt=np.array([r.cnt for r in res]),
         lik=np.concatenate([r.trace[3] for r in res]), launches=s.stats["launches"])
print("batch ok", [r.m for r in res])
