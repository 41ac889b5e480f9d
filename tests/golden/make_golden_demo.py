"""Expected posterior means for the bundled demo (demo.sh: separated t1 model and joint model).

    python tests/golden/make_golden_demo.py

The checked-in reference outputs (data/out/*.dice) cannot be reproduced with the reference's
CURRENT code (they differ from it by up to 0.25 in isoform ratio -- they predate the hard-coded
mate_mode="single"), so the expectation is generated here: read->isoform matrices from the
read model that is pinned bit-for-bit to the reference's TranSplice (host_matrices.npz), then
the oracle sampler pipeline (bit-exact restatement of Psi_GP_MH + get_psi's pre-runs),
averaged over 4 seeds.  Output: tests/golden/demo_expected.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from diceseq_b200.annotation import loadgene  # noqa: E402
from diceseq_b200.run_utils import gene_inputs  # noqa: E402
from oracle import psi_gp_mh_oracle as orc  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
for g in genes:
    g.add_premRNA()
out = {}
for name, times, X in (("joint", (1, 2, 3), np.array([1.5, 2.5, 5.0])), ("t1", (1,), np.arange(1))):
    theta2 = ((max(X) - min(X) + 0.1) / 3.0) ** 2
    sam = [[os.path.join(DATA, "reads_t%d.sorted.bam" % t)] for t in times]
    for g in genes:
        R, L, P, count = gene_inputs(g, sam)
        acc = []
        for seed in range(4):
            with np.errstate(all="ignore"):
                res, _var, _ = orc.prerun_and_main([r.copy() for r in R], [l.copy() for l in L], [p.copy() for p in P],
                                                   X, 3.0, theta2, 20000, 1000, 100, rng=np.random.RandomState(seed))
            Psi, m = res[0], res[6]
            acc.append(Psi[int(m / 4):].mean(axis=0))
        out["%s_%s_psi" % (name, g.geneID)] = np.mean(acc, axis=0)
        out["%s_%s_sd" % (name, g.geneID)] = np.std(acc, axis=0)
        out["%s_%s_count" % (name, g.geneID)] = count
        print(name, g.geneID, np.round(out["%s_%s_psi" % (name, g.geneID)][0], 3), np.round(np.std(acc, axis=0)[0], 3))
np.savez_compressed(os.path.join(HERE, "demo_expected.npz"), **out)
