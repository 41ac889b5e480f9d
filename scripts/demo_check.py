import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200.annotation import loadgene
from diceseq_b200.run_utils import gene_inputs
from diceseq_b200.packer import clean_inputs, pack_genes
from diceseq_b200.batch import BatchSampler
D = "tests/data"
genes = loadgene(D + "/yeast_RNA_splicing.gtf")
X = np.array([1.5, 2.5, 5.0]); theta2 = ((5 - 1.5 + 0.1) / 3) ** 2
sam = [[D + "/reads_t%d.sorted.bam" % t] for t in (1, 2, 3)]
ins = []
for g in genes:
    g.add_premRNA()
    R, L, P, c = gene_inputs(g, sam)
    clean_inputs(R, L, P)
    ins.append(([r * p for r, p in zip(R, P)], L))
packed = pack_genes(ins)
E = np.load("tests/golden/demo_expected.npz")
s = BatchSampler(0)
acc = []
for seed in range(8):
    res = s.run(packed, X=X, theta1=3.0, theta2=theta2, M=20000, initial=1000, gap=100, seed=seed)
    acc.append([r.psi_mean[0] for r in res])
    print(seed, [(r.m, r.cnt) for r in res], np.round(res[3].psi_mean[0], 3), np.round(res[3].var, 3))
acc = np.array(acc)
for k, g in enumerate(genes):
    print(g.geneID, "gpu mean", np.round(acc[:, k].mean(0), 3), "sd", np.round(acc[:, k].std(0), 3), "oracle", np.round(E["joint_%s_psi" % g.geneID][0], 3))
