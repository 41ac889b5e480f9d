import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
genes = list(range(40, 44))
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
packed = pack_genes([synth.gene_W("timeseries", g, reads=reads, T=T) for g in genes])
print("C", packed.n_iso)
s = BatchSampler(0)
res = s.run(packed, M=1500, initial=1000, gap=100, seed=5, gene_ids=genes, keep_trace=True)
for r in res:
    print(r.m, r.cnt, r.state, r.flags, r.var, r.psi_mean[:, 0])
print(s.stats)
