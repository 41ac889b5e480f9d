"""Small fixed case for ncu source-level captures: python scripts/ncu_case.py <config> <G> <M> [C]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
cfg, G, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = {"C": int(sys.argv[4])} if len(sys.argv) > 4 else {}
packed = pack_genes([synth.gene_W(cfg, g, **kw) for g in range(G)])
s = BatchSampler(0)
s.run(packed, M=M, initial=10 * M, gap=100, seed=1, summarize=False)
print(s.stats)
