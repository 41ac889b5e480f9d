"""Fixed small case for ncu: python scripts/ncu_case_c.py <C> <G> <M>  (env DICEGP_WIDE selects the kernel of the throughput pass)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
C, G, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
packed = pack_genes([synth.gene_W("timeseries", g, C=C) for g in range(G)])
s = BatchSampler(0)
s.run(packed, M=M, initial=10 * M, gap=100, seed=1, summarize=False)
print(s.stats)
