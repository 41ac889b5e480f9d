"""Phase profile of one (C, G) time-series group (DG_PROFILE build via DICEGP_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
C, G = int(sys.argv[1]), int(sys.argv[2])
packed = pack_genes([synth.gene_W("timeseries", g, C=C) for g in range(G)])
s = BatchSampler(0)
s.run(packed, M=2000, initial=100000, gap=100, seed=1, summarize=False)
print("C=%d G=%d main %.2f ms" % (C, G, s.stats["main_kernel_ms"]))
