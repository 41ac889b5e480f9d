"""Phase profile of the straggler (latency-shape) pass for one isoform count (DG_PROFILE build via DICEGP_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
C, G = int(sys.argv[1]), int(sys.argv[2])
initial = int(sys.argv[3]) if len(sys.argv) > 3 else 10**6
os.environ["DICEGP_STRAGGLER_STEPS"] = "100"
packed = pack_genes([synth.gene_W("timeseries", g, C=C) for g in range(G)])
s = BatchSampler(0)
for rep in range(2):
    s.run(packed, M=20000, initial=initial, gap=100, seed=1, summarize=False)
print("C=%d G=%d main %.2f ms -> %.2f us/step" % (C, G, s.stats["main_kernel_ms"], s.stats["main_kernel_ms"] * 1e3 / 20000))
