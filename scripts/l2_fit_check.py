"""Is the streamed tier bound by DRAM?  Same isoform count, same number of chains, fewer reads per
gene: when the W of all co-resident chains fits the 126 MB L2 the DRAM traffic disappears."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
C, G = int(sys.argv[1]), int(sys.argv[2])
s = BatchSampler(0)
for reads in (1000, 700, 500, 350):
    packed = pack_genes([synth.gene_W("timeseries", g, C=C, reads=reads) for g in range(G)])
    for rep in range(2):
        s.run(packed, M=2000, initial=100000, gap=100, seed=1, summarize=False)
    st = s.stats
    print("C=%d G=%d reads/time point %d (W %.0f KB per chain): main %.2f ms  %.3e evals/s" % (
        C, G, reads, 8 * C * 8 * reads / 1024, st["main_kernel_ms"], st["main_read_evals"] / st["main_kernel_ms"] * 1e3), flush=True)
