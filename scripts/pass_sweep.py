"""Sweep the straggler pass plan (DICEGP_STRAGGLER_STEPS) on one packed batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
cfg = sys.argv[1]; G = int(sys.argv[2])
def gen(g):
    return synth.gene_W(cfg, g)
with Pool(min(32, os.cpu_count())) as p:
    genes = p.map(gen, range(G), chunksize=16)
packed = pack_genes(genes)
s = BatchSampler(0)
s.run(packed, seed=1, summarize=False)
for plan in sys.argv[3:]:
    os.environ["DICEGP_STRAGGLER_STEPS"] = plan
    s.run(packed, seed=1, summarize=False)
    st = s.stats
    print("plan %-12s prerun %.1f ms main %.1f ms  steps %d" % (plan, st["prerun_kernel_ms"], st["main_kernel_ms"], st["steps"]), flush=True)
