"""One packed batch, several environment settings (each argument: ENV=value[,ENV=value...]; ':' stands
for ',' inside a value).  Prints the main-run kernel time; DICEGP_VERBOSE=1 adds the per-pass lines."""
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
for spec in sys.argv[3:]:
    for kv in spec.split(","):
        k, v = kv.split("=")
        os.environ[k] = v.replace(":", ",")
    res = s.run(packed, seed=1, summarize=False)
    st = s.stats
    print("%-50s prerun %.1f main %.1f ms  (sum m %d)" % (spec, st["prerun_kernel_ms"], st["main_kernel_ms"], int(sum(r.m for r in res))), flush=True)
