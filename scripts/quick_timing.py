"""Quick device-stream timing of the batched pipeline (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes

cfg = sys.argv[1] if len(sys.argv) > 1 else "timeseries"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 512
M = int(sys.argv[3]) if len(sys.argv) > 3 else 20000

def gen(g):
    return synth.gene_W(cfg, g)

t0 = time.time()
with Pool(min(32, os.cpu_count())) as p:
    genes = p.map(gen, range(G), chunksize=16)
packed = pack_genes(genes)
print("gen+pack %.1fs  W=%.2f GB  cpus=%d" % (time.time() - t0, packed.w_bytes / 1e9, os.cpu_count()))
s = BatchSampler(0)
for rep in range(2):
    t0 = time.time()
    res = s.run(packed, M=M, seed=rep, summarize=False)
    wall = time.time() - t0
    st = s.stats
    m = np.array([r.m for r in res]); state = np.array([r.state for r in res])
    print("rep %d wall %.2fs kernel %.1f ms evals %.3e  -> %.3e evals/s (kernel)  steps %d  conv %.2f  m median %d max %d" % (
        rep, wall, st["kernel_ms"], st["read_evals"], st["read_evals"] / st["kernel_ms"] * 1e3, st["steps"],
        (state == 1).mean(), np.median(m), m.max()))
