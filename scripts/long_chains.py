"""Per-step latency of the chains that never converge (they bound the straggler pass)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
G = int(sys.argv[1])
def gen(g):
    return synth.gene_W("timeseries", g)
with Pool(min(32, os.cpu_count())) as p:
    genes = p.map(gen, range(G), chunksize=16)
s = BatchSampler(0)
res = s.run(pack_genes(genes), seed=1, summarize=False)
state = np.array([r.state for r in res]); cnt = np.array([r.cnt for r in res]); m = np.array([r.m for r in res])
C = np.array([g[0][0].shape[1] for g in genes])
bad = np.where(state != 1)[0]
print("never converged:", len(bad), "C:", np.bincount(C[bad], minlength=9), "accept rate:", np.round(cnt[bad] / 20000.0, 3))
print("all: accept rate mean %.3f" % (cnt / np.maximum(m, 1)).mean())
os.environ["DICEGP_STRAGGLER_STEPS"] = "100"
for name, ids in (("never-converged", bad), ("ordinary, same C", None)):
    if ids is None:
        ids = np.concatenate([np.where((C == c) & (state == 1))[0][:n] for c, n in enumerate(np.bincount(C[bad], minlength=9)) if n])
    sub = pack_genes([genes[i] for i in ids])
    for rep in range(2):
        r2 = s.run(sub, seed=1, gene_ids=ids, initial=10**6, summarize=False)
    os.environ["DICEGP_STRAGGLER_STEPS"] = "4096"
    r3 = s.run(sub, seed=1, gene_ids=ids, summarize=False)
    print(name, "with Geweke checks: main %.1f ms, m = %s" % (s.stats["main_kernel_ms"], sorted(set(int(r.m) for r in r3))[:6]))
    os.environ["DICEGP_STRAGGLER_STEPS"] = "100"
    print(name, "n=%d main %.1f ms for 20000 steps each -> %.2f us/step" % (len(ids), s.stats["main_kernel_ms"], s.stats["main_kernel_ms"] * 1e3 / 20000))
