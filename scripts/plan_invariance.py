"""Same batch, same seed under different pass plans / kernels (environment settings given as arguments, applied one after
the other in ONE process): m, cnt and the posterior means must be identical (the chains see the same counter-based stream
and the same arithmetic per step whichever kernel advances them), the marginal likelihood within 1e-12.
    plan_invariance.py <G> "ENV=v;ENV=v" ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
G = int(sys.argv[1])
def gen(g): return synth.gene_W("timeseries", g)
with Pool(min(32, os.cpu_count())) as p:
    genes = p.map(gen, range(G), chunksize=16)
packed = pack_genes(genes)
s = BatchSampler(0)
ref = s.run(packed, seed=3)
print("default plan: m median %d, converged %.3f" % (np.median(ref.m), (ref.state == 1).mean()), flush=True)
for env in sys.argv[2:]:
    kv = [e.split("=", 1) for e in env.split(";") if e]
    for k, v in kv:
        os.environ[k] = v
    a = s.run(packed, seed=3)
    for k, _ in kv:
        os.environ.pop(k, None)
    same_m, same_cnt = np.array_equal(a.m, ref.m), np.array_equal(a.cnt, ref.cnt)
    same_psi = np.array_equal(a.psi_mean, ref.psi_mean)
    rel = np.max(np.abs(a.lik_marg - ref.lik_marg) / np.abs(ref.lik_marg))
    print("[%s]: m %s cnt %s psi_mean %s lik_marg rel %.2e (%d genes differ in m)" % (
        env, same_m, same_cnt, same_psi, rel, int((a.m != ref.m).sum())), flush=True)
