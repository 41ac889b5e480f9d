"""Warp-per-chain pre-run kernel against the general chain kernel (DICEGP_PRERUN_KERNEL=0) on the
same batch and seed: jump variances, main-run m / cnt, timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
cfg = sys.argv[1]; G = int(sys.argv[2]); M = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
def gen(g):
    return synth.gene_W(cfg, g)
with Pool(min(32, os.cpu_count())) as p:
    genes = p.map(gen, range(G), chunksize=16)
packed = pack_genes(genes)
s = BatchSampler(0)
out = {}
for mode in ("0", "1", "0", "1"):
    os.environ["DICEGP_PRERUN_KERNEL"] = mode
    res = s.run(packed, seed=3, M=M, summarize=True)
    out[mode] = (res.var.copy(), res.m.copy(), res.cnt.copy(), res.psi_mean.copy())
    print("kernel %s: prerun %.2f ms (%d launches, %.3e evals)  main %.1f ms" % (
        "warp-per-chain" if mode == "1" else "general", s.stats["prerun_kernel_ms"], s.stats["prerun_launches"],
        s.stats["prerun_read_evals"], s.stats["main_kernel_ms"]), flush=True)
a, b = out["0"], out["1"]
rel = np.abs(a[0] - b[0]) / np.maximum(np.abs(a[0]), 1e-300)
rel = rel[a[0] != 0]
print("jump variances: max rel diff %.3e, share within 1e-10: %.4f, bit-identical: %.4f" % (rel.max(), (rel < 1e-10).mean(), (rel == 0).mean()))
print("main run: same m %.4f, same cnt %.4f, psi_mean max abs diff (genes with same m) %.3e" % (
    (a[1] == b[1]).mean(), (a[2] == b[2]).mean(),
    0.0 if not (a[1] == b[1]).any() else np.abs(a[3] - b[3]).max()))
