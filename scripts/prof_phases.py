"""Per-phase cycle counts (needs the DG_PROFILE build: DICEGP_LIB=diceseq_b200/libdicegp_prof.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes

s = BatchSampler(0)
def run(cfg, genes, M, initial, **kw):
    packed = pack_genes([synth.gene_W(cfg, g, **kw) for g in genes])
    print("== %s G=%d C=%s M=%d initial=%d %s" % (cfg, len(genes), np.unique(packed.n_iso), M, initial, kw), flush=True)
    sys.stderr.flush()
    s.run(packed, M=M, initial=initial, gap=100, seed=1, summarize=False)
    st = s.stats
    print("   kernel %.2f ms, %.3e evals/s" % (st["kernel_ms"], st["read_evals"] / st["kernel_ms"] * 1e3), flush=True)

run("static", [0], 4000, 100000)
run("static", range(4096), 3000, 1000)
for C in (2, 4, 8):
    run("timeseries", [0], 3000, 100000, C=C)
for C in (2, 4, 8):
    run("timeseries", range(296), 2000, 100000, C=C)
