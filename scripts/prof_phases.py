"""Per-phase cycle counts (needs a DG_PROFILE build: DICEGP_LIB=diceseq_b200/libdicegp_prof_dN.so)."""
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
    print("   main kernel %.2f ms, %.3e evals/s" % (st["main_kernel_ms"], st["main_read_evals"] / st["main_kernel_ms"] * 1e3), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "small"):
    run("static", [0], 4000, 100000)
    run("static", range(1184), 3000, 100000)
if which in ("all", "big"):
    for C in (2, 4, 8):
        run("timeseries", range(148), 2000, 100000, C=C)
    for C in (4, 8):
        run("timeseries", range(592), 2000, 100000, C=C)
