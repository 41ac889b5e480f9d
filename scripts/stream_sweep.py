"""Sweep helper: time the streaming tier for one C (python scripts/stream_sweep.py C G)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
C, G = int(sys.argv[1]), int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else None
packed = pack_genes([synth.gene_W("timeseries", g, C=C, T=T) for g in range(G)])
s = BatchSampler(0)
for rep in range(2):
    s.run(packed, M=2000, initial=100000, gap=100, seed=1, summarize=False)
st = s.stats
print("T=%s C=%d G=%d maxres=%s tiers=%s/%s: main %.2f ms  %.3e evals/s" % (T, C, G, os.environ.get("DICEGP_MAX_RESIDENT_TIER"), os.environ.get("DICEGP_TIER_THREADS"), os.environ.get("DICEGP_TIER_BPS"),
      st["main_kernel_ms"], st["main_read_evals"] / st["main_kernel_ms"] * 1e3))
