"""Golden read->isoform matrices from the UNMODIFIED reference read model.

    python tests/golden/make_golden_host.py

The reference's `TranSplice` needs pysam (absent here).  A stand-in `pysam` module backed by
`diceseq_b200.bam.BamFile` (raw BGZF/BAM decoding only) plus a stub `matplotlib` let the
reference package import; everything after the raw records -- region filters, mate matching,
compatibility, fragment-length model, probabilities -- is then the reference's own code
(diceseq/utils/sam_utils.py, tran_utils.py, gtf_utils.py).  Output: tests/golden/host_matrices.npz
with Rmat / proU / efflen_unif / count per gene and time point for the bundled yeast data
(tests/data, copied from /root/reference/data), in the CLI's mode (mate_mode="single", unif).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from diceseq_b200 import bam as mybam  # noqa: E402

# ---- stand-in modules -------------------------------------------------------------------
pysam = types.ModuleType("pysam")
pysam.AlignmentFile = lambda path, mode="rb": mybam.load_samfile(path)
pysam.idxstats = lambda path: "\n".join("%s\t%d\t%d\t%d" % r for r in mybam.load_samfile(path).idxstats())
pysam.FastaFile = object
sys.modules["pysam"] = pysam
for name in ("matplotlib", "matplotlib.pyplot", "pylab"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, "/root/reference")
import diceseq  # noqa: E402  (the unmodified reference package)
from diceseq.utils.gtf_utils import loadgene  # noqa: E402
from diceseq.utils.tran_utils import TranSplice  # noqa: E402
from diceseq.utils.sam_utils import load_samfile  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
out = {}
for g in genes:
    g.add_premRNA()
    for ti in (1, 2, 3, 4):
        sam = load_samfile(os.path.join(DATA, "reads_t%d.sorted.bam" % ti))
        t = TranSplice(g)
        t.set_reads(sam)
        t.get_ready("unif", None, None, "single", 200)
        key = "%s_t%d" % (g.geneID, ti)
        out[key + "_R"] = t.Rmat
        out[key + "_proU"] = t.proU
        out[key + "_efflen"] = t.efflen_unif
        t.get_ready("unif", None, None, "pair", 200)
        out[key + "_count"] = np.sum(t.Rmat.sum(axis=1) > 0)
        print(key, t.Rmat.shape, int(out[key + "_count"]), t.efflen_unif)
out["gene_ids"] = np.array([g.geneID for g in genes])
out["tran_ids"] = np.array([",".join(t.tranID for t in g.trans) for g in genes])
out["tran_len"] = np.array([",".join(str(t.tranL) for t in g.trans) for g in genes])
np.savez_compressed(os.path.join(HERE, "host_matrices.npz"), **out)
