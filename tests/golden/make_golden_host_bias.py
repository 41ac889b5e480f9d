"""Golden read->isoform matrices of the BIAS modes from the UNMODIFIED reference read model.

    python tests/golden/make_golden_host_bias.py

Same stand-ins as make_golden_host.py (a `pysam` module backed by raw BAM decoding, a stub matplotlib);
`pysam.FastaFile` serves the seeded synthetic genome of synth_fasta.py, the bias parameters are the
reference's bundled data/bias/yeast_4tU.bias (copied to tests/data).  Everything else -- set_sequence,
set_bias("seq"), the end5 / end3 / both weighting, the bias-weighted effective lengths -- is the reference's
own code (diceseq/utils/tran_utils.py, bias_utils.py), driven the way get_psi drives it
(run_utils.py:62-87, mate_mode="single").  Output: tests/golden/host_matrices_bias.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from diceseq_b200 import bam as mybam  # noqa: E402
import synth_fasta  # noqa: E402

GENOME = {}


class _Fasta:
    def __init__(self, path):
        pass

    def fetch(self, ref, start, stop):
        return GENOME[ref][max(start, 0):stop]


pysam = types.ModuleType("pysam")
pysam.AlignmentFile = lambda path, mode="rb": mybam.load_samfile(path)
pysam.FastaFile = _Fasta
sys.modules["pysam"] = pysam
for name in ("matplotlib", "matplotlib.pyplot", "pylab"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, "/root/reference")
import diceseq  # noqa: E402,F401  (the unmodified reference package)
from diceseq.utils.gtf_utils import loadgene  # noqa: E402
from diceseq.utils.tran_utils import TranSplice  # noqa: E402
from diceseq.utils.sam_utils import load_samfile  # noqa: E402
from diceseq.utils.bias_utils import BiasFile, FastaFile  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
GENOME.update(synth_fasta.genome_for(genes))
fasta = FastaFile("synthetic")
out = {}
for g in genes:
    g.add_premRNA()
    for ti in (1, 3):
        for mode in ("end5", "end3", "both"):
            sam = load_samfile(os.path.join(DATA, "reads_t%d.sorted.bam" % ti))
            bias = BiasFile(os.path.join(DATA, "yeast_4tU.bias"))
            t = TranSplice(g)
            t.set_reads(sam)
            t.set_sequence(fasta)
            t.set_bias(bias, "seq")
            t.get_ready(mode, None, None, "single", 200)
            key = "%s_t%d_%s" % (g.geneID, ti, mode)
            out[key + "_R"] = t.Rmat
            out[key + "_proB"] = t.proB
            out[key + "_efflen_bias"] = t.efflen_bias
            out[key + "_efflen_unif"] = t.efflen_unif
            print(key, t.Rmat.shape, t.efflen_bias, np.nanmax(t.proB) if t.Rmat.any() else None)
# one case with a fixed fragment-length model (FLmean / FLstd from the bias file, run_utils.py:74-77)
g = genes[1]
bias = BiasFile(os.path.join(DATA, "yeast_4tU.bias"))
t = TranSplice(g)
t.set_reads(load_samfile(os.path.join(DATA, "reads_t2.sorted.bam")))
t.set_sequence(fasta)
t.set_bias(bias, "seq")
t.get_ready("both", bias.flen_mean, bias.flen_std, "single", 200)
out["flen_case_gene"] = np.array(g.geneID)
out["flen_case_proB"] = t.proB
out["flen_case_efflen_bias"] = t.efflen_bias
out["flen_mean_std"] = np.array([bias.flen_mean, bias.flen_std])
out["gene_ids"] = np.array([g.geneID for g in genes])
np.savez_compressed(os.path.join(HERE, "host_matrices_bias.npz"), **out)
