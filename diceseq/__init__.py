"""Import shim: `import diceseq` resolves to the B200-native implementation of the Psi_GP_MH path.

The names below are the ones of the reference's `diceseq/__init__.py` that this repository implements
(sampler, GP helpers, annotation and BAM access); `diceseq.diceseq.main` is the command line.  The
remaining reference exports (static models, Bayes factors, plotting and the bias-training tools) are outside
the accelerated path: asking for one raises an ImportError that says so instead of an AttributeError.
"""
from diceseq_b200 import __version__  # noqa: F401
from diceseq_b200.model_GP import GP_K, Geweke_Z, Psi_GP_MH, normal_pdf  # noqa: F401
from diceseq_b200.annotation import Gene, Transcript, loadgene  # noqa: F401
from diceseq_b200.bam import load_samfile  # noqa: F401
from diceseq_b200.read_model import fetch_reads  # noqa: F401
from diceseq_b200.bias import BiasFile, FastaFile  # noqa: F401
from diceseq_b200.read_model import TranSplice  # noqa: F401

load_annotation = loadgene

_NOT_PORTED = ("ReadSet", "DiceFile", "SampleFile", "TranUnits", "mcmc_sampler", "miso_BF", "dicediff_BF",
               "get_BioVar", "Psi_MCMC_MH", "Psi_analytic", "Psi_junction")


def __getattr__(name):
    if name in _NOT_PORTED:
        raise ImportError("diceseq.%s belongs to the reference package only; this build accelerates the "
                          "Psi_GP_MH path (sampler, read model, CLI)" % name)
    raise AttributeError(name)
