"""diceseq_b200 -- B200-native DICEseq GP/MH isoform sampler (the Psi_GP_MH hot path).

Public surface mirrors what the reference exports for this path
(diceseq/__init__.py:15): Psi_GP_MH, normal_pdf, GP_K, Geweke_Z.
"""
from .model_GP import GP_K, Geweke_Z, Psi_GP_MH, normal_pdf  # noqa: F401

__version__ = "0.1.0"
