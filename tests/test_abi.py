"""The C-ABI library loads and exports every symbol include/dicegp.h declares (CPU only:
no compute call is made; without a GPU dicegp_create must fail loudly, not fall back)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "diceseq_b200", "libdicegp.so")):
        ge.build()
    from diceseq_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from diceseq_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dicegp.h")).read()
    declared = set(re.findall(r"\b(dicegp_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    from diceseq_b200 import _lib
    assert lib.dicegp_abi_version() == _lib.ABI_VERSION == 2
    assert _lib.shape_limits() == {"max_isoforms": 32, "max_times": 32, "max_latent": 448}


def test_chain_desc_layout():
    from diceseq_b200 import _lib
    assert ctypes.sizeof(_lib.ChainDesc) == 72       # 6 x int32 + 5 x int64 + double


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from diceseq_b200 import _lib
    with pytest.raises(_lib.DiceGPError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the package must not import, call or load it."""
    pkg = os.path.join(ROOT, "diceseq_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|psi_gp_mh_oracle|oracle/_ref", re.M)
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f
