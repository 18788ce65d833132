"""CPU tier: the C-ABI library loads and exports every symbol include/spade.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "spade.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spade_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    from pyspade_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/spade.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.load().spade_abi_version() == _lib.SPADE_ABI_VERSION


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path fails loudly; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyspade_b200.engine import Engine, SpadeError
    with pytest.raises(SpadeError):
        Engine.from_cpm(sp.random(6, 9, 0.4, format="csr", random_state=0))


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "pyspade_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
