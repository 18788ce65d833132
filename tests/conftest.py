import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (``*.so`` is git-ignored): build it once, in-tree,
    when nvcc is there, so that the suite does not depend on having run build() first."""
    import shutil
    try:
        from pyspade_b200 import _build
        if _build.needs_build() and shutil.which(os.environ.get("NVCC", "nvcc")):
            _build.build()
    except Exception as exc:                                   # noqa: BLE001 - tests will say why
        print(f"[conftest] could not build libspade_b200: {exc}")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_case():
    return np.load(os.path.join(GOLDEN, "small_case.npz"))


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def hypergeom_grid():
    return np.load(os.path.join(GOLDEN, "hypergeom_grid.npz"))


@pytest.fixture(scope="session")
def cli_case():
    return np.load(os.path.join(GOLDEN, "cli_case.npz"))
