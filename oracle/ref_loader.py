"""Import the unmodified reference (``/root/reference/pySpade``) in the dev container.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference imports ``tables`` and ``matplotlib`` at module level
(``pySpade/utils.py:17-18``, ``pySpade/DEobs.py:8``); neither is installed here
and neither is used on the hot path, so empty stand-in modules are registered
before import.  Nothing is copied: the reference code runs from where it lies.

``/root/reference`` does not exist on the GPU box; ``available()`` is False
there and every caller must then rely on the committed golden fixtures.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYSPADE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pySpade", "utils.py"))


def _stub_missing(names):
    for name in names:
        try:
            importlib.import_module(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__dict__["__stub__"] = True
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)


def load():
    """Return ``(utils, DEobs, DErand)`` modules of the reference.

    Importing ``DEobs`` / ``DErand`` runs ``np.random.seed(0)`` at module level
    (``DEobs.py:25``, ``DErand.py:27``); callers that need the reference's RNG
    stream must reseed explicitly.
    """
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _stub_missing(["tables", "matplotlib", "matplotlib.pyplot"])
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    utils = importlib.import_module("pySpade.utils")
    deobs = importlib.import_module("pySpade.DEobs")
    derand = importlib.import_module("pySpade.DErand")
    return utils, deobs, derand
